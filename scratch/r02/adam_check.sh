#!/bin/bash
# fused Adam: kernel test, every reconstruction test (they now run on it), and the recon half of the bench before/after
mkdir -p gpurun_out/r02
timeout 170 python -m pytest tests/ -x -q -m gpu -s -k "fused_adam or recon or trace or walk or memoised or checkpointed" > gpurun_out/r02/pytest_adam.txt 2>&1
echo rc=$? >> gpurun_out/r02/pytest_adam.txt
tail -25 gpurun_out/r02/pytest_adam.txt
