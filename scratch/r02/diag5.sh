#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "epilogue_fusions" 2>&1 | tail -5 > gpurun_out/r02/pytest_gemm2.txt
cat gpurun_out/r02/pytest_gemm2.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02/pytest_all.txt; cat gpurun_out/r02/pytest_all.txt
