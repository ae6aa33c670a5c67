#!/bin/bash
mkdir -p gpurun_out/r02
timeout 60 python scratch/r02/adam_bench.py > gpurun_out/r02/adam_bench.txt 2>&1
timeout 200 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02/pytest_gpu_final2.txt 2>&1
echo rc=$? >> gpurun_out/r02/pytest_gpu_final2.txt
tail -2 gpurun_out/r02/adam_bench.txt; tail -5 gpurun_out/r02/pytest_gpu_final2.txt
