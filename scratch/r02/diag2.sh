#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r02
timeout 300 python scratch/mma_peak.py > gpurun_out/r02/mma_peak2.txt 2>&1
W4=0 NOLIB=1 python scratch/bench_gemm.py > gpurun_out/r02/bench_gemm_s8.txt 2>&1
cat gpurun_out/r02/mma_peak2.txt
