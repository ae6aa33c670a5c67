#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/r02
for c in 1 2; do echo "#### v2 ctas=$c"; EDADM_GEMM_CTAS=$c timeout 200 python scratch/r02/gemm_check.py 2>&1 | tail -16; done > gpurun_out/r02/gemm_check_v2.txt 2>&1
cat gpurun_out/r02/gemm_check_v2.txt
