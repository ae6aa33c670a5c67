#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --tb=short -k "qgemm or codes or fusion or geglu or transformer or split" 2>&1 | tail -3
echo "== weight-resident"; timeout 300 python scratch/r02/codes_bench.py 2>&1 | grep -v Warn
echo "== EDADM_GEMM_NO_BRES=1"; EDADM_GEMM_NO_BRES=1 timeout 300 python scratch/r02/codes_bench.py 2>&1 | grep -v Warn
