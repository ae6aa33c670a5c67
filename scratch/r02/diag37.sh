#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for c in 1 2; do echo "== EDADM_GEMM_CTAS=$c"; EDADM_GEMM_CTAS=$c timeout 300 python scratch/r02/codes_bench.py 2>&1 | grep -v Warn; done
echo "== default"; timeout 300 python scratch/r02/codes_bench.py 2>&1 | grep -v Warn
