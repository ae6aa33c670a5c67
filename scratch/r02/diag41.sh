#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --tb=short -k "attn or attention or transformer" 2>&1 | tail -3
for r in 0 1; do echo "== EDADM_ATTN_QRES=$r"; EDADM_ATTN_QRES=$r timeout 300 python scratch/r02/attn_bench.py 2>&1 | grep -v Warn; done
