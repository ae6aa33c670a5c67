#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for a in "2 64 128 64 64 3 4" "4 128 128 32 32 3 2" "4 128 128 16 16 3 1" "8 128 128 8 8 3 1" "2 64 72 128 128 1 16" "8 192 192 32 32 3 16"; do timeout 120 python scratch/r02/dbg_wgrad.py $a 2>&1 | grep -v Warn | tail -1; done
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=line -k "conv_bf16x3" 2>&1 | tail -5
