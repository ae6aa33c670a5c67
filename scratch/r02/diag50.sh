#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=line -k "bf16x3" 2>&1 | tail -4
timeout 600 python scratch/r02/wgrad_bench.py 2>&1 | grep -v Warn | tail -6
