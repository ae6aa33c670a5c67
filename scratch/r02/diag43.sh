#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --tb=short 2>&1 | tail -3
timeout 300 python scratch/r02/actq_sweep.py 2>&1 | grep -v Warn | head -2
