#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --tb=short -k "conv_bf16x3 or linear_bf16x3" 2>&1 | tail -8
timeout 600 python scratch/r02/prof_recon_tp.py resblock 2>&1 | tail -32
