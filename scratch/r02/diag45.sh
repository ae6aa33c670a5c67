#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --tb=line -k "conv_bf16x3" 2>&1 | tail -12
