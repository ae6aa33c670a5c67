#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -s --tb=short -k "bf16x3" 2>&1 | grep -E "rel-L2|passed|failed|Error|assert" | head -20
