#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_api.py -x -q -m gpu --tb=short -k export 2>&1 | tail -15
