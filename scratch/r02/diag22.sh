#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_api.py -x -q -m gpu --tb=short 2>&1 | tail -15
