#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python scratch/r02/dbg_det.py 2>&1 | grep -v Warning | grep "knob off: None"
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --tb=short -k "lazy or reproducible or search or upsample" 2>&1 | tail -6
