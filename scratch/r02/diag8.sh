#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/r02
python scratch/r02/dbg_graph.py 2>&1 | grep -v Warn | tail -24
timeout 900 python -m pytest tests/test_gpu_api.py -q -m gpu --tb=short -k walk 2>&1 | tail -15
