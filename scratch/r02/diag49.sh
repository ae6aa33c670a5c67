#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python scratch/r02/wgrad_bench.py 2>&1 | grep -v Warn | tail -6
