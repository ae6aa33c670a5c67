#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python scratch/r02/trace_cmp.py 2>&1 | tail -12
