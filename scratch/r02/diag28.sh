#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python scratch/r02/dbg_lazy.py 2>&1 | grep -v Warning | head -8
