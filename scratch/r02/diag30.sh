#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python scratch/r02/dbg_det.py 2>&1 | grep -v Warning | tail -20
