#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python scratch/r02/gemm_table.py imagenet 2>&1 | grep -v Warn | head -26
