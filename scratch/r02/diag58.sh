#!/bin/bash
cd "$GRAFT_REPO_ROOT"
EDADM_GEMM_V1=1 timeout 600 python scratch/r02/gemm_table.py imagenet 2>&1 | grep -v Warn | grep "eager step\|    27 \|   192  1 \|   384  1 -\|   576  1 -\|   960  1 -"
