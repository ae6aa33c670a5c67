#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for d in 0 1 2 3 4 7 8 15; do echo "== EDADM_EPI_DEBUG=$d"; EDADM_EPI_DEBUG=$d timeout 300 python scratch/r02/codes_bench.py 2>&1 | grep -v Warn | grep "3072\|131072, 384"; done
