#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for bn in 64 96 128; do echo "== bn $bn"; EDADM_WGRAD_BN=$bn timeout 600 python scratch/r02/wgrad_bench.py 2>&1 | grep -v Warn | grep "576, 576\|192, 192, 64\|384, 384"; done
