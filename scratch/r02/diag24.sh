#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for cfg in "4 3" "3 4" "2 4" "3 3" "4 2"; do set -- $cfg; echo "== stages $1 blocks/SM $2"; EDADM_ACTQ_STAGES=$1 EDADM_ACTQ_BPS=$2 timeout 300 python scratch/r02/actq_sweep.py; done
