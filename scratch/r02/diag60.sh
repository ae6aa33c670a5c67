#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python scratch/r02/trace_gemm2.py 2>&1 | grep -v Warn | tail -32
