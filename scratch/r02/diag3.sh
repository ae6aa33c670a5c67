#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/r02
for dbg in 0 4; do echo "#### debug=$dbg"; EDADM_GEMM_DEBUG=$dbg python scratch/r02/trace_gemm.py 2>&1 | grep -E "==|tile [23]:"; done > gpurun_out/r02/trace_gemm_dbg4.txt 2>&1
cat gpurun_out/r02/trace_gemm_dbg4.txt
