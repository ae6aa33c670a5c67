#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --tb=line -k "split_shortcut or lazy or conv_bf16x3 or upsample or ties or epilogue_fusions or codes or small_n or qattn or reproducible" > gpurun_out/memcheck_r02.log 2>&1
grep -n "=========" gpurun_out/memcheck_r02.log | grep -v "Host Frame\|Saved host" | head -20
