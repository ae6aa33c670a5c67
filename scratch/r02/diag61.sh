#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu --tb=short -k "qgemm or codes or fusion or geglu or transformer or split or lazy" 2>&1 | tail -3
timeout 300 python scratch/r02/codes_bench.py 2>&1 | grep -v Warn
timeout 600 python scratch/r02/gemm_table.py imagenet 2>&1 | grep -v Warn | head -14
