#!/bin/bash
# round-2 baseline diagnostics: int8 tensor peak, per-shape GEMM numbers, ImageNet step vs batch (L2 residency)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02/smi.txt
python scratch/mma_peak.py > gpurun_out/r02/mma_peak.txt 2>&1
NOLIB=1 python scratch/bench_gemm.py > gpurun_out/r02/bench_gemm.txt 2>&1
for b in 128 32 16; do
  python bench.py --workload imagenet --batch $b --no-recon --no-cpu-baseline --steps 10 > gpurun_out/r02/bench_imagenet_b$b.json 2> gpurun_out/r02/bench_imagenet_b$b.err
done
python bench.py --workload church --no-recon --no-cpu-baseline --steps 10 > gpurun_out/r02/bench_church.json 2> gpurun_out/r02/bench_church.err
tail -n 3 gpurun_out/r02/*.txt gpurun_out/r02/*.json
