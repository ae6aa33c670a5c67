#!/bin/bash
# launch lists (device time + DRAM bytes per launch) of one eager step, ImageNet b128 and church b100
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/r02
for wl in imagenet church; do
EDADM_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02/launches_$wl.csv python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline --no-recon --no-graph --no-secondary > gpurun_out/r02/ll_$wl.log 2>&1
tail -1 gpurun_out/r02/ll_$wl.log | cut -c1-150
done
ls -la gpurun_out/r02/launches_*.csv
