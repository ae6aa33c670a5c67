#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/r02
for u in resblock transformer_block; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02/launches_recon3_$u.csv python scratch/r02/prof_recon.py $u > gpurun_out/r02/ll_recon3_$u.log 2>&1; tail -1 gpurun_out/r02/ll_recon3_$u.log
python scratch/r02/agg_launches.py gpurun_out/r02/launches_recon3_$u.csv gpurun_out/r02/recon3_${u}_summary.json gpurun_out/r02/recon3_${u}_iteration.csv | head -14
done
