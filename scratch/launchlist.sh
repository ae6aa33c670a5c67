EDADM_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_church.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-recon --no-graph > gpurun_out/ll.log 2>&1
tail -2 gpurun_out/ll.log | cut -c1-200
