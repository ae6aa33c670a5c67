EDADM_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_imagenet.csv python bench.py --workload imagenet --steps 1 --warmup 1 --no-cpu-baseline --no-recon --no-graph > gpurun_out/ll_in.log 2>&1
tail -1 gpurun_out/ll_in.log | cut -c1-200
