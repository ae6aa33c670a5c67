timeout 1500 python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_church.json; cut -c1-250 gpurun_out/bench_church.json
timeout 900 python bench.py --workload imagenet --steps 10 2>/dev/null | tail -1 > gpurun_out/bench_imagenet.json; cut -c1-250 gpurun_out/bench_imagenet.json
EDADM_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_church_dram.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-recon --no-graph > gpurun_out/ll.log 2>&1
EDADM_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_imagenet_dram.csv python bench.py --workload imagenet --steps 1 --warmup 1 --no-cpu-baseline --no-recon --no-graph > gpurun_out/ll_in.log 2>&1
echo done
