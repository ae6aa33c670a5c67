set -x
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 1500 python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_church.json; cat gpurun_out/bench_church.json | cut -c1-400
timeout 900 python bench.py --workload imagenet --steps 10 2>/dev/null | tail -1 > gpurun_out/bench_imagenet.json; cut -c1-300 gpurun_out/bench_imagenet.json
EDADM_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_church_dram.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-recon --no-graph > gpurun_out/ll.log 2>&1
EDADM_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_imagenet_dram.csv python bench.py --workload imagenet --steps 1 --warmup 1 --no-cpu-baseline --no-recon --no-graph > gpurun_out/ll_in.log 2>&1
W4=0 ONLY="church 32x32" NOLIB=1 REPS=3 timeout 500 ncu --set full --clock-control none --import-source on -k regex:qgemm_i8 -s 2 -c 1 -o gpurun_out/qgemm_c192 -f python scratch/bench_gemm.py > gpurun_out/qg_ncu.log 2>&1
ONE=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:qattn_kernel -s 3 -c 1 -o gpurun_out/qattn_t1024 -f python scratch/bench_attn.py > gpurun_out/qattn_ncu.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:act_quant_nhwc_tma -s 2 -c 1 -o gpurun_out/actq_tma -f python scratch/one_actq.py > gpurun_out/actq_ncu.log 2>&1
ls gpurun_out/*.ncu-rep
