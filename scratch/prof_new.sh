EDADM_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"layernorm_quant_rows|geglu_quant_rows" -c 2 -o gpurun_out/tf_producers -f python bench.py --workload imagenet --steps 1 --warmup 1 --no-cpu-baseline --no-recon --no-graph > gpurun_out/tfp.log 2>&1
tail -2 gpurun_out/tfp.log | cut -c1-150
