timeout 1500 python bench.py 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_church.json
cat gpurun_out/bench_church.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.json
cat gpurun_out/bench_ref.json
