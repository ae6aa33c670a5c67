python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench2_err.log | tail -1 > gpurun_out/bench_church_2gpu.json
cut -c1-300 gpurun_out/bench_church_2gpu.json; tail -3 gpurun_out/bench2_err.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
