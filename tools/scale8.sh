#!/bin/bash
# 8-GPU bench lines (run under gpurun --gpus 8): config 2 weak scaling, config 5 strong scaling
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r1b_bench_c2_8gpu.json
cut -c1-200 gpurun_out/r1b_bench_c2_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 8 --workload c5 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r1b_bench_c5_8gpu.json
cut -c1-200 gpurun_out/r1b_bench_c5_8gpu.json
