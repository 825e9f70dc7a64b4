#!/bin/bash
# N-GPU bench line (run under gpurun --gpus N): the default line (config 2 weak scaling + per_config records of
# configs 3, 4, 5 on N ranks), as the driver launches it
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/r2_scale_${N}gpu.err | tail -1 > gpurun_out/r2_bench_c2_${N}gpu.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_c2_${N}gpu.json"))
print("N", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"])
for k, v in d.get("per_config", {}).items():
    print(" ", k, v.get("value"), v.get("unit"), v.get("ms_per_iteration", v.get("ms_per_step")), v.get("halo_exchange_ms"))
PY
