timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py 2>/tmp/bench_err.log > gpurun_out/r2_bench_c2_1gpu.json; tail -2 /tmp/bench_err.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c2_1gpu.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'res',d['e2e_hessian_resident']['value'],'frac',d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'launches', d['gpu_launches'])
for k,v in d['per_config'].items(): print(k, v.get('value'), v.get('ms_per_iteration', v.get('ms_per_step')), v.get('parity_vs_oracle'))
print(d['cpu_baseline']['value'])
PY
