timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-per-config 2>/tmp/bench_err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step']); print('e2e',json.dumps(d['e2e'])[:700]); print(d['e2e_hessian_resident'])"
tail -3 /tmp/bench_err.log
