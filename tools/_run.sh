timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['matches_resident_path'])
for k,v in d['per_config'].items(): print(k, v.get('value'), v.get('ms_per_iteration', v.get('ms_per_step')), v.get('parity_vs_oracle'))"
