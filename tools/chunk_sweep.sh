#!/bin/bash
for c in 0 128 192 256 384 512; do
  MOP_STREAM_CHUNK=$c python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunk $c: value %.4g ms %.3f e2e %.4g parity %.2e' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['parity_vs_oracle']))"
done
