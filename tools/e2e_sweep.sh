#!/bin/bash
# e2e chunk-size sweep of bench.py (diagnostics)
run() {
  MOP_BENCH_E2E_SPLIT=$1 MOP_BENCH_E2E_STREAMS=$2 MOP_BENCH_E2E_PRIO=$3 python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('split $1 streams $2 prio $3: value %.4g e2e %.4g ok %s resident %.4g' % (d['value'], d['e2e']['value'], d['e2e']['matches_resident_path'], d['e2e_hessian_resident']['value']))"
}
run 136,296,296,296 4 1
run 256,256,256,256 4 1
run 128,128,128,128,128,128,128,128 8 1
run 64,192,256,256,192,64 6 1
run 148,148,148,148,148,148,136 7 1
