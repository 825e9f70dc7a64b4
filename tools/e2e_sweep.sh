#!/bin/bash
# e2e chunk-size sweep of bench.py (diagnostics): MOP_BENCH_E2E_SPLIT / _STREAMS feed HostStepPipeline(chunks, nstream)
run() {
  MOP_BENCH_E2E_SPLIT=$1 MOP_BENCH_E2E_STREAMS=$2 python bench.py --no-per-config --steps 5 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('split $1 streams $2: value %.4g e2e %.4g ok %s resident %.4g' % (d['value'], d['e2e']['value'], d['e2e']['matches_resident_path'], d['e2e_hessian_resident']['value']))"
}
run 136,296,296,296 4
run 136,296,296,148,148 5
run 136,296,296,222,74 5
