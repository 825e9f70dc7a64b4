#!/bin/bash
# Round evidence on one B200 (run under gpurun): GPU tests, the bench line, the ncu launch list of the bench
# command, one ncu --set full capture of the dominant kernels, and the config-5 bench line.
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -1
python bench.py > gpurun_out/r1b_bench_c2_1gpu.json 2> gpurun_out/r1b_bench_err.log
tail -c 300 gpurun_out/r1b_bench_c2_1gpu.json
python bench.py --steps 2 --warmup 1 > gpurun_out/plain_bench.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches.csv \
      python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_ll.log 2>&1
DIAG_B=1024 python tools/prof_step.py > gpurun_out/plain_prof.log 2>&1 &&
  DIAG_B=1024 ncu --set full --clock-control none --import-source on \
      -k regex:"k_tridiag_rwf|k_spectrum_step|k_project_trrot|k_upd_apply|k_upd_matvec" -s 4 -c 5 \
      -o gpurun_out/r1b_full python tools/prof_step.py > gpurun_out/ncu_full.log 2>&1
python bench.py --workload c5 --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r1b_bench_c5_1gpu.json
cut -c1-300 gpurun_out/r1b_bench_c5_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r1b_bench_reference_arm.json
cut -c1-300 gpurun_out/r1b_bench_reference_arm.json
