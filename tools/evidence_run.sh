#!/bin/bash
# Round-2 evidence on one B200 (run under gpurun, one part per call: gpurun returns at most 64 MiB):
#   part 1: GPU tests, smoke, the bench line (with per_config records), the ncu launch list of the bench command, the
#           reference arm;  part 2: ncu --set full of the C2 step kernels and of the C5 cluster tridiagonalisation;
#   part 3: ncu --set full of the producers.  Every ncu run follows a plain run of the same command that exited 0.
P=r2
part=${1:-1}
if [ "$part" = 1 ]; then
  timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
  python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' 2>&1 | tail -1
  python bench.py > gpurun_out/${P}_bench_c2_1gpu.json 2> gpurun_out/${P}_bench_err.log
  tail -c 300 gpurun_out/${P}_bench_c2_1gpu.json
  python bench.py --no-per-config --steps 2 --warmup 3 > gpurun_out/plain_bench.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches.csv \
        python bench.py --no-per-config --steps 2 --warmup 3 > gpurun_out/ncu_ll.log 2>&1
  python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${P}_bench_reference_arm.json
  cut -c1-300 gpurun_out/${P}_bench_reference_arm.json
elif [ "$part" = 2 ]; then
  DIAG_B=1024 python tools/prof_step.py > gpurun_out/plain_prof.log 2>&1 &&
    DIAG_B=1024 ncu --set full --clock-control none -k regex:"k_tridiag_blk|k_spectrum_step" -s 6 -c 6 \
        -o gpurun_out/${P}_full python tools/prof_step.py > gpurun_out/ncu_full.log 2>&1
  CL=0 python tools/run_cluster_once.py 600 256 > gpurun_out/plain_cluster.log 2>&1 &&
    CL=0 ncu --set full --clock-control none -k regex:"k_lg_tridiag_blk|k_lg_trieig" -s 2 -c 2 \
        -o gpurun_out/${P}_c5 python tools/run_cluster_once.py 600 256 > gpurun_out/ncu_c5.log 2>&1
  ls -la gpurun_out/*.ncu-rep
else
  python tools/measure_rows.py > gpurun_out/${P}_row_table.md 2> gpurun_out/rows_err.log
  python tools/prof_producers.py > gpurun_out/plain_producers.log 2>&1 &&
    ncu --set full --clock-control none \
        -k regex:"k_swart|k_lindh|k_ric|k_model_hessian|k_afir|k_bias_terms|k_project_trrot" -c 12 \
        -o gpurun_out/${P}_producers python tools/prof_producers.py > gpurun_out/ncu_producers.log 2>&1
  ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/plain_producers.log
fi
