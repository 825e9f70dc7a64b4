timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
PROF=1 DIAG_B=1024 python tools/prof_step.py 2>&1 | grep "k_tridiag_blk<5"
