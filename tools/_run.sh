python tools/prof_lindh_rows.py 2>&1 | grep -v Warn | tail -8
