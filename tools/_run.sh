timeout 600 python -m pytest tests/test_eigh_large.py tests/test_rsprfo.py tests/test_full_size.py -m gpu -x -q 2>&1 | tail -3
python tools/time_cluster.py 600 256 2>&1 | grep "cluster=2 sym=1\|cluster=1"
python tools/time_cluster.py 600 32 2>&1 | grep "cluster=4 sym=0"
python tools/prof_configs.py c5 2>&1 | grep -v Warn | head -8
