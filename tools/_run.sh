timeout 600 python -m pytest tests/test_eigh_large.py tests/test_rsprfo.py -m gpu -x -q 2>&1 | tail -2
python tools/time_cluster.py 600 256 2>&1 | grep "ms per mop_eigh"
python tools/time_cluster.py 600 32 2>&1 | grep "ms per mop_eigh"
