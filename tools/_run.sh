timeout 600 python -m pytest tests/test_eigh_large.py tests/test_rsprfo.py -m gpu -x -q 2>&1 | tail -2
python tools/trieig_phases.py 600 256
