timeout 600 python -m pytest tests/test_crsirfo.py tests/test_bias2.py tests/test_neb.py -m gpu -x -q 2>&1 | tail -25
