timeout 600 python -m pytest tests/test_edge_new.py -m gpu -x -q 2>&1 | tail -15
