timeout 300 python -m pytest tests/test_swart.py -m gpu -x -q 2>&1 | tail -3
TIME=1 python tools/prof_producers.py 2>&1 | grep swart
MOP_SWP_MINB=2 TIME=1 python tools/prof_producers.py 2>&1 | grep swart
