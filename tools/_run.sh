timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_pipeline or packed" 2>&1 | tail -3
