timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_swart.py tests/test_crsirfo.py tests/test_neb.py tests/test_modelhess_d3.py tests/test_bias2.py -m gpu -x -q 2>&1 | tail -8
echo "rc=$?"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_pipeline or packed or eigh or c2 or step" 2>&1 | tail -6
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_eigh_large.py -m gpu -x -q 2>&1 | tail -6
