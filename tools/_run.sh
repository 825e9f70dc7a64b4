timeout 600 python -m pytest tests/test_eigh_large.py tests/test_rsprfo.py -m gpu -x -q 2>&1 | tail -2
python tools/trieig_phases.py 600 256
python tools/prof_eigh.py 600 256 2>&1 | grep "mop::" | head -5
MOP_TRIEIG_1024=1 python tools/prof_eigh.py 600 256 2>&1 | grep "k_lg_trieig"
python tools/prof_eigh.py 600 32 2>&1 | grep "k_lg_trieig"
MOP_TRIEIG_1024=1 python tools/prof_eigh.py 600 32 2>&1 | grep "k_lg_trieig"
