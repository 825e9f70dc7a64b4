timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/trieig_phases.py 600 256
python tools/prof_eigh.py 600 256 2>&1 | grep "k_lg_trieig"
python tools/prof_eigh.py 150 1024 2>&1 | grep "k_lg_trieig"
