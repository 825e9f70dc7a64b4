python tools/prof_configs.py c4 24 2>&1 | grep -v Warn | head -14
python tools/prof_configs.py sp 24 2>&1 | tail -10
