timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/population_sweep.py c5 1024 4096 16384 65536 2>&1 | tail -4
