timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/population_sweep.py c2 2048 8192 2>&1 | tail -2
python tools/population_sweep.py c1 4096 2>&1 | tail -1
python tools/population_sweep.py c5 1024 2>&1 | tail -1
python tools/population_sweep.py c4 256 2>&1 | tail -1
