timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/opt_sweep.py c4 256 2>&1 | grep "outer  32 newton 2\|outer  32 newton 1"
python tools/population_sweep.py c4 256 4096 2>&1 | tail -2
