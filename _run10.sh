python tools/opt_sweep.py c4 256 2>&1 | tail -14
python tools/opt_sweep.py c4 1024 2>&1 | grep "outer  32 newton 2"
