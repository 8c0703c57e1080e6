set -x
./tools/microbench > gpurun_out/microbench.txt 2>&1; tail -3 gpurun_out/microbench.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -c 600 gpurun_out/bench3.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench3_ref.json 2> gpurun_out/bench3_ref.err; tail -c 800 gpurun_out/bench3_ref.json
nproc; lscpu | head -20
