show() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); t=d['trials']; print(sys.argv[1], '%.0f trials/s' % t['value'], t['stage_ms'], 'pop %.0f' % t['population']['value'])" "$1"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for w in c1 c2 c5; do
  python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --trials-workload $w 2>/dev/null | show "rcp $w"
done
python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu 2>/dev/null | show "rcp c4"
python tools/opt_sweep.py 2>&1 | tail -4
