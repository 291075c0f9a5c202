python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in cfg5 cfg3 cfg2; do python bench.py --workload $w --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w', d['value'], d['roofline']['frac'], d['parity'])"; done
DPOSE_TILE_THREADS=768 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('thr768', d['value'], d['roofline']['frac'])"
DPOSE_TILE_REPLICAS=8 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rep8', d['value'], d['roofline']['frac'])"
