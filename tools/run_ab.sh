DPOSE_NVCC_EXTRA=-DDPOSE_TILE_CHECK=1 python -m dpose_b200.build --force >/dev/null 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in cfg5 cfg3 cfg2; do python bench.py --workload $w --no-cpu-baseline --steps 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w checked build', d['value'], d['parity'])"; done
