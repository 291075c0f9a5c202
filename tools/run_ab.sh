q() { python bench.py --workload ${2:-cfg5} --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['roofline']['frac'], d['parity']['pass_rate_vs_shifted_reference'])"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
q "fx1 cfg5"
q "fx1 cfg3" cfg3
DPOSE_NVCC_EXTRA=-DDPOSE_TILE_FX=0 python -m dpose_b200.build --force >/dev/null 2>&1
q "fx0 cfg5"
q "fx0 cfg3" cfg3
