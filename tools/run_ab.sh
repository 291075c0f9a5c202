q() { python bench.py --workload ${2:-cfg5} --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['roofline']['frac'], d['parity']['pass_rate_vs_shifted_reference'])"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
q cfg3 cfg3
DPOSE_TILE_THREADS=768 q "cfg3 768" cfg3
q cfg5 cfg5
