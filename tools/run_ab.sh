q() { python bench.py --workload ${2:-cfg5} --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['roofline']['frac'], d['parity']['pass_rate_vs_shifted_reference'])"; }
q "allslots 1024 rep4"
DPOSE_NVCC_EXTRA=-DDPOSE_TILE_COMPACT=1 python -m dpose_b200.build --force >/dev/null 2>&1
q "compact 1024 rep8"
DPOSE_TILE_REPLICAS=4 q "compact 1024 rep4"
q "compact cfg3" cfg3
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
