q() { python bench.py --workload ${2:-cfg5} --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['roofline']['frac'], d['parity']['pass_rate_vs_shifted_reference'])"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
q "lds+lds (rep4)"
DPOSE_NVCC_EXTRA=-DDPOSE_TILE_TEX=1 python -m dpose_b200.build --force >/dev/null 2>&1
q "lds(rep8)+tex"
DPOSE_TILE_REPLICAS=4 q "lds(rep4)+tex"
q "lds+tex cfg3" cfg3
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
