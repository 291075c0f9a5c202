python -m pytest tests -m gpu -x -q 2>&1 | tail -3
q() { python bench.py --workload ${2:-cfg5} --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['roofline']['frac'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])"; }
q cfg2 cfg2
q cfg3 cfg3
q cfg5 cfg5
python tools/bench_kernels.py loop4096 cfg4 single 2>&1 | cut -c1-600
