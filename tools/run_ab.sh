# usage (under gpurun, one GPU): bash tools/run_ab.sh "<nvcc extra flags for variant B>" [workload]
# A/B of two builds of the library on ONE box: the tree as it is, then rebuilt with the extra flags
# (e.g. "-DDPOSE_TILE_COMPACT=0", "-DDPOSE_TILE_WALK=0", "-DDPOSE_TILE_CVT=0", "-DDPOSE_TILE_CHECK=1").
extra=${1:-}; w=${2:-cfg5}
q() { python bench.py --workload $w --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', '%.4e poses/s' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'parity', d['parity']['pass_rate_vs_shifted_reference'])"; }
q "A (tree as is)"
if [ -n "$extra" ]; then
  DPOSE_NVCC_EXTRA="$extra" python -m dpose_b200.build --force >/dev/null 2>&1
  q "B ($extra)"
fi
