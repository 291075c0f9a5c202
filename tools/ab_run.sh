# usage: bash tools/ab_run.sh <tag> [workloads...]   (under gpurun, one GPU): every variant of dpose_b200/lib/ab/*.so (tools/ab_build.py)
# through bench.py, one line per (variant, workload) in gpurun_out/ab_<tag>.txt; the default library is put back at the end
tag=${1:-x}; shift; W=${@:-cfg5 cfg2 cfg3}
O=gpurun_out/ab_${tag}.txt; mkdir -p gpurun_out; : > $O
L=dpose_b200/lib/libdpose_b200.so
for so in dpose_b200/lib/ab/*.so; do
  v=$(basename $so .so); cp $so $L; touch $L
  for w in $W; do
    timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('$v', '$w', '%.4e' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'ms %.4f' % d['ms_per_step'], 'parity', d['parity']['pass_rate_vs_shifted_reference'], 'clean', (d.get('clean_map') or {}).get('value'))
" | tee -a $O
  done
done
cp dpose_b200/lib/ab/default.so $L; touch $L
