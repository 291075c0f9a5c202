# usage: bash tools/ab_cfg4.sh <tag>  (under gpurun): bench.py --workload cfg4 for every variant in dpose_b200/lib/ab/*.so
tag=${1:-x}; L=dpose_b200/lib/libdpose_b200.so
for so in dpose_b200/lib/ab/*.so; do v=$(basename $so .so); cp $so $L; touch $L
  timeout 300 python bench.py --workload cfg4 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('$v', 'device us %.1f' % (1e3 * d['ms_per_step']), 'call us %.1f' % (1e3 * d['e2e']['ms_per_step']), 'evals/s %.3e' % d['value'], d.get('parity'))
"; done | tee gpurun_out/ab_${tag}_cfg4.txt
cp dpose_b200/lib/ab/default.so $L; touch $L
