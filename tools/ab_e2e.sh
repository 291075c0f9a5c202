# usage: bash tools/ab_e2e.sh <tag>  (under gpurun): the end-to-end leg of bench.py for every variant in dpose_b200/lib/ab/*.so
tag=${1:-x}; L=dpose_b200/lib/libdpose_b200.so; mkdir -p gpurun_out
for so in dpose_b200/lib/ab/*.so; do v=$(basename $so .so); cp $so $L; touch $L
  python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); e = d['e2e']; print('$v', 'e2e ms', round(e['ms_per_step'], 3), 'floor', round(e['pcie_bidirectional_ms_per_step'], 3), 'cost_only', '%.3e' % d['e2e_variants']['cost_only']['value'], 'sweep', '%.3e' % d['e2e_variants']['sweep']['value'], 'value ms', round(d['ms_per_step'], 4))
" | tee -a gpurun_out/ab_${tag}_e2e.txt
done
cp dpose_b200/lib/ab/default.so $L; touch $L
