# usage: bash tools/ab_batch.sh <tag>  (under gpurun): device time per batched launch (tools/bench_solver.py batch) for every variant in
# dpose_b200/lib/ab/*.so -> gpurun_out/ab_<tag>_batch.txt
tag=${1:-x}; L=dpose_b200/lib/libdpose_b200.so
for so in dpose_b200/lib/ab/*.so; do v=$(basename $so .so); cp $so $L; touch $L; echo "== $v"; python tools/bench_solver.py batch 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln)
    print(d['poses'], 'volatile' if d['volatile_map'] else 'clean', round(d['us_per_call'], 2))
"; done | tee gpurun_out/ab_${tag}_batch.txt
cp dpose_b200/lib/ab/default.so $L; touch $L
