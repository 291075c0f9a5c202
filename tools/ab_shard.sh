# usage: bash tools/ab_shard.sh <tag>  (under gpurun): tools/bench_solver.py shard for every variant in dpose_b200/lib/ab/*.so
tag=${1:-x}; L=dpose_b200/lib/libdpose_b200.so
for so in dpose_b200/lib/ab/*.so; do v=$(basename $so .so); cp $so $L; touch $L; echo "== $v"; python tools/bench_solver.py shard 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print(d['poses'], 'volatile' if d['volatile_map'] else 'clean', round(d['us_per_step'], 1))
"; done | tee gpurun_out/ab_${tag}_shard.txt
cp dpose_b200/lib/ab/default.so $L; touch $L
