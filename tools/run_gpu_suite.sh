# usage: bash tools/run_gpu_suite.sh <tag>   (under gpurun) — GPU tests, default bench, other workloads
tag=${1:-vx}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json
for w in cfg2 cfg3; do python bench.py --workload $w > gpurun_out/bench_${tag}_$w.json 2>> gpurun_out/bench_$tag.err; cat gpurun_out/bench_${tag}_$w.json; done
tail -5 gpurun_out/bench_$tag.err
