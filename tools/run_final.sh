# usage: bash tools/run_final.sh <tag>  (under gpurun, one GPU): smoke, tests, both bench arms, all workloads, side kernels
tag=${1:-final}
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_${tag}_ref.json
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; cut -c1-400 gpurun_out/bench_${tag}.json
for w in cfg2 cfg3; do python bench.py --workload $w > gpurun_out/bench_${tag}_$w.json 2>> gpurun_out/bench_${tag}.err; done
python tools/bench_kernels.py > gpurun_out/kernels_${tag}.jsonl 2>&1; tail -3 gpurun_out/kernels_${tag}.jsonl | cut -c1-300
