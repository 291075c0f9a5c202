# usage: bash tools/run_ncu.sh <tag> [kernel-regex]   (under gpurun; one GPU)
tag=${1:-vx}; pat=${2:-k_cost_batch_tile}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c 1 -f -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_f_$tag.log 2>&1
tail -2 gpurun_out/plain_$tag.log; tail -3 gpurun_out/ncu_f_$tag.log
