# usage: bash tools/cap_tile.sh <tag> [workload]   (under gpurun, one GPU): one ncu --set full capture of k_cost_batch_tile on a
# bench workload; summaries + the SASS-level source page land in gpurun_out/r2_<tag>_tile_<workload>_*
tag=${1:-vx}; w=${2:-cfg5}; O=gpurun_out/r2_${tag}_tile_$w
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --workload $w"
ncu --set full --clock-control none --import-source on -k "regex:k_cost_batch_tile" -s 3 -c 1 -f -o $O $B > /dev/null 2>&1
python tools/ncu_metrics.py $O.ncu-rep > ${O}_metrics.txt 2>&1
ncu -i $O.ncu-rep --page source --csv --print-source cuda,sass > $O.src.csv 2>/dev/null
python tools/ncu_lines.py $O.src.csv 45 > ${O}_lines.txt 2>&1
ncu -i $O.ncu-rep --page source --csv --print-source sass > ${O}_sass.csv 2>/dev/null
rm -f $O.src.csv $O.ncu-rep
head -40 ${O}_metrics.txt
