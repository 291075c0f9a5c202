# usage: bash tools/run_round2.sh <tag>   (under gpurun, one GPU): smoke, GPU tests, every bench workload + reference arms,
# side kernels, ncu launch lists and full captures of the dominant kernels.  Everything lands in gpurun_out/r2_<tag>_*.
tag=${1:-vx}; O=gpurun_out/r2_${tag}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee ${O}_gpu_tests.log
for w in cfg5 cfg2 cfg3 cfg4 cfg1; do
  timeout 600 python bench.py --workload $w > ${O}_bench_$w.json 2>> ${O}_bench.err; cut -c1-160 ${O}_bench_$w.json
done
for w in cfg5 cfg4; do
  timeout 600 python bench.py --impl reference --workload $w --steps 3 --warmup 1 > ${O}_bench_${w}_reference_arm.json 2>> ${O}_bench.err
done
timeout 900 python tools/bench_kernels.py > ${O}_side_kernels.jsonl 2>> ${O}_bench.err
timeout 600 python tools/bench_solver.py > ${O}_solver.jsonl 2>> ${O}_bench.err
# launch lists (a number printed under ncu is never a bench value)
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
for w in cfg5 cfg2 cfg4; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_$w.csv $B --workload $w > /dev/null 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_lethal.csv python tools/lethal_only.py > /dev/null 2>&1
# full captures of the dominant kernels; summarised on the box (the .ncu-rep files are too big to travel back together)
cap() {  # cap <name> <kernel regex> <skip> <count> <command...>
  name=$1; pat=$2; skip=$3; cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k "regex:$pat" -s $skip -c $cnt -f -o ${O}_$name "$@" > /dev/null 2>&1
  python tools/ncu_metrics.py ${O}_$name.ncu-rep > ${O}_${name}_metrics.txt 2>&1
  ncu -i ${O}_$name.ncu-rep --page source --csv --print-source cuda,sass > ${O}_$name.src.csv 2>/dev/null
  python tools/ncu_lines.py ${O}_$name.src.csv 45 > ${O}_${name}_lines.txt 2>&1
  rm -f ${O}_$name.src.csv
  case $name in tile_*)  # the cell walk, instruction by instruction (executions, active threads, stall samples and reasons)
    ncu -i ${O}_$name.ncu-rep --page source --csv --print-source sass > ${O}_$name.sass.csv 2>/dev/null
    python tools/ncu_sass_loop.py ${O}_$name.sass.csv DADD.RM 14 26 > ${O}_${name}_sass_loop.txt 2>&1
    rm -f ${O}_$name.sass.csv;;
  esac
  if [ "$name" != "tile_cfg5" ]; then rm -f ${O}_$name.ncu-rep; fi
}
cap tile_cfg5 k_cost_batch_tile 3 1 $B --workload cfg5
cap tile_cfg2 k_cost_batch_tile 3 1 $B --workload cfg2
cap tile_cfg3 k_cost_batch_tile 3 1 $B --workload cfg3
cap solve_cfg4 k_solve 6 1 $B --workload cfg4
cap coop_4096 k_cost_batch_coop 20 1 python tools/bench_solver.py batch
cap lethal "k_pack_tiles|k_count|k_scan|k_write" 8 4 python tools/lethal_only.py
du -sh gpurun_out; ls gpurun_out | head -60
