python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_kernels.py precompute lethal > gpurun_out/kernels_v10.jsonl 2>&1 && cat gpurun_out/kernels_v10.jsonl | cut -c1-400 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_lethal.csv python tools/bench_kernels.py lethal > /dev/null 2>&1
grep -E "k_count|k_scan|k_write" gpurun_out/launches_lethal.csv | awk -F'","' '{print $5, $NF}' | head -24
