python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline 2>/dev/null | cut -c1-900
python tools/bench_kernels.py lethal single loop4096 2>&1 | tee gpurun_out/kernels_v8.jsonl
