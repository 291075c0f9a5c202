python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_kernels.py cfg4 loop4096 2>&1 | tee gpurun_out/kernels_v9.jsonl
python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-200
