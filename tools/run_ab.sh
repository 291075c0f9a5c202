python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_kernels.py lethal 2>&1 | tee gpurun_out/kernels_v15.jsonl | cut -c1-420
