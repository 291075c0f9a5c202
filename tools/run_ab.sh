python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -12
python tools/bench_kernels.py cfg4 2>&1 | cut -c1-400
