python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/bench_kernels.py lethal 2>&1 | cut -c1-420
