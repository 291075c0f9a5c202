# usage (under gpurun, one GPU, after `python tools/ab_build.py check="-DDPOSE_TILE_CHECK=1"`): the parity suites, the small-size sweep over every
# kernel and every bench workload (+ one row-band shard) on the bounds-checking build: any access outside the coefficient
# image, the tile plane or the row masks traps
L=dpose_b200/lib/libdpose_b200.so
cp dpose_b200/lib/ab/check.so $L; touch $L
(timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_solver.py tests/test_gpu_vs_reference_build.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/exercise_all_kernels.py 2>&1 | tail -1
for w in cfg5 cfg2 cfg3 cfg4; do timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-120; done
timeout 300 python bench.py --workload cfg5 --no-cpu-baseline --no-e2e --as-rank 5/8 2>/dev/null | cut -c1-120) | tee gpurun_out/r2_tile_check_build_tests.log
cp dpose_b200/lib/ab/default.so $L; touch $L
