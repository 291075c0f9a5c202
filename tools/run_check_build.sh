L=dpose_b200/lib/libdpose_b200.so
cp dpose_b200/lib/ab/check.so $L; touch $L
(timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_solver.py tests/test_gpu_vs_reference_build.py -m gpu -q 2>&1 | tail -3
for w in cfg5 cfg2 cfg3 cfg4; do timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e 2>&1 | cut -c1-120; done) | tee gpurun_out/r2_v26_tile_check_build_tests.log
cp dpose_b200/lib/ab/default.so $L; touch $L
