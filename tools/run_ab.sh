python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; cat gpurun_out/bench_v5.json
q() { python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['roofline']['frac'])"; }
for t in 512 1024; do DPOSE_TILE_THREADS=$t q "threads$t"; done
for r in 1 8; do DPOSE_TILE_REPLICAS=$r q "rep$r"; done
DPOSE_TILE_REPLICAS=8 DPOSE_TILE_THREADS=512 q "rep8_thr512"
DPOSE_NVCC_EXTRA=-DDPOSE_TILE_CVT=0 python -m dpose_b200.build --force >/dev/null 2>&1
q "cvt0_768"
DPOSE_TILE_THREADS=1024 q "cvt0_1024"
