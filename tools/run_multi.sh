# usage: bash tools/run_multi.sh N <tag>   (under gpurun --gpus N)
N=${1:-2}; tag=${2:-vx}
nvidia-smi topo -m 2>/dev/null | head -12
python -m pytest tests -m gpu -x -q -k "multi or shard" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err
cat gpurun_out/bench_${tag}_n$N.json; tail -3 gpurun_out/bench_${tag}_n$N.err
