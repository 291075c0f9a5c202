# usage: bash tools/run_multi.sh N <tag>   (under gpurun --gpus N): the in-process multi-GPU test and the torchrun bench
# (strong scaling by default; the weak-scaling, clean-map, e2e, e2e_multi and e2e_variants legs are in the same line)
N=${1:-2}; tag=${2:-vx}; O=gpurun_out/r2_${tag}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12 > ${O}_topo_n$N.txt; nproc >> ${O}_topo_n$N.txt
timeout 300 python -m pytest tests -m gpu -x -q -s -k "multi or shard" 2>&1 | tail -4 | tee ${O}_multi_gpu_test_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 3 > ${O}_bench_cfg5_n$N.json 2> ${O}_bench_cfg5_n$N.err
cut -c1-400 ${O}_bench_cfg5_n$N.json; tail -3 ${O}_bench_cfg5_n$N.err
