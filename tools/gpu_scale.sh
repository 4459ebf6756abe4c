# multi-GPU bench of ONE sharded job (strong scaling) under torchrun, as the driver launches it.  usage: bash tools/gpu_scale.sh <N> <tag>
N=${1:-2}; tag=${2:-scale}
o=gpurun_out; mkdir -p $o
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 2 --warmup 2 \
  > $o/bench_${tag}_n$N.json 2> $o/bench_${tag}_n$N.err; echo "bench n=$N rc=$?"
tail -c 1800 $o/bench_${tag}_n$N.json; tail -3 $o/bench_${tag}_n$N.err
