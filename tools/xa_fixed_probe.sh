# fixed cost of the cross-attention phase: per-phase times with the piece merge (32), the math (16) or the whole stream (1) left out
lib=$PWD/tools/probe/libs/probe.so
for cfg in "60 30" "8 4"; do set -- $cfg
for s in 0 32 48 1; do
  echo "== batch $1 skip=$s"
  WXB200_LIB=$lib WXB_DEC_SKIP=$s WXB_DEC_PROF=1 timeout 300 python bench.py --allow-env --no-align --no-extras --batch-size $1 --minutes $2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sample-len 224 2>&1 >/dev/null | grep "wxb dec prof" | tail -1 | cut -c60-330
done; done
