# per-lib, per-config phase profile: "skip:l2pf:l2xa" triples
for lib in tools/probe/libs/*.so; do
for cfg in $1; do
  IFS=: read sk m xa <<< "$cfg"
  echo "$lib skip=$sk l2pf=$m xa=$xa"
  WXB200_LIB=$PWD/$lib WXB_DEC_SKIP=$sk WXB_DEC_L2PF=$m WXB_DEC_L2XA=$xa WXB_DEC_PROF=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sample-len ${2:-40} 2>&1 >/dev/null | grep "wxb dec prof" | tail -1 | cut -c60-400
done; done
