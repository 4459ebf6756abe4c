# per-phase timing of the persistent decode kernel (WXB_DEC_PROF) for a list of WXB_DEC_SKIP masks
for s in ${PROBE_MASKS:-0}; do
  echo "skip=$s"
  WXB_DEC_SKIP=$s WXB_DEC_PROF=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sample-len ${PROBE_LEN:-40} --batch-size ${PROBE_B:-60} --minutes ${PROBE_MIN:-30} 2>&1 >/dev/null | grep "wxb dec prof" | tail -3
done
