# Per-phase profile of the persistent decode kernel for a list of WXB_DEC_SKIP probe masks (results of a masked run are
# numerically meaningless; timings only).  usage: bash tools/dec_phase_probe.sh "0 64" [sample_len]
for s in ${1:-0}; do
  echo "skip=$s"
  WXB_DEC_SKIP=$s WXB_DEC_PROF=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sample-len ${2:-40} 2>&1 >/dev/null | grep "wxb dec prof" | tail -1 | cut -c60-400
done
