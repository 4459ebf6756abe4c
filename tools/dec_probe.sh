# decode-step phase probe: WXB_DEC_SKIP bitmask (1 cross-attn, 2 GEMV, 4 self-attn, 8 LayerNorm); results are timing only
for s in ${PROBE_MASKS:-0 15 14 13 11 7}; do
  WXB_DEC_SKIP=$s timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sample-len ${PROBE_LEN:-40} --batch-size ${PROBE_B:-60} --minutes ${PROBE_MIN:-30} 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('skip=$s ms/step %.3f  decode ms %.1f'%(r['ms_per_launch'], r['stages']['decode_steps']['ms']))"
done
