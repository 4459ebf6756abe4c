set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
for g in 1 2; do for s in 0 1 2 4 6; do
  WXB_DEC_GROUPS=$g WXB_DEC_SKIP=$s timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sample-len 40 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('groups=$g skip=$s ms/step %.3f  decode ms %.1f'%(r['ms_per_launch'], r['stages']['decode_steps']['ms']))"
done; done
