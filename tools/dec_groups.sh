# decoder parity + per-phase timing for 1 / 2 / 3 co-resident sequence groups (WXB_DEC_GROUPS)
o=gpurun_out; mkdir -p $o
for g in ${GROUPS_LIST:-1 2 3}; do
  echo "== groups=$g parity"
  WXB_DEC_GROUPS=$g timeout 600 python -m pytest tests/test_gpu_decoder.py -x -q 2>&1 | tail -3
done
for g in ${GROUPS_LIST:-1 2 3}; do
  echo "== groups=$g bench (sample-len 40)"
  WXB_DEC_GROUPS=$g WXB_DEC_PROF=${PROF:-1} timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sample-len 40 --batch-size ${PROBE_B:-60} 2> $o/dec_g$g.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('groups=$g ms/step %.3f frac %.3f decode ms %.1f'%(r['ms_per_launch'], r['frac'], r['stages']['decode_steps']['ms']))"
  grep "wxb dec prof" $o/dec_g$g.err | tail -3
  grep -v "wxb dec prof" $o/dec_g$g.err | tail -5
done
