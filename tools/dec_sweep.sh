# timing sweep of the decode step: "groups:prefetch[:skipmask]" tuples in SWEEP
o=gpurun_out; mkdir -p $o
for cfg in ${SWEEP:-1:0 2:0 2:2 2:4 3:0 3:2}; do
  g=${cfg%%:*}; rest=${cfg#*:}; pf=${rest%%:*}; sk=0; [ "$rest" != "$pf" ] && sk=${rest#*:}
  WXB_DEC_GROUPS=$g WXB_XA_PF=$pf WXB_DEC_SKIP=$sk WXB_DEC_PROF=${PROF:-1} timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sample-len ${PROBE_LEN:-40} --batch-size ${PROBE_B:-60} 2> $o/dec_sweep.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('groups=$g pf=$pf skip=$sk ms/step %.3f frac %.3f decode ms %.1f'%(r['ms_per_launch'], r['frac'], r['stages']['decode_steps']['ms']))"
  grep "wxb dec prof" $o/dec_sweep.err | tail -3
  grep -v "wxb dec prof" $o/dec_sweep.err | tail -3
done
