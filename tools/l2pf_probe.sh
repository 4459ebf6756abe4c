# A/B of the decoder's L2 staging plan inside ONE gpurun call: "mask:xa" pairs, per-phase profile + step time.
# usage: bash tools/l2pf_probe.sh "0:24 1:24 7:24" [sample_len]
for cfg in $1; do
  m=${cfg%%:*}; xa=${cfg##*:}
  WXB_DEC_L2PF=$m WXB_DEC_L2XA=$xa WXB_DEC_PROF=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sample-len ${2:-224} 2> gpurun_out/l2pf.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('l2pf=$m xa=$xa ms/step %.3f frac %.3f value %.1f' % (r['ms_per_step'], r['frac'], d['value']))"
  grep "wxb dec prof" gpurun_out/l2pf.err | tail -1 | cut -c60-400
done
