# A/B of library builds inside ONE gpurun call (boxes differ): tools/probe/libs/*.so, alternating, PROBE_LEN positions
for rep in 1 2; do
for lib in tools/probe/libs/*.so; do
  WXB200_LIB=$PWD/$lib WXB_DEC_PROF=1 timeout 300 python bench.py --allow-env --no-align --no-extras --batch-size ${PROBE_B:-60} --minutes ${PROBE_MIN:-30} --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sample-len ${PROBE_LEN:-224} 2> gpurun_out/ab.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$lib ms/step %.3f frac %.3f' % (r['ms_per_step'], r['frac']))"
  grep "wxb dec prof" gpurun_out/ab.err | tail -1 | cut -c60-330
done; done
