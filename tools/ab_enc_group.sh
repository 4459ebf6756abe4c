# A/B of the encoder's chunk-group size inside ONE gpurun call (results do not depend on it): encoder stage ms / TFLOP/s of the
# 60-chunk job for each group size.  usage: bash tools/ab_enc_group.sh "0 8 10 12 15 20 30"
for g in ${1:-0 8 10 12 15 20 30}; do
  timeout 300 python bench.py --no-align --no-extras --no-e2e --no-cpu-baseline --sample-len ${PROBE_LEN:-8} --steps 3 --warmup 2 --enc-group $g 2> gpurun_out/abe.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['roofline']['stages']['encoder']
print('enc_group $g: encoder %.2f ms  %.1f TFLOP/s  burst %.3f sustained %.3f' % (e['ms'], e['TFLOP/s'], e['frac_bf16_burst'], e['frac_bf16_sustained']))"
  grep -i "error\|fail" gpurun_out/abe.err | head -3
done
