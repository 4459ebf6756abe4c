# parity tests + default bench + the tcgen05 encoder-attention variant
o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -x -q > $o/pytest_q.log 2>&1; echo "pytest rc=$?"; tail -3 $o/pytest_q.log
timeout 600 python bench.py --no-cpu-baseline > $o/bench_q.json 2> $o/bench_q.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value %.1f e2e %.1f ms/step %.1f | dec ms/step %.3f frac %.3f | enc ms %.1f frac %.3f | mel ms %.3f | ckv %.1f ctc %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['ms_per_step'], r['frac'], r['stages']['encoder']['ms'], r['stages']['encoder']['frac_bf16_burst'], r['stages']['mel']['ms'], r['stages']['cross_kv_gemm']['ms'], r['stages']['ctc']['ms']))
PY
WXB_ATTN=tc timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 2 > $o/bench_q_tc.json 2> $o/bench_q_tc.err; echo "bench tc rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q_tc.json').read().strip().splitlines()[-1]); r=d['roofline']
print('[WXB_ATTN=tc] value %.1f | enc ms %.1f frac %.3f' % (d['value'], r['stages']['encoder']['ms'], r['stages']['encoder']['frac_bf16_burst']))
PY
