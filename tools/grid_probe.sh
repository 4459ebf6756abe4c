# per-phase times of the decode kernel on a narrower grid (-DWXB_PROBE build): what one sequence group costs on half the SMs
lib=$PWD/tools/probe/libs/probe.so
run() { # batch minutes grid
  echo "== batch $1 grid $3"
  WXB200_LIB=$lib WXB_DEC_GRID=$3 WXB_DEC_PROF=1 timeout 300 python bench.py --allow-env --no-align --no-extras --batch-size $1 --minutes $2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sample-len ${PROBE_LEN:-224} 2> gpurun_out/gp.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms/step %.3f frac %.3f' % (r['ms_per_step'], r['frac']))"
  grep "wxb dec prof" gpurun_out/gp.err | tail -1 | cut -c60-330
}
run 60 30 148
run 30 15 148
run 30 15 74
run 30 15 100
run 20 10 50
