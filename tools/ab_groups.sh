# A/B of the decoder's sequence-group schedule inside ONE gpurun call: the in-tree library vs tools/probe/libs/*.so, group counts
# and start offsets.  usage: bash tools/ab_groups.sh  (PROBE_B rows, PROBE_LEN positions, DELAYS list)
run() {  # lib label extra-args
  WXB200_LIB=$1 WXB_DEC_PROF=1 timeout 300 python bench.py --allow-env --no-align --no-extras --batch-size ${PROBE_B:-60} --minutes ${PROBE_MIN:-30} \
    --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --sample-len ${PROBE_LEN:-224} ${@:3} 2> gpurun_out/abg.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$2 ms/step %.3f frac %.3f' % (r['ms_per_step'], r['frac']))"
  grep "wxb dec prof" gpurun_out/abg.err | tail -${PROF_LINES:-2} | cut -c15-400
  grep -i "error\|fail" gpurun_out/abg.err | head -3
}
base=$PWD/whisperx-mlx_b200/lib/libwxb200.so
[ -z "$NO_BASE" ] && run $base "base"
for lib in ${LIBS:-tools/probe/libs/*.so}; do
  L=$PWD/$lib
  for dl in ${DELAYS:-0 110000}; do
    run $L "$(basename $lib) G=${GROUPS_N:-2} dl=$dl" --dec-groups ${GROUPS_N:-2} --dec-group-delay-ns $dl
  done
done
