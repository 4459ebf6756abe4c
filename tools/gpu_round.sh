# One GPU visit: parity tests, smoke, bench (both arms), ncu launch list, ncu --set full of the decode step kernel.
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag>
tag=${1:-head}
o=gpurun_out
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $o/gpu_$tag.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > $o/pytest_$tag.log 2>&1; echo "pytest rc=$?"
timeout 300 python __graft_entry__.py smoke > $o/smoke_$tag.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; echo "bench rc=$?"
tail -c 3000 $o/bench_$tag.json
if [ -z "$SKIP_REF" ]; then timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $o/bench_ref_$tag.json 2>/dev/null; echo "ref rc=$?"; fi
if [ -z "$SKIP_NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $o/launches_r2_$tag.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras --sample-len 20 > $o/ncu_launch_$tag.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dec_step_kernel -s ${NCU_SKIP:-3} -c 1 -f -o $o/r2_dec_step_$tag \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-extras --no-align --sample-len 20 > $o/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
fi
