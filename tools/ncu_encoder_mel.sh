# ncu --set full of the log-mel kernel(s), the encoder GEMM shapes, the attention kernel and LayerNorm in ONE profiled run.
# usage (GPU box, repo root): bash tools/ncu_encoder_mel.sh <tag> [count]
tag=${1:-r2}; cnt=${2:-14}
o=gpurun_out; mkdir -p $o
cmd="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --sample-len 4"
timeout 600 $cmd > $o/plain_$tag.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:logmel|gemm_tc|attention_tc|layernorm|mel_transpose|w2v|groupnorm|posconv' -c $cnt -f -o $o/${tag}_enc_mel $cmd > $o/ncu_$tag.log 2>&1
echo "ncu rc=$?"; tail -5 $o/ncu_$tag.log
