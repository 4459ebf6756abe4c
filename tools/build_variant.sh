# Build a variant of libwxb200.so with extra nvcc defines into tools/probe/libs/<name>.so (for tools/ab.sh).
# usage: bash tools/build_variant.sh <name> [-DWXB_XA_NS=1 ...]
set -e
name=$1; shift
src=whisperx-mlx_b200/csrc; out=tools/probe/libs; obj=$out/obj_$name
mkdir -p $obj
for f in wxb_api wxb_ctc wxb_logmel wxb_gemm wxb_attn wxb_encoder wxb_decoder wxb_model wxb_w2v wxb_dtw wxb_vad; do
  if [ "$f" = "wxb_decoder" ] || [ -n "$ALL" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr "$@" -c $src/$f.cu -o $obj/$f.o 2> $obj/$f.log
    grep -A2 "dec_step_kernelILi4" $obj/$f.log | tail -2 || true
  else
    cp whisperx-mlx_b200/lib/obj/$f.o $obj/$f.o
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/$name.so $obj/*.o -cudart static
rm -rf $obj
