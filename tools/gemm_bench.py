"""Stand-alone timing of the encoder GEMM shapes (large-v3, 60 chunks: M = 90000)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "whisperx-mlx_b200"))
from whisperx._native import get_context
ctx = get_context(0)
M, d = 90000, 1280
for name, N, K, gelu, f32 in (("qkv", 3 * d, d, False, False), ("out", d, d, False, True), ("fc1", 4 * d, d, True, False), ("fc2", d, 4 * d, False, True)):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    for _ in range(3):
        ctx.gemm_bf16(A, W, b, gelu=gelu, out_f32=f32)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        ctx.gemm_bf16(A, W, b, gelu=gelu, out_f32=f32)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name}: N={N} K={K} gelu={gelu} f32={f32}: {ms * 1e3:.0f} us, {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s")
