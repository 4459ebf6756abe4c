"""Stand-alone timing of the encoder self-attention kernel (large-v3 shape: 20 heads, T = 1500)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "whisperx-mlx_b200"))
from whisperx._native import get_context
ctx = get_context(0)
B, T, H = int(os.environ.get("ATT_B", 60)), 1500, 20
d = 64 * H
qkv = (torch.randn(B * T, 3 * d, device="cuda") * 1.0).to(torch.bfloat16)
for _ in range(3):
    ctx.encoder_attention(qkv, B, T, H)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    ctx.encoder_attention(qkv, B, T, H)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 4.0 * B * H * T * T * 64
print(f"attention B={B}: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s (WXB_ATTN={os.environ.get('WXB_ATTN', 'default')})")
