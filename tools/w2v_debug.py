"""Bring-up aid: every stage of the native wav2vec2 forward against torchaudio's intermediates (one short segment).
usage (GPU box): python tools/w2v_debug.py [seconds]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "whisperx-mlx_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torchaudio  # noqa: E402
from fake_ctc_model import synthetic_speech  # noqa: E402
from whisperx.align_model import Wav2Vec2B200, W2V_BASE_DIMS, kernel_layout_to_torchaudio, random_init_torchaudio  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
params = torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H._params
ref = random_init_torchaudio(params, 0)
ours = Wav2Vec2B200(ref.state_dict(), "cuda", W2V_BASE_DIMS)
torch.nn.utils.parametrize.remove_parametrizations(ref.encoder.transformer.pos_conv_embed.conv, "weight")
ref.load_state_dict(kernel_layout_to_torchaudio(ours.kernel_weights, ours.dims))
ref.eval()
wave = synthetic_speech(secs, seed=31)
S = len(wave)
P = max(2, -(-S // 320))
ctx = ours.ctx


def report(name, got, want):
    got, want = got.float().cpu(), want.float().cpu()
    err = (got - want).abs()
    print(f"{name:28s} shape {tuple(want.shape)} max-abs err {float(err.max()):.5f} mean {float(err.mean()):.6f} | ref std {float(want.std()):.4f} "
          f"| worst row {int(err.amax(dim=-1).argmax())}", flush=True)


with torch.inference_mode():
    x = torch.from_numpy(wave)[None, None]
    feats = []
    for layer in ref.feature_extractor.conv_layers:
        x, _ = layer(x, None)
        feats.append(x[0].t().contiguous())  # [T, 512]
    f = feats[-1][None]
    proj = ref.encoder.feature_projection(f)
    xpos = proj + ref.encoder.transformer.pos_conv_embed(proj)
    layers = []
    h = xpos
    for layer in ref.encoder.transformer.layers:
        h, _ = layer(h)
        layers.append(h[0])
    want_logits = ref.aux(ref.encoder.transformer.layer_norm(h))[0]

dev, offs, lens = ours.upload([wave])
for stage in range(0, 9 + 12):
    ctx.debug_set("w2v_stop", stage)
    try:
        ours.emissions_device(dev, offs, lens)
    finally:
        ctx.debug_set("w2v_stop", -1)
    if stage <= 6:
        T = feats[stage].shape[0]
        got = ctx.debug_buffer(f"w2v.c{stage}", (T, 512), torch.bfloat16)
        report(f"conv layer {stage}", got, feats[stage])
    elif stage == 7:
        T = proj.shape[1]
        report("feature projection", ctx.debug_buffer("w2v.x", (P, 768), torch.float32)[:T], proj[0])
    elif stage == 8:
        T = proj.shape[1]
        report("x + pos conv", ctx.debug_buffer("w2v.x", (P, 768), torch.float32)[:T], xpos[0])
    else:
        l = stage - 9
        T = proj.shape[1]
        if l == 0:
            qkv = ctx.debug_buffer("w2v.qkv", (P, 3 * 768), torch.bfloat16)[:T]
            a = ref.encoder.transformer.layers[0].attention
            with torch.inference_mode():
                wq = torch.cat([a.q_proj(xpos), a.k_proj(xpos), a.v_proj(xpos)], -1)[0]
            report("layer 0 qkv", qkv, wq)
            att = ctx.debug_buffer("w2v.att", (P, 768), torch.bfloat16)[:T]
            with torch.inference_mode():
                q, k, v = [t.view(1, T, 12, 64).transpose(1, 2) for t in (a.q_proj(xpos), a.k_proj(xpos), a.v_proj(xpos))]
                wa = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(T, 768)
            report("layer 0 attention", att, wa)
        report(f"transformer layer {l}", ctx.debug_buffer("w2v.x", (P, 768), torch.float32)[:T], layers[l])
emis, _ = ours.emissions_device(dev, offs, lens)
report("emission logits", emis, want_logits)
