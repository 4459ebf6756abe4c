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

configs = sys.argv[1:] or ["3.0"]   # each argument: comma-separated segment durations of one batched call
params = torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H._params
ref = random_init_torchaudio(params, 0)
ours = Wav2Vec2B200(ref.state_dict(), "cuda", W2V_BASE_DIMS)
torch.nn.utils.parametrize.remove_parametrizations(ref.encoder.transformer.pos_conv_embed.conv, "weight")
ref.load_state_dict(kernel_layout_to_torchaudio(ours.kernel_weights, ours.dims))
ref.eval()
torch.set_num_threads(16)
ctx = ours.ctx


def report(name, got, want):
    got, want = got.float().cpu(), want.float().cpu()
    err = (got - want).abs()
    flag = "  <<<<<<" if float(err.max()) > 0.1 * max(float(want.std()), 1e-3) + 0.01 else ""
    print(f"  {name:26s} shape {tuple(want.shape)} max-abs err {float(err.max()):.5f} mean {float(err.mean()):.6f} | ref std {float(want.std()):.4f} "
          f"| worst row {int(err.amax(dim=-1).argmax())}{flag}", flush=True)


def oracle(wave):
    with torch.inference_mode():
        x = torch.from_numpy(wave)[None, None]
        feats = []
        for layer in ref.feature_extractor.conv_layers:
            x, _ = layer(x, None)
            feats.append(x[0].t().contiguous())
        proj = ref.encoder.feature_projection(feats[-1][None])
        xpos = ref.encoder.transformer._preprocess(proj)  # LayerNorm(x + pos conv): the base model's Transformer is built layer_norm_first
        layers, h = [], xpos
        for layer in ref.encoder.transformer.layers:
            h, _ = layer(h)
            layers.append(h[0])
        logits = ref.aux(h)[0]
        full, _ = ref(torch.from_numpy(wave)[None])
        assert float((full[0] - logits).abs().max()) < 1e-4, "staged oracle differs from the module's own forward"
    return feats, proj[0], xpos[0], layers, logits


for cfg in configs:
    durs = [float(v) for v in cfg.split(",")]
    waves = [synthetic_speech(d, seed=31 + i) for i, d in enumerate(durs)]
    P = max(2, -(-max(len(w) for w in waves) // 320))
    B = len(waves)
    print(f"=== batch of {durs} s, P = {P}", flush=True)
    refs = [oracle(w) for w in waves]
    dev, offs, lens = ours.upload(waves)
    for stage in [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 20]:
        ctx.debug_set("w2v_stop", stage)
        try:
            ours.emissions_device(dev, offs, lens)
        finally:
            ctx.debug_set("w2v_stop", -1)
        for b in range(B):
            feats, proj, xpos, layers, _ = refs[b]
            T = proj.shape[0]
            if stage <= 6:
                rows = P << (6 - stage)
                Tl = feats[stage].shape[0]
                got = ctx.debug_buffer(f"w2v.c{stage}", (Tl, 512), torch.bfloat16, offset_bytes=b * rows * 512 * 2)
                report(f"seg {b} conv layer {stage}", got, feats[stage])
            elif stage == 7:
                report(f"seg {b} feature projection", ctx.debug_buffer("w2v.x", (T, 768), torch.float32, offset_bytes=b * P * 768 * 4), proj)
            elif stage == 8:
                report(f"seg {b} LN(x + pos conv)", ctx.debug_buffer("w2v.x", (T, 768), torch.float32, offset_bytes=b * P * 768 * 4), xpos)
            else:
                l = stage - 9
                if l == 0:
                    a = ref.encoder.transformer.layers[0].attention
                    xp = xpos[None]
                    with torch.inference_mode():
                        q, k, v = a.q_proj(xp), a.k_proj(xp), a.v_proj(xp)
                        wq = torch.cat([q, k, v], -1)[0]
                        qh, kh, vh = [t.view(1, T, 12, 64).transpose(1, 2) for t in (q, k, v)]
                        wa = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(T, 768)
                    report(f"seg {b} layer 0 qkv", ctx.debug_buffer("w2v.qkv", (T, 2304), torch.bfloat16, offset_bytes=b * P * 2304 * 2), wq)
                    report(f"seg {b} layer 0 attention", ctx.debug_buffer("w2v.att", (T, 768), torch.bfloat16, offset_bytes=b * P * 768 * 2), wa)
                report(f"seg {b} transformer layer {l}", ctx.debug_buffer("w2v.x", (T, 768), torch.float32, offset_bytes=b * P * 768 * 4), layers[l])
    emis, t_off = ours.emissions_device(dev, offs, lens)
    for b in range(B):
        report(f"seg {b} emission logits", emis[t_off[b]:t_off[b + 1]], refs[b][4])
