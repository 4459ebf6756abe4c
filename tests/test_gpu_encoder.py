"""GPU parity for K2: the tcgen05 GEMM (vs a torch fp32 reference of the same op) and the full
encoder (vs the torch-CPU oracle run on the same bf16-rounded weights)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 128, 64), (256, 512, 128), (1000, 384, 384), (1500, 1280, 1280),
                                   (3000, 3840, 1280), (2900, 3840, 1280), (777, 5120, 1280), (513, 1280, 5120), (64, 1536, 512), (300, 200, 72)])
def test_gemm_bf16_vs_torch(wxb_ctx, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = A.float() @ W.float().t() + bias
    got = wxb_ctx.gemm_bf16(A, W, bias, out_f32=True)
    # fp32 accumulation of exact bf16 products: only summation-order noise
    assert _rel(got, ref) < 2e-5, _rel(got, ref)
    got16 = wxb_ctx.gemm_bf16(A, W, bias, gelu=True)
    ref16 = torch.nn.functional.gelu(ref)
    assert got16.dtype == torch.bfloat16
    assert _rel(got16.float(), ref16) < 6e-3  # bf16 output rounding (2^-8 relative)
    nob = wxb_ctx.gemm_bf16(A, W, None, out_f32=True)
    assert _rel(nob, A.float() @ W.float().t()) < 2e-5


def _tiny_dims(n_mels=80, d=128, heads=2, layers=2, vocab=1000):
    return dict(n_mels=n_mels, n_audio_ctx=1500, n_audio_state=d, n_audio_head=heads, n_audio_layer=layers,
                n_vocab=vocab, n_text_ctx=448, n_text_state=d, n_text_head=heads, n_text_layer=layers)


@pytest.mark.parametrize("cfg", ["mini80", "mini128", "tiny", "base"])
def test_encoder_vs_oracle(wxb_ctx, cfg):
    """Encoder hidden states vs the fp32 oracle on the same (bf16-rounded) weights.  Tolerance is the
    bf16 budget: activations are rounded to bf16 (2^-9 relative) at every GEMM input; the stated
    max-abs error is relative to the output scale (ln_post output is O(1))."""
    from oracle import whisper as ow
    from whisperx.backends import b200_weights as bw
    from fake_ctc_model import synthetic_speech
    import whisperx.audio as wa

    if cfg in ("tiny", "base"):  # "base" is the architecture of BASELINE config 2
        dims = bw.dims_for(cfg)
    elif cfg == "mini80":
        dims = _tiny_dims(80, 128, 2, 2)
    else:
        dims = _tiny_dims(128, 256, 4, 1)
    w = bw.init_random_weights(dims, seed=3, std=0.05)
    kw = bw.to_kernel_layout(w, dims, "cuda")
    wxb_ctx.set_model(dims, kw)
    B = 2
    chunks = [synthetic_speech(30.0, seed=21), synthetic_speech(9.0, seed=22)]
    mel = wa.log_mel_chunks(chunks, dims["n_mels"])
    got = wxb_ctx.encode(mel).float().cpu()
    assert got.shape == (B, 1500, dims["n_audio_state"])
    w_ref = bw.kernel_layout_to_openai_fp32(kw, dims)
    with torch.no_grad():
        ref = ow.encoder_forward(w_ref, dims, mel.cpu())
    err = (got - ref).abs()
    scale = float(ref.abs().max())
    print(f"[{cfg}] encoder max-abs err {float(err.max()):.4f} (mean {float(err.mean()):.5f}) at output scale {scale:.2f}")
    assert float(err.max()) <= 0.06 * max(1.0, scale), float(err.max())
    assert float(err.mean()) <= 0.006 * max(1.0, scale)


def test_encoder_chunk_groups_bit_identical(wxb_ctx):
    """The layer stack over groups of chunks (wxb_encoder.cu: chunk groups, `enc_group`) returns the same bits as the whole
    batch at once: group sizes 1, 2 (ragged last group of 1) and 4 on a 5-chunk batch."""
    from whisperx.backends import b200_weights as bw
    from fake_ctc_model import synthetic_speech
    import whisperx.audio as wa

    dims = _tiny_dims(80, 128, 2, 2)
    wxb_ctx.set_model(dims, bw.to_kernel_layout(bw.init_random_weights(dims, seed=4, std=0.05), dims, "cuda"))
    chunks = [synthetic_speech(30.0 if i % 2 == 0 else 7.0 + i, seed=40 + i) for i in range(5)]
    mel = wa.log_mel_chunks(chunks, 80)
    try:
        wxb_ctx.debug_set("enc_group", 0)
        whole = wxb_ctx.encode(mel).clone()
        for g in (1, 2, 4):
            wxb_ctx.debug_set("enc_group", g)
            assert torch.equal(wxb_ctx.encode(mel), whole), f"enc_group={g}"
    finally:
        wxb_ctx.debug_set("enc_group", -1)


@pytest.mark.parametrize("B,T,H", [(1, 1500, 2), (2, 1500, 6), (3, 700, 3), (1, 128, 1), (2, 129, 2)])
def test_encoder_attention_vs_torch(wxb_ctx, B, T, H):
    """Stand-alone self-attention vs torch fp32 SDPA on the same bf16 inputs.  P is rounded to bf16 before the PV
    product (2^-9 relative per term), the output to bf16: max-abs tolerance 0.02 at |V| ~ 1."""
    d = 64 * H
    g = torch.Generator(device="cuda").manual_seed(B * 100 + T + H)
    qkv = torch.randn(B * T, 3 * d, device="cuda", generator=g)
    qkv[:, :d] *= 2.0  # scores of a few units: a real softmax, not a near-uniform one
    qkv = qkv.to(torch.bfloat16)
    got = wxb_ctx.encoder_attention(qkv, B, T, H).float()
    q, k, v = [x.float().view(B, T, H, 64).transpose(1, 2) for x in qkv.split(d, dim=1)]
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, d)
    err = float((got - ref).abs().max())
    print(f"[B={B} T={T} H={H}] attention max-abs err {err:.4f}")
    assert err < 0.02, err
