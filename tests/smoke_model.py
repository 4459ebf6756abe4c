"""Model half of __graft_entry__.smoke(): one tiny encode + greedy decode on cuda:0 checked against the oracle."""
import numpy as np
import torch


def smoke_model(ctx):
    from fake_ctc_model import synthetic_speech
    from oracle import whisper as ow
    from whisperx.backends import b200_weights as bw
    import whisperx.audio as wa

    dims = dict(n_mels=80, n_audio_ctx=1500, n_audio_state=128, n_audio_head=2, n_audio_layer=1,
                n_vocab=1000, n_text_ctx=448, n_text_state=128, n_text_head=2, n_text_layer=1)
    w = bw.init_random_weights(dims, seed=1, std=0.05)
    kw = bw.to_kernel_layout(w, dims, "cuda")
    ctx.set_model(dims, kw)
    mel = wa.log_mel_chunks([synthetic_speech(3.0, seed=4), synthetic_speech(1.0, seed=5)], 80)
    enc = ctx.encode(mel)
    w_ref = bw.kernel_layout_to_openai_fp32(kw, dims)
    with torch.no_grad():
        ref_enc = ow.encoder_forward(w_ref, dims, mel.cpu())
        toks = np.array([[1, 2, 4, 9], [1, 2, 4, 11]], np.int32)
        cache = ow.DecoderCache(w_ref, dims, enc.float().cpu())
        ref_logits = ow.decoder_forward(w_ref, dims, torch.from_numpy(toks).long(), cache)
    e_enc = float((enc.float().cpu() - ref_enc).abs().max())
    e_log = float((ctx.decoder_logits(enc, toks).cpu() - ref_logits).abs().max())
    assert e_enc < 0.08 and e_log < 0.05, (e_enc, e_log)
    r = ctx.decode_greedy(enc, [1, 2, 4], 3, no_speech=5, sample_len=8)
    assert r["tokens"].shape == (2, 8)
    print(f"smoke model ok: encoder max-abs err {e_enc:.4f}, logits max-abs err {e_log:.4f}, launches={ctx.launches}")
