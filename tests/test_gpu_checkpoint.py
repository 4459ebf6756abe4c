"""A checkpoint in the real on-disk format -> the kernels: a whisper-tiny sized transformers model is saved with
save_pretrained (safetensors), loaded through whisperx.load_model(..., weights=<dir>), and the GPU encoder / decoder must reproduce
the transformers model's own outputs (fp32 CPU) within the bf16 tolerance of the oracle tests."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_hf_safetensors_checkpoint_end_to_end(wxb_ctx, tmp_path):
    import whisperx
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    torch.manual_seed(1)
    cfg = WhisperConfig(vocab_size=51865, num_mel_bins=80, d_model=384, encoder_layers=4, decoder_layers=4, encoder_attention_heads=6,
                        decoder_attention_heads=6, encoder_ffn_dim=1536, decoder_ffn_dim=1536, max_source_positions=1500,
                        max_target_positions=448, pad_token_id=50257, bos_token_id=50257, eos_token_id=50257, decoder_start_token_id=50258)
    hf = WhisperForConditionalGeneration(cfg).eval()
    with torch.no_grad():
        for p in hf.parameters():
            p.add_(0.02 * torch.randn_like(p))
    hf.save_pretrained(tmp_path / "whisper-tiny", safe_serialization=True)
    with pytest.raises(ValueError):  # the same files under another architecture name must be refused, not mis-read
        whisperx.load_model("base", device="cuda", backend="b200", vad_method=None, language="en", weights=str(tmp_path / "whisper-tiny"))
    model = whisperx.load_model("openai/whisper-tiny", device="cuda", backend="b200", vad_method=None, language="en",
                                weights=str(tmp_path / "whisper-tiny"))
    be = model.backend
    be._bind()
    mel = torch.randn(2, 80, 3000, generator=torch.Generator().manual_seed(3)) * 0.5
    toks = np.random.RandomState(0).randint(0, 50000, size=(2, 9)).astype(np.int32)
    enc = be.ctx.encode(mel.cuda())
    got = be.ctx.decoder_logits(enc, toks).float().cpu()
    with torch.no_grad():
        ref_enc = hf.model.encoder(mel).last_hidden_state
        ref = hf(input_features=mel, decoder_input_ids=torch.from_numpy(toks).long()).logits
    e_err = (enc.float().cpu() - ref_enc).abs().max().item()
    l_err = (got - ref).abs()
    print(f"transformers checkpoint: encoder max-abs err {e_err:.4f} at scale {ref_enc.abs().max().item():.2f}; logits max-abs err "
          f"{l_err.max().item():.4f} (mean {l_err.mean().item():.5f}) at logit std {ref.std().item():.3f}")
    assert e_err <= 0.06 * max(1.0, ref_enc.abs().max().item())
    # the whole chain (bf16 weights, bf16 encoder output, bf16 activations) against an fp32 model: looser than the decoder-only gate
    assert l_err.max().item() <= 0.08 * max(ref.std().item(), 1.0) + 0.03 and l_err.mean().item() <= 0.012 * max(ref.std().item(), 1.0) + 0.004
    # a download_root that holds the checkpoint under the model's name is picked up without `weights=`
    m2 = whisperx.load_model("whisper-tiny", device="cuda", backend="b200", vad_method=None, language="en", download_root=str(tmp_path))
    assert torch.equal(m2.backend.kernel_weights["dec.emb"], be.kernel_weights["dec.emb"])
