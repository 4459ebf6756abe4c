"""Checkpoint formats: Hugging Face safetensors (single file, sharded index), OpenAI .pt, mlx-community weights.npz
all load into the same OpenAI-named fp32 state dict.  CPU only (the GPU side is tests/test_gpu_checkpoint.py)."""
import json
import os

import numpy as np
import pytest
import torch

from whisperx.backends import b200_weights as bw


@pytest.fixture(scope="module")
def hf_model():
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    torch.manual_seed(0)
    cfg = WhisperConfig(vocab_size=300, num_mel_bins=80, d_model=64, encoder_layers=2, decoder_layers=3, encoder_attention_heads=1,
                        decoder_attention_heads=1, encoder_ffn_dim=256, decoder_ffn_dim=256, max_source_positions=1500,
                        max_target_positions=448, pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1)
    return WhisperForConditionalGeneration(cfg).eval()


def _same(a, b):
    assert set(a) == set(b), set(a) ^ set(b)
    for k in a:
        assert torch.equal(a[k].float(), b[k].float()), k


def test_hf_directory_single_and_sharded(hf_model, tmp_path):
    want = bw.from_hf_state_dict(hf_model.state_dict())
    one = tmp_path / "one"
    hf_model.save_pretrained(one, safe_serialization=True)
    _same(bw.load_checkpoint(str(one)), want)
    _same(bw.load_checkpoint(str(one / "model.safetensors")), want)
    many = tmp_path / "many"
    hf_model.save_pretrained(many, safe_serialization=True, max_shard_size="1MB")
    assert os.path.exists(many / "model.safetensors.index.json")
    _same(bw.load_checkpoint(str(many)), want)
    dims = bw.infer_dims(want)
    assert (dims["n_audio_state"], dims["n_audio_layer"], dims["n_text_layer"], dims["n_vocab"], dims["n_mels"]) == (64, 2, 3, 300, 80)
    assert (dims["n_audio_ctx"], dims["n_text_ctx"]) == (1500, 448)


def test_openai_pt_and_mlx_npz(hf_model, tmp_path):
    want = bw.from_hf_state_dict(hf_model.state_dict())
    torch.save({"dims": {"n_mels": 80}, "model_state_dict": want}, tmp_path / "tiny.pt")
    _same(bw.load_checkpoint(str(tmp_path / "tiny.pt")), want)
    # mlx-community layout: fp16 npz, Conv1d kernels [out, k, in], no encoder positional table
    mlx = {k: v.numpy().astype(np.float16) for k, v in want.items() if k != "encoder.positional_embedding"}
    for k in ("encoder.conv1.weight", "encoder.conv2.weight"):
        mlx[k] = np.ascontiguousarray(mlx[k].transpose(0, 2, 1))
    os.makedirs(tmp_path / "mlx")
    np.savez(tmp_path / "mlx" / "weights.npz", **mlx)
    with open(tmp_path / "mlx" / "config.json", "w") as fh:
        json.dump({"n_mels": 80}, fh)
    got = bw.load_checkpoint(str(tmp_path / "mlx"))
    assert set(got) == set(want)
    for k in want:
        ref = want[k] if k == "encoder.positional_embedding" else want[k].half().float()
        assert got[k].shape == want[k].shape and torch.allclose(got[k], ref, atol=1e-6), k


def test_rejects_foreign_files(tmp_path):
    torch.save({"some": torch.zeros(3)}, tmp_path / "x.pt")
    with pytest.raises(ValueError):
        bw.load_checkpoint(str(tmp_path / "x.pt"))
    with pytest.raises(FileNotFoundError):
        bw.load_checkpoint(str(tmp_path / "missing"))
