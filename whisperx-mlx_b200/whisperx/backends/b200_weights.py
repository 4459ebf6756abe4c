"""
Model dimensions by name, random initialisation (no checkpoints are available offline) and
state-dict conversion for the B200 backend.  Tensors use OpenAI-Whisper parameter names
("encoder.blocks.0.attn.query.weight", ...), which is also the naming wxb_set_model expects.
"""
from typing import Dict

import torch

# name -> ModelDimensions
_DIMS = {
    "tiny": (80, 384, 6, 4, 51865, 384, 6, 4),
    "base": (80, 512, 8, 6, 51865, 512, 8, 6),
    "small": (80, 768, 12, 12, 51865, 768, 12, 12),
    "medium": (80, 1024, 16, 24, 51865, 1024, 16, 24),
    "large-v2": (80, 1280, 20, 32, 51865, 1280, 20, 32),
    "large-v3": (128, 1280, 20, 32, 51866, 1280, 20, 32),
    "large-v3-turbo": (128, 1280, 20, 32, 51866, 1280, 20, 4),
}
_ALIASES = {"turbo": "large-v3-turbo", "large": "large-v3", "whisper-large-v3": "large-v3",
            "whisper-large-v3-turbo": "large-v3-turbo"}


def canonical_name(name: str) -> str:
    n = name.lower().replace("openai/", "").replace("mlx-community/", "")
    if n.endswith("-mlx"):
        n = n[:-4]
    if n.startswith("whisper-"):
        n = n[len("whisper-"):]
    n = _ALIASES.get(n, n)
    if n not in _DIMS:
        raise ValueError(f"unknown Whisper architecture '{name}' (known: {sorted(_DIMS)})")
    return n


def dims_for(name: str) -> Dict[str, int]:
    m, ad, ah, al, v, td, th, tl = _DIMS[canonical_name(name)]
    return dict(n_mels=m, n_audio_ctx=1500, n_audio_state=ad, n_audio_head=ah, n_audio_layer=al,
                n_vocab=v, n_text_ctx=448, n_text_state=td, n_text_head=th, n_text_layer=tl)


def special_tokens(dims: Dict[str, int]) -> Dict[str, int]:
    """Multilingual special-token ids: large-v3 family has one more language token."""
    v3 = dims["n_vocab"] >= 51866
    base = dict(eot=50257, sot=50258, blank=220)
    if v3:
        base.update(transcribe=50360, translate=50359, no_speech=50363, no_timestamps=50364, timestamp_begin=50365)
    else:
        base.update(transcribe=50359, translate=50358, no_speech=50362, no_timestamps=50363, timestamp_begin=50364)
    return base


def _sinusoids(length: int, channels: int) -> torch.Tensor:
    import math
    inc = math.log(10000.0) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float32))
    t = torch.arange(length, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.cat([t.sin(), t.cos()], dim=1)


def init_random_weights(dims: Dict[str, int], seed: int = 0, std: float = 0.02, bias_std: float = 0.02,
                        ln_jitter: float = 0.1) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights (fp32, CPU).  Matrices N(0, std); biases N(0, bias_std);
    LayerNorm weight 1 + N(0, ln_jitter), bias N(0, bias_std) so that every parameter matters in
    parity tests."""
    g = torch.Generator().manual_seed(seed)

    def mat(*shape, s=std):
        return torch.randn(*shape, generator=g) * s

    w: Dict[str, torch.Tensor] = {}
    d, dm = dims["n_audio_state"], dims["n_mels"]
    w["encoder.conv1.weight"], w["encoder.conv1.bias"] = mat(d, dm, 3), mat(d, s=bias_std)
    w["encoder.conv2.weight"], w["encoder.conv2.bias"] = mat(d, d, 3), mat(d, s=bias_std)
    w["encoder.positional_embedding"] = _sinusoids(dims["n_audio_ctx"], d)

    def ln(prefix, width):
        w[prefix + ".weight"] = 1.0 + mat(width, s=ln_jitter)
        w[prefix + ".bias"] = mat(width, s=bias_std)

    def attn(prefix, width):
        for nm in ("query", "key", "value", "out"):
            w[f"{prefix}.{nm}.weight"] = mat(width, width)
            if nm != "key":
                w[f"{prefix}.{nm}.bias"] = mat(width, s=bias_std)

    def mlp(prefix, width):
        w[prefix + ".0.weight"], w[prefix + ".0.bias"] = mat(4 * width, width), mat(4 * width, s=bias_std)
        w[prefix + ".2.weight"], w[prefix + ".2.bias"] = mat(width, 4 * width), mat(width, s=bias_std)

    for i in range(dims["n_audio_layer"]):
        p = f"encoder.blocks.{i}"
        ln(p + ".attn_ln", d); attn(p + ".attn", d); ln(p + ".mlp_ln", d); mlp(p + ".mlp", d)
    ln("encoder.ln_post", d)
    t = dims["n_text_state"]
    w["decoder.token_embedding.weight"] = mat(dims["n_vocab"], t)
    w["decoder.positional_embedding"] = mat(dims["n_text_ctx"], t)
    for i in range(dims["n_text_layer"]):
        p = f"decoder.blocks.{i}"
        ln(p + ".attn_ln", t); attn(p + ".attn", t)
        ln(p + ".cross_attn_ln", t); attn(p + ".cross_attn", t)
        ln(p + ".mlp_ln", t); mlp(p + ".mlp", t)
    ln("decoder.ln", t)
    return w


_HF_RENAMES = (
    ("model.encoder.", "encoder."), ("model.decoder.", "decoder."),
    (".layers.", ".blocks."), ("embed_positions.weight", "positional_embedding"),
    ("embed_tokens.", "token_embedding."), (".self_attn_layer_norm.", ".attn_ln."),
    (".encoder_attn_layer_norm.", ".cross_attn_ln."), (".final_layer_norm.", ".mlp_ln."),
    (".self_attn.", ".attn."), (".encoder_attn.", ".cross_attn."),
    (".q_proj.", ".query."), (".k_proj.", ".key."), (".v_proj.", ".value."), (".out_proj.", ".out."),
    (".fc1.", ".mlp.0."), (".fc2.", ".mlp.2."),
    ("encoder.layer_norm.", "encoder.ln_post."), ("decoder.layer_norm.", "decoder.ln."),
)


def from_hf_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """transformers' WhisperForConditionalGeneration names -> OpenAI names (fp32 copies)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("proj_out."):
            continue  # tied to the token embedding
        n = k
        for a, b in _HF_RENAMES:
            n = n.replace(a, b)
        out[n] = v.detach().float().clone()
    return out


def round_to_bf16(w: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """What the kernels see, as fp32: every >=2-D matrix rounded to bf16 (biases / LN / positions of
    the ENCODER stay fp32; the decoder's positional table is bf16 like its embedding)."""
    out = {}
    for k, v in w.items():
        if v.dim() >= 2 and k != "encoder.positional_embedding":
            out[k] = v.to(torch.bfloat16).float()
        else:
            out[k] = v.float()
    return out
