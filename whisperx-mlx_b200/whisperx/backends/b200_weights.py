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


def infer_dims(w: Dict[str, torch.Tensor]) -> Dict[str, int]:
    """ModelDimensions from the shapes of an OpenAI-named state dict (what a checkpoint's config would say)."""
    def n_blocks(prefix):
        return 1 + max(int(k.split(".")[2]) for k in w if k.startswith(prefix + ".blocks."))
    n_mels = int(w["encoder.conv1.weight"].shape[1])
    ad = int(w["encoder.conv1.weight"].shape[0])
    n_vocab, td = (int(v) for v in w["decoder.token_embedding.weight"].shape)
    return dict(n_mels=n_mels, n_audio_ctx=int(w["encoder.positional_embedding"].shape[0]), n_audio_state=ad, n_audio_head=ad // 64,
                n_audio_layer=n_blocks("encoder"), n_vocab=n_vocab, n_text_ctx=int(w["decoder.positional_embedding"].shape[0]),
                n_text_state=td, n_text_head=td // 64, n_text_layer=n_blocks("decoder"))


def _from_mlx_npz(path: str) -> Dict[str, torch.Tensor]:
    """mlx-community `weights.npz` (what the reference's backends load, mlx_lightning.py:73): OpenAI names, fp16, Conv1d kernels
    stored [out, k, in] (MLX layout) instead of [out, in, k]."""
    import numpy as np
    out = {}
    with np.load(path) as z:
        for k in z.files:
            t = torch.from_numpy(np.asarray(z[k]).astype(np.float32))
            if k in ("encoder.conv1.weight", "encoder.conv2.weight") and t.shape[1] == 3:
                t = t.permute(0, 2, 1).contiguous()
            out[k] = t
    return out


def load_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """A checkpoint on disk -> OpenAI-named fp32 state dict.  Accepted: a Hugging Face directory (model.safetensors, sharded
    model.safetensors.index.json, or pytorch_model.bin), a single .safetensors file, an OpenAI .pt ({"dims", "model_state_dict"}),
    an mlx-community directory / weights.npz, or a torch file holding a plain state dict."""
    import json
    import os
    if os.path.isdir(path):
        for name in ("model.safetensors", "model.safetensors.index.json", "weights.safetensors", "weights.npz", "pytorch_model.bin"):
            f = os.path.join(path, name)
            if os.path.exists(f):
                return load_checkpoint(f)
        raise FileNotFoundError(f"no model.safetensors / weights.npz / pytorch_model.bin under {path}")
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    if path.endswith(".index.json"):
        from safetensors.torch import load_file
        with open(path) as fh:
            shards = sorted(set(json.load(fh)["weight_map"].values()))
        sd = {}
        for shard in shards:
            sd.update(load_file(os.path.join(os.path.dirname(path), shard)))
    elif path.endswith(".safetensors"):
        from safetensors.torch import load_file
        sd = load_file(path)
    elif path.endswith(".npz"):
        sd = _from_mlx_npz(path)
    else:
        sd = torch.load(path, map_location="cpu", weights_only=True)
        sd = sd.get("model_state_dict", sd)
    if any(k.startswith("model.encoder.") for k in sd):
        return from_hf_state_dict(sd)
    if "encoder.conv1.weight" not in sd or "decoder.token_embedding.weight" not in sd:
        raise ValueError(f"{path}: not a Whisper checkpoint (neither transformers nor OpenAI / MLX parameter names)")
    sd = {k: v.detach().float() for k, v in sd.items()}
    if "encoder.positional_embedding" not in sd:  # HF / MLX exports may leave the fixed sinusoid table out
        sd["encoder.positional_embedding"] = _sinusoids(1500, int(sd["encoder.conv1.weight"].shape[0]))
    return sd


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


def to_kernel_layout(w: Dict[str, torch.Tensor], dims: Dict[str, int], device) -> Dict[str, torch.Tensor]:
    """OpenAI-named fp32 weights -> the device tensors wxb_set_model consumes (csrc/wxb_model.cu):
    bf16 [N, K] matrices (Q|K|V fused, conv kernels flattened tap-major for the implicit-im2col GEMM,
    cross K|V fused), f32 biases / LayerNorm parameters / positional tables."""
    bf, f32 = torch.bfloat16, torch.float32
    out: Dict[str, torch.Tensor] = {}

    def put(name, t, dtype):
        out[name] = t.to(device=device, dtype=dtype).contiguous()

    def conv_flat(k):  # [co, ci, 3] -> [co, 3*ci] with column index tap*ci + channel
        return k.permute(0, 2, 1).reshape(k.shape[0], -1)

    put("enc.conv1.w", conv_flat(w["encoder.conv1.weight"]), bf)
    put("enc.conv1.b", w["encoder.conv1.bias"], f32)
    put("enc.conv2.w", conv_flat(w["encoder.conv2.weight"]), bf)
    put("enc.conv2.b", w["encoder.conv2.bias"], f32)
    put("enc.pos", w["encoder.positional_embedding"], f32)

    def fused_qkv(p, width):
        wq, wk, wv = w[p + ".query.weight"], w[p + ".key.weight"], w[p + ".value.weight"]
        bq, bv = w[p + ".query.bias"], w[p + ".value.bias"]
        return torch.cat([wq, wk, wv], 0), torch.cat([bq, torch.zeros(width), bv], 0)

    for i in range(dims["n_audio_layer"]):
        s, d = f"encoder.blocks.{i}", f"enc.{i}"
        width = dims["n_audio_state"]
        put(d + ".ln1.w", w[s + ".attn_ln.weight"], f32); put(d + ".ln1.b", w[s + ".attn_ln.bias"], f32)
        qw, qb = fused_qkv(s + ".attn", width)
        put(d + ".qkv.w", qw, bf); put(d + ".qkv.b", qb, f32)
        put(d + ".out.w", w[s + ".attn.out.weight"], bf); put(d + ".out.b", w[s + ".attn.out.bias"], f32)
        put(d + ".ln2.w", w[s + ".mlp_ln.weight"], f32); put(d + ".ln2.b", w[s + ".mlp_ln.bias"], f32)
        put(d + ".fc1.w", w[s + ".mlp.0.weight"], bf); put(d + ".fc1.b", w[s + ".mlp.0.bias"], f32)
        put(d + ".fc2.w", w[s + ".mlp.2.weight"], bf); put(d + ".fc2.b", w[s + ".mlp.2.bias"], f32)
    put("enc.ln_post.w", w["encoder.ln_post.weight"], f32); put("enc.ln_post.b", w["encoder.ln_post.bias"], f32)

    put("dec.emb", w["decoder.token_embedding.weight"], bf)
    put("dec.pos", w["decoder.positional_embedding"].to(bf).float(), f32)
    width = dims["n_text_state"]
    for i in range(dims["n_text_layer"]):
        s, d = f"decoder.blocks.{i}", f"dec.{i}"
        put(d + ".ln1.w", w[s + ".attn_ln.weight"], f32); put(d + ".ln1.b", w[s + ".attn_ln.bias"], f32)
        qw, qb = fused_qkv(s + ".attn", width)
        put(d + ".qkv.w", qw, bf); put(d + ".qkv.b", qb, f32)
        put(d + ".out.w", w[s + ".attn.out.weight"], bf); put(d + ".out.b", w[s + ".attn.out.bias"], f32)
        put(d + ".ln2.w", w[s + ".cross_attn_ln.weight"], f32); put(d + ".ln2.b", w[s + ".cross_attn_ln.bias"], f32)
        put(d + ".cq.w", w[s + ".cross_attn.query.weight"], bf); put(d + ".cq.b", w[s + ".cross_attn.query.bias"], f32)
        put(d + ".ckv.w", torch.cat([w[s + ".cross_attn.key.weight"], w[s + ".cross_attn.value.weight"]], 0), bf)
        put(d + ".ckv.b", torch.cat([torch.zeros(width), w[s + ".cross_attn.value.bias"]], 0), f32)
        put(d + ".cout.w", w[s + ".cross_attn.out.weight"], bf); put(d + ".cout.b", w[s + ".cross_attn.out.bias"], f32)
        put(d + ".ln3.w", w[s + ".mlp_ln.weight"], f32); put(d + ".ln3.b", w[s + ".mlp_ln.bias"], f32)
        put(d + ".fc1.w", w[s + ".mlp.0.weight"], bf); put(d + ".fc1.b", w[s + ".mlp.0.bias"], f32)
        put(d + ".fc2.w", w[s + ".mlp.2.weight"], bf); put(d + ".fc2.b", w[s + ".mlp.2.bias"], f32)
    put("dec.ln.w", w["decoder.ln.weight"], f32); put("dec.ln.b", w["decoder.ln.bias"], f32)
    return out


def expected_kernel_tensors(dims: Dict[str, int]) -> Dict[str, tuple]:
    """name -> (shape, dtype) of every tensor wxb_set_model borrows for these dims (the library reads them through raw
    pointers, so a checkpoint of another architecture or dtype must be rejected before it gets there)."""
    bf, f32 = torch.bfloat16, torch.float32
    exp: Dict[str, tuple] = {}
    d, nm = dims["n_audio_state"], dims["n_mels"]
    exp["enc.conv1.w"], exp["enc.conv1.b"] = ((d, 3 * nm), bf), ((d,), f32)
    exp["enc.conv2.w"], exp["enc.conv2.b"] = ((d, 3 * d), bf), ((d,), f32)
    exp["enc.pos"] = ((dims["n_audio_ctx"], d), f32)

    def block(p, width, cross):
        for ln in (".ln1", ".ln2") + ((".ln3",) if cross else ()):
            exp[p + ln + ".w"], exp[p + ln + ".b"] = ((width,), f32), ((width,), f32)
        exp[p + ".qkv.w"], exp[p + ".qkv.b"] = ((3 * width, width), bf), ((3 * width,), f32)
        exp[p + ".out.w"], exp[p + ".out.b"] = ((width, width), bf), ((width,), f32)
        if cross:
            exp[p + ".cq.w"], exp[p + ".cq.b"] = ((width, width), bf), ((width,), f32)
            exp[p + ".ckv.w"], exp[p + ".ckv.b"] = ((2 * width, width), bf), ((2 * width,), f32)
            exp[p + ".cout.w"], exp[p + ".cout.b"] = ((width, width), bf), ((width,), f32)
        exp[p + ".fc1.w"], exp[p + ".fc1.b"] = ((4 * width, width), bf), ((4 * width,), f32)
        exp[p + ".fc2.w"], exp[p + ".fc2.b"] = ((width, 4 * width), bf), ((width,), f32)

    for i in range(dims["n_audio_layer"]):
        block(f"enc.{i}", d, False)
    exp["enc.ln_post.w"], exp["enc.ln_post.b"] = ((d,), f32), ((d,), f32)
    t = dims["n_text_state"]
    exp["dec.emb"] = ((dims["n_vocab"], t), bf)
    exp["dec.pos"] = ((dims["n_text_ctx"], t), f32)
    for i in range(dims["n_text_layer"]):
        block(f"dec.{i}", t, True)
    exp["dec.ln.w"], exp["dec.ln.b"] = ((t,), f32), ((t,), f32)
    return exp


def validate_kernel_weights(k: Dict[str, torch.Tensor], dims: Dict[str, int]) -> None:
    """Raise ValueError unless `k` holds exactly the tensors of expected_kernel_tensors(dims), on a CUDA device, contiguous."""
    exp = expected_kernel_tensors(dims)
    missing = sorted(set(exp) - set(k))
    if missing:
        raise ValueError(f"weights: {len(missing)} tensors missing for these dims, e.g. {missing[:4]}")
    for name, (shape, dtype) in exp.items():
        t = k[name]
        if tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            raise ValueError(f"weights: '{name}' is {tuple(t.shape)} {t.dtype}, expected {tuple(shape)} {dtype} "
                             "(checkpoint of another architecture, or not in kernel layout)")
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError(f"weights: '{name}' must be a contiguous CUDA tensor")


def random_kernel_weights_on_device(dims: Dict[str, int], device, seed: int = 0, std: float = 0.02):
    """Large models: draw the random-init weights directly on the GPU in kernel layout (avoids a
    multi-GB CPU generate + copy).  Same tensor set as to_kernel_layout(init_random_weights(...))."""
    g = torch.Generator(device=device).manual_seed(seed)
    bf, f32 = torch.bfloat16, torch.float32
    out: Dict[str, torch.Tensor] = {}

    def mat(name, *shape, dtype=bf, s=std):
        out[name] = (torch.randn(*shape, generator=g, device=device, dtype=f32) * s).to(dtype)

    def ln(name, width):
        out[name + ".w"] = 1.0 + torch.randn(width, generator=g, device=device) * 0.1
        mat(name + ".b", width, dtype=f32)

    d, nm = dims["n_audio_state"], dims["n_mels"]
    mat("enc.conv1.w", d, 3 * nm); mat("enc.conv1.b", d, dtype=f32)
    mat("enc.conv2.w", d, 3 * d); mat("enc.conv2.b", d, dtype=f32)
    out["enc.pos"] = _sinusoids(dims["n_audio_ctx"], d).to(device)

    def block(p, width, cross):
        ln(p + ".ln1", width)
        mat(p + ".qkv.w", 3 * width, width); mat(p + ".qkv.b", 3 * width, dtype=f32)
        out[p + ".qkv.b"][width:2 * width] = 0
        mat(p + ".out.w", width, width); mat(p + ".out.b", width, dtype=f32)
        ln(p + ".ln2", width)
        if cross:
            mat(p + ".cq.w", width, width); mat(p + ".cq.b", width, dtype=f32)
            mat(p + ".ckv.w", 2 * width, width); mat(p + ".ckv.b", 2 * width, dtype=f32)
            out[p + ".ckv.b"][:width] = 0
            mat(p + ".cout.w", width, width); mat(p + ".cout.b", width, dtype=f32)
            ln(p + ".ln3", width)
        mat(p + ".fc1.w", 4 * width, width); mat(p + ".fc1.b", 4 * width, dtype=f32)
        mat(p + ".fc2.w", width, 4 * width); mat(p + ".fc2.b", width, dtype=f32)

    for i in range(dims["n_audio_layer"]):
        block(f"enc.{i}", d, False)
    ln("enc.ln_post", d)
    t = dims["n_text_state"]
    mat("dec.emb", dims["n_vocab"], t)
    out["dec.pos"] = (torch.randn(dims["n_text_ctx"], t, generator=g, device=device) * std).to(bf).float()
    for i in range(dims["n_text_layer"]):
        block(f"dec.{i}", t, True)
    ln("dec.ln", t)
    return out


def kernel_layout_to_openai_fp32(k: Dict[str, torch.Tensor], dims: Dict[str, int]) -> Dict[str, torch.Tensor]:
    """Inverse of to_kernel_layout (fp32, CPU): lets the oracle run on exactly the (bf16-rounded)
    numbers the kernels hold, whichever way the model was initialised."""
    w: Dict[str, torch.Tensor] = {}
    c = lambda n: k[n].detach().float().cpu()  # noqa: E731
    nm, d = dims["n_mels"], dims["n_audio_state"]
    w["encoder.conv1.weight"] = c("enc.conv1.w").view(d, 3, nm).permute(0, 2, 1).contiguous()
    w["encoder.conv1.bias"] = c("enc.conv1.b")
    w["encoder.conv2.weight"] = c("enc.conv2.w").view(d, 3, d).permute(0, 2, 1).contiguous()
    w["encoder.conv2.bias"] = c("enc.conv2.b")
    w["encoder.positional_embedding"] = c("enc.pos")

    def unfuse(src, dst, width):
        qw, qb = c(src + ".qkv.w"), c(src + ".qkv.b")
        w[dst + ".query.weight"], w[dst + ".key.weight"], w[dst + ".value.weight"] = qw[:width], qw[width:2 * width], qw[2 * width:]
        w[dst + ".query.bias"], w[dst + ".value.bias"] = qb[:width], qb[2 * width:]
        w[dst + ".out.weight"], w[dst + ".out.bias"] = c(src + ".out.w"), c(src + ".out.b")

    def mlp(src, dst):
        w[dst + ".0.weight"], w[dst + ".0.bias"] = c(src + ".fc1.w"), c(src + ".fc1.b")
        w[dst + ".2.weight"], w[dst + ".2.bias"] = c(src + ".fc2.w"), c(src + ".fc2.b")

    for i in range(dims["n_audio_layer"]):
        s, o = f"enc.{i}", f"encoder.blocks.{i}"
        w[o + ".attn_ln.weight"], w[o + ".attn_ln.bias"] = c(s + ".ln1.w"), c(s + ".ln1.b")
        unfuse(s, o + ".attn", d)
        w[o + ".mlp_ln.weight"], w[o + ".mlp_ln.bias"] = c(s + ".ln2.w"), c(s + ".ln2.b")
        mlp(s, o + ".mlp")
    w["encoder.ln_post.weight"], w["encoder.ln_post.bias"] = c("enc.ln_post.w"), c("enc.ln_post.b")
    t = dims["n_text_state"]
    w["decoder.token_embedding.weight"] = c("dec.emb")
    w["decoder.positional_embedding"] = c("dec.pos")
    for i in range(dims["n_text_layer"]):
        s, o = f"dec.{i}", f"decoder.blocks.{i}"
        w[o + ".attn_ln.weight"], w[o + ".attn_ln.bias"] = c(s + ".ln1.w"), c(s + ".ln1.b")
        unfuse(s, o + ".attn", t)
        w[o + ".cross_attn_ln.weight"], w[o + ".cross_attn_ln.bias"] = c(s + ".ln2.w"), c(s + ".ln2.b")
        ckw, ckb = c(s + ".ckv.w"), c(s + ".ckv.b")
        w[o + ".cross_attn.query.weight"], w[o + ".cross_attn.query.bias"] = c(s + ".cq.w"), c(s + ".cq.b")
        w[o + ".cross_attn.key.weight"], w[o + ".cross_attn.value.weight"] = ckw[:t], ckw[t:]
        w[o + ".cross_attn.value.bias"] = ckb[t:]
        w[o + ".cross_attn.out.weight"], w[o + ".cross_attn.out.bias"] = c(s + ".cout.w"), c(s + ".cout.b")
        w[o + ".mlp_ln.weight"], w[o + ".mlp_ln.bias"] = c(s + ".ln3.w"), c(s + ".ln3.b")
        mlp(s, o + ".mlp")
    w["decoder.ln.weight"], w["decoder.ln.bias"] = c("dec.ln.w"), c("dec.ln.b")
    return w
