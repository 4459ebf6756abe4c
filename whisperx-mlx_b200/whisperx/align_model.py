"""
The B200 alignment model: wav2vec2-base CTC (torchaudio WAV2VEC2_ASR_BASE_960H architecture) whose forward pass runs on
the hand-written kernels of csrc/wxb_w2v.cu (tcgen05 GEMMs + flash attention + fp32 conv front layer), BATCHED over all
segments of a transcript.  It replaces the torch model the reference's align() calls once per segment
(/root/reference/whisperx/alignment.py:240-258); `whisperx.load_align_model` returns it for base-architecture bundles.

The object also answers the torchaudio call convention (`model(waveform [1, S], lengths=None) -> (logits [1, T, V], None)`), so
it can be handed to any code written for the reference's models.  There is no CPU fallback.
"""
import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

W2V_BASE_DIMS = dict(conv_dim=512, embed_dim=768, n_heads=12, n_layers=12, ff_dim=3072, pos_kernel=128, pos_groups=16, n_out=29)
_CONV = ((10, 5), (3, 2), (3, 2), (3, 2), (3, 2), (2, 2), (2, 2))
MIN_SAMPLES = 400  # the model's receptive field: shorter segments are zero-padded (alignment.py:243-249)


def frames_for(n_samples: int) -> int:
    t = int(n_samples)
    for k, s in _CONV:
        t = (t - k) // s + 1 if t >= k else 0
    return t


def forward_flops(n_samples: Sequence[int], dims: Dict[str, int] = W2V_BASE_DIMS) -> float:
    """Algorithmic FLOPs of one forward pass over segments of these lengths (convs, projections, attention, MLP, head)."""
    d, ff, L, c = dims["embed_dim"], dims["ff_dim"], dims["n_layers"], dims["conv_dim"]
    total = 0.0
    for S in n_samples:
        t = int(S)
        cin = 1
        for k, s in _CONV:
            t = (t - k) // s + 1 if t >= k else 0
            total += 2.0 * t * c * cin * k
            cin = c
        T = t
        total += 2.0 * T * c * d                                          # feature projection
        total += 2.0 * T * d * (d // dims["pos_groups"]) * dims["pos_kernel"]  # grouped positional conv
        total += L * (2.0 * T * d * 3 * d + 4.0 * T * T * d + 2.0 * T * d * d + 4.0 * T * d * ff)
        total += 2.0 * T * d * dims["n_out"]
    return total


def from_torchaudio_state_dict(sd: Dict[str, torch.Tensor], dims: Dict[str, int], device) -> Dict[str, torch.Tensor]:
    """torchaudio `Wav2Vec2Model.state_dict()` -> the kernel-layout tensors of wxb_set_align_model (include/wxb200.h)."""
    bf, f32 = torch.bfloat16, torch.float32
    d, G = dims["embed_dim"], dims["pos_groups"]
    out: Dict[str, torch.Tensor] = {}

    def put(name, t, dtype):
        out[name] = t.detach().to(device=device, dtype=dtype).contiguous()

    fe = "feature_extractor.conv_layers."
    if fe + "1.layer_norm.weight" in sd:
        raise ValueError("layer-norm feature extractors (wav2vec2 LARGE / XLSR) are not supported by the B200 alignment model")
    put("w2v.conv0.w", sd[fe + "0.conv.weight"].reshape(dims["conv_dim"], 10), f32)
    put("w2v.gn.w", sd[fe + "0.layer_norm.weight"], f32)
    put("w2v.gn.b", sd[fe + "0.layer_norm.bias"], f32)
    for l in range(1, 7):
        w = sd[fe + f"{l}.conv.weight"]                      # [co, ci, k] -> [co, k * ci], column = tap * ci + channel
        put(f"w2v.conv{l}.w", w.permute(0, 2, 1).reshape(w.shape[0], -1), bf)
    fp = "encoder.feature_projection."
    put("w2v.fp.ln.w", sd[fp + "layer_norm.weight"], f32); put("w2v.fp.ln.b", sd[fp + "layer_norm.bias"], f32)
    put("w2v.fp.w", sd[fp + "projection.weight"], bf); put("w2v.fp.b", sd[fp + "projection.bias"], f32)
    pc = "encoder.transformer.pos_conv_embed.conv."
    if pc + "weight" in sd:
        w = sd[pc + "weight"].float()
    else:  # weight norm over dim 2: w[:, :, k] = g[k] * v[:, :, k] / ||v[:, :, k]||
        g = sd.get(pc + "parametrizations.weight.original0", sd.get(pc + "weight_g")).float()
        v = sd.get(pc + "parametrizations.weight.original1", sd.get(pc + "weight_v")).float()
        w = g * v / v.norm(dim=(0, 1), keepdim=True)
    put("w2v.pos.w", w.permute(0, 2, 1).reshape(d, -1), bf)   # [d, 128 * d/G], column = tap * (d/G) + channel of the group
    put("w2v.pos.b", sd[pc + "bias"], f32)
    for i in range(dims["n_layers"]):
        s, o = f"encoder.transformer.layers.{i}.", f"w2v.{i}."
        put(o + "qkv.w", torch.cat([sd[s + "attention.q_proj.weight"], sd[s + "attention.k_proj.weight"], sd[s + "attention.v_proj.weight"]], 0), bf)
        put(o + "qkv.b", torch.cat([sd[s + "attention.q_proj.bias"], sd[s + "attention.k_proj.bias"], sd[s + "attention.v_proj.bias"]], 0), f32)
        put(o + "out.w", sd[s + "attention.out_proj.weight"], bf); put(o + "out.b", sd[s + "attention.out_proj.bias"], f32)
        put(o + "ln1.w", sd[s + "layer_norm.weight"], f32); put(o + "ln1.b", sd[s + "layer_norm.bias"], f32)
        put(o + "fc1.w", sd[s + "feed_forward.intermediate_dense.weight"], bf); put(o + "fc1.b", sd[s + "feed_forward.intermediate_dense.bias"], f32)
        put(o + "fc2.w", sd[s + "feed_forward.output_dense.weight"], bf); put(o + "fc2.b", sd[s + "feed_forward.output_dense.bias"], f32)
        put(o + "ln2.w", sd[s + "final_layer_norm.weight"], f32); put(o + "ln2.b", sd[s + "final_layer_norm.bias"], f32)
    put("w2v.ln.w", sd["encoder.transformer.layer_norm.weight"], f32); put("w2v.ln.b", sd["encoder.transformer.layer_norm.bias"], f32)
    put("w2v.aux.w", sd["aux.weight"], bf); put("w2v.aux.b", sd["aux.bias"], f32)
    return out


def kernel_layout_to_torchaudio(k: Dict[str, torch.Tensor], dims: Dict[str, int]) -> Dict[str, torch.Tensor]:
    """Inverse (fp32, CPU, weight norm removed: plain `conv.weight`): lets a torch oracle run on exactly the bf16-rounded
    numbers the kernels hold (tests)."""
    c = lambda n: k[n].detach().float().cpu()  # noqa: E731
    d, G, cd = dims["embed_dim"], dims["pos_groups"], dims["conv_dim"]
    sd = {}
    fe = "feature_extractor.conv_layers."
    sd[fe + "0.conv.weight"] = c("w2v.conv0.w").view(cd, 1, 10)
    sd[fe + "0.layer_norm.weight"], sd[fe + "0.layer_norm.bias"] = c("w2v.gn.w"), c("w2v.gn.b")
    for l, (kk, _) in enumerate(_CONV):
        if l:
            sd[fe + f"{l}.conv.weight"] = c(f"w2v.conv{l}.w").view(cd, kk, cd).permute(0, 2, 1).contiguous()
    fp = "encoder.feature_projection."
    sd[fp + "layer_norm.weight"], sd[fp + "layer_norm.bias"] = c("w2v.fp.ln.w"), c("w2v.fp.ln.b")
    sd[fp + "projection.weight"], sd[fp + "projection.bias"] = c("w2v.fp.w"), c("w2v.fp.b")
    pc = "encoder.transformer.pos_conv_embed.conv."
    sd[pc + "weight"] = c("w2v.pos.w").view(d, dims["pos_kernel"], d // G).permute(0, 2, 1).contiguous()
    sd[pc + "bias"] = c("w2v.pos.b")
    for i in range(dims["n_layers"]):
        s, o = f"encoder.transformer.layers.{i}.", f"w2v.{i}."
        qw, qb = c(o + "qkv.w"), c(o + "qkv.b")
        for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
            sd[s + f"attention.{nm}.weight"], sd[s + f"attention.{nm}.bias"] = qw[j * d:(j + 1) * d], qb[j * d:(j + 1) * d]
        sd[s + "attention.out_proj.weight"], sd[s + "attention.out_proj.bias"] = c(o + "out.w"), c(o + "out.b")
        sd[s + "layer_norm.weight"], sd[s + "layer_norm.bias"] = c(o + "ln1.w"), c(o + "ln1.b")
        sd[s + "feed_forward.intermediate_dense.weight"], sd[s + "feed_forward.intermediate_dense.bias"] = c(o + "fc1.w"), c(o + "fc1.b")
        sd[s + "feed_forward.output_dense.weight"], sd[s + "feed_forward.output_dense.bias"] = c(o + "fc2.w"), c(o + "fc2.b")
        sd[s + "final_layer_norm.weight"], sd[s + "final_layer_norm.bias"] = c(o + "ln2.w"), c(o + "ln2.b")
    sd["encoder.transformer.layer_norm.weight"], sd["encoder.transformer.layer_norm.bias"] = c("w2v.ln.w"), c("w2v.ln.b")
    sd["aux.weight"], sd["aux.bias"] = c("w2v.aux.w"), c("w2v.aux.b")
    return sd


class Wav2Vec2B200:
    """wav2vec2-base CTC model resident on one B200.  `weights` = torchaudio state dict (or kernel-layout dict)."""

    is_b200_native = True

    def __init__(self, weights: Dict[str, torch.Tensor], device="cuda", dims: Optional[Dict[str, int]] = None):
        from ._native import get_context
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("the B200 alignment model runs on CUDA sm_100a only; there is no CPU fallback")
        self.ctx = get_context(dev.index if dev.index is not None else torch.cuda.current_device())
        self.device = self.ctx.device
        self.dims = dict(dims or W2V_BASE_DIMS)
        if "aux.weight" in weights:
            self.dims["n_out"] = int(weights["aux.weight"].shape[0])
            self.kernel_weights = from_torchaudio_state_dict(weights, self.dims, self.device)
        else:
            self.dims["n_out"] = int(weights["w2v.aux.w"].shape[0])
            self.kernel_weights = {k: v.to(self.device).contiguous() for k, v in weights.items()}
        self.last_stats: Dict[str, float] = {}
        self._staging = None
        self._staging_done = None

    # torch.nn.Module look-alikes used by callers of the reference API
    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("the B200 alignment model cannot be moved off the GPU (no CPU fallback)")
        return self

    def eval(self):
        return self

    def _bind(self):
        if self.ctx.align_owner is not self:
            self.ctx.set_align_model(self.dims, self.kernel_weights, owner=self)

    # ------------------------------------------------------------------ device-resident batched forward
    def emissions_device(self, audio_dev: torch.Tensor, offs: np.ndarray, lens: np.ndarray) -> Tuple[torch.Tensor, np.ndarray]:
        """audio_dev f32 cuda; segment b = lens[b] (>= 400) samples at offs[b].  Returns (logits f32 [sum T, n_out], t_off
        int32 [B + 1])."""
        self._bind()
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        T = np.array([frames_for(int(s)) for s in lens], dtype=np.int64)
        t_off = np.concatenate([[0], np.cumsum(T)]).astype(np.int32)
        emis = torch.empty((max(int(t_off[-1]), 1), self.dims["n_out"]), dtype=torch.float32, device=self.device)
        self.ctx.w2v_emissions(audio_dev, offs, lens, emis, t_off)
        self.last_stats.update(flops=forward_flops(lens.tolist(), self.dims), frames=int(t_off[-1]), segments=len(lens))
        return emis[: int(t_off[-1])], t_off

    def upload(self, waves: List[np.ndarray]):
        """Host waveforms -> one packed device buffer through a pinned staging buffer.  Segments shorter than the receptive
        field are zero-padded to 400 samples (the reference pads them too, alignment.py:243-249)."""
        lens = np.array([max(len(w), MIN_SAMPLES) for w in waves], dtype=np.int32)
        offs = np.zeros(len(waves), dtype=np.int64)
        if len(waves) > 1:
            offs[1:] = np.cumsum(lens[:-1])
        total = int(lens.sum())
        if self._staging is None or self._staging.numel() < max(total, 1):
            self._staging = torch.empty(max(total, 1), dtype=torch.float32).pin_memory()
            self._staging_done = None
        if self._staging_done is not None:
            self._staging_done.synchronize()
        hv = self._staging.numpy()
        dev = torch.empty(max(total, 1), dtype=torch.float32, device=self.device)
        # Groups of segments: host copy into the pinned buffer, then an asynchronous H2D of that range, so the PCIe transfer of
        # one group runs under the host copy of the next; segments that are back-to-back views of one float32 array (what
        # align() cuts from a recording) are copied as one range with torch's multi-threaded copy (as the ASR backend does).
        from .backends.b200 import B200WhisperBackend
        n_seg, group = len(waves), 12
        for g0 in range(0, n_seg, group):
            g1 = min(n_seg, g0 + group)
            a, b = int(offs[g0]), int(offs[g1 - 1] + lens[g1 - 1])
            same_len = all(len(w) == int(l) for w, l in zip(waves[g0:g1], lens[g0:g1]))
            run = B200WhisperBackend._contiguous_run(waves[g0:g1], lens[g0:g1]) if same_len else None
            if run is not None:
                self._staging[a:b].copy_(torch.from_numpy(run))
            else:
                for w, o, l in zip(waves[g0:g1], offs[g0:g1], lens[g0:g1]):
                    n = len(w)
                    hv[o:o + n] = np.asarray(w, dtype=np.float32)
                    if n < l:
                        hv[o + n:o + l] = 0.0
            if b > a:
                dev[a:b].copy_(self._staging[a:b], non_blocking=True)
        self._staging_done = torch.cuda.Event()
        self._staging_done.record()
        self.last_stats["h2d_bytes"] = total * 4
        return dev, offs, lens

    def emissions(self, waves: List[np.ndarray]):
        dev, offs, lens = self.upload(waves)
        return self.emissions_device(dev, offs, lens)

    # ------------------------------------------------------------------ torchaudio call convention (one segment)
    def __call__(self, waveform: torch.Tensor, lengths: Optional[torch.Tensor] = None):
        if waveform.dim() == 1:
            waveform = waveform[None]
        outs = []
        for row in waveform:
            w = row.detach().to(self.device, dtype=torch.float32).contiguous()
            n = w.numel()
            if n < MIN_SAMPLES:
                w = torch.nn.functional.pad(w, (0, MIN_SAMPLES - n))
            e, _ = self.emissions_device(w, np.array([0]), np.array([w.numel()]))
            outs.append(e)
        return torch.stack(outs, 0), None


def random_init_torchaudio(params: dict, seed: int = 0):
    """torchaudio's own module for these architecture parameters, seeded random init (no checkpoint offline)."""
    import torchaudio
    torch.manual_seed(seed)
    return torchaudio.models.wav2vec2_model(**params).eval()
