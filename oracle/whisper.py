"""
ORACLE — test infrastructure only.  Never imported by the product path (whisperx-mlx_b200/);
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.

Torch-CPU fp32 restatement of the Whisper model and of the reference's batched greedy loop.

The model arithmetic is NOT in /root/reference: the reference delegates it to the un-vendored
third-party package `mlx-whisper` (pyproject.toml:13, git branch pin `whisperx-optimizations`,
no commit hash; on top of mlx>=0.26.0).  This file restates the published OpenAI Whisper
architecture (the one mlx_whisper ports) and is cross-checked in tests/test_oracle_cpu.py against
transformers 5.5.0 `WhisperModel` (modeling_whisper.py:541-648 encoder, :650-798 decoder) built
from an explicit WhisperConfig.  PARITY UNPINNED for this piece: the reference's own tests hold no
known-answer vectors for mel twin / encoder / decoder; parity is anchored on the
reference's call sites and in-tree loop:
    /root/reference/mlx_whisper_batch_decoder.py:267-303  BatchGreedyDecoder.update
    /root/reference/mlx_whisper_batch_decoder.py:317-384  BatchDecodingTask._main_loop_batch
    /root/reference/mlx_whisper_batch_decoder.py:386-468  run (EOT trimming, avg_logprob)
    /root/reference/whisperx/backends/mlx_lightning.py:187-196  DecodingOptions(temperature=0)
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """OpenAI Whisper encoder positions (HF modeling_whisper.py:55-62)."""
    assert channels % 2 == 0
    log_inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-log_inc * torch.arange(channels // 2, dtype=torch.float32))
    t = torch.arange(length, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.cat([t.sin(), t.cos()], dim=1)


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def _mha(q, k, v, n_head, causal_offset: Optional[int] = None):
    """q [B,Tq,d], k/v [B,Tk,d]; softmax(q k^T / sqrt(dh)) v.  causal_offset = absolute position of
    q[0] for causal masking over absolute key positions (None = no mask)."""
    B, Tq, d = q.shape
    Tk = k.shape[1]
    dh = d // n_head
    qh = q.view(B, Tq, n_head, dh).transpose(1, 2)
    kh = k.view(B, Tk, n_head, dh).transpose(1, 2)
    vh = v.view(B, Tk, n_head, dh).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * (dh ** -0.5)
    if causal_offset is not None:
        qi = torch.arange(Tq)[:, None] + causal_offset
        ki = torch.arange(Tk)[None, :]
        s = s.masked_fill(ki > qi, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B, Tq, d)


def _lin(x, w: Dict[str, torch.Tensor], name: str):
    return F.linear(x, w[name + ".weight"], w.get(name + ".bias"))


def encoder_forward(w: Dict[str, torch.Tensor], dims, mel: torch.Tensor, return_layers: bool = False):
    """mel f32 [B, n_mels, 3000] -> [B, 1500, d].  conv1 k3 p1, conv2 k3 s2 p1, GELU(erf) after each,
    + sinusoid positions, pre-LN blocks, ln_post."""
    x = F.gelu(F.conv1d(mel, w["encoder.conv1.weight"], w["encoder.conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, w["encoder.conv2.weight"], w["encoder.conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1) + w["encoder.positional_embedding"][None]
    layers = [x]
    for i in range(dims["n_audio_layer"]):
        p = f"encoder.blocks.{i}"
        h = _ln(x, w[p + ".attn_ln.weight"], w[p + ".attn_ln.bias"])
        a = _mha(_lin(h, w, p + ".attn.query"), _lin(h, w, p + ".attn.key"), _lin(h, w, p + ".attn.value"),
                 dims["n_audio_head"])
        x = x + _lin(a, w, p + ".attn.out")
        h = _ln(x, w[p + ".mlp_ln.weight"], w[p + ".mlp_ln.bias"])
        x = x + _lin(F.gelu(_lin(h, w, p + ".mlp.0")), w, p + ".mlp.2")
        layers.append(x)
    x = _ln(x, w["encoder.ln_post.weight"], w["encoder.ln_post.bias"])
    return (x, layers) if return_layers else x


class DecoderCache:
    """Self-attention K/V grow by one position per call; cross K/V are computed once."""

    def __init__(self, w, dims, enc_out: torch.Tensor):
        self.self_k: List[Optional[torch.Tensor]] = [None] * dims["n_text_layer"]
        self.self_v: List[Optional[torch.Tensor]] = [None] * dims["n_text_layer"]
        self.cross_k, self.cross_v = [], []
        for i in range(dims["n_text_layer"]):
            p = f"decoder.blocks.{i}.cross_attn"
            self.cross_k.append(_lin(enc_out, w, p + ".key"))
            self.cross_v.append(_lin(enc_out, w, p + ".value"))
        self.pos = 0


def decoder_forward(w, dims, tokens: torch.Tensor, cache: DecoderCache, positions=None, cross_qk=None) -> torch.Tensor:
    """tokens int64 [B, n] appended at positions cache.pos.. ; returns f32 logits [B, n, V]
    (tied output head, HF modeling_whisper.py:971).  `positions` (indices into the n new tokens) limits the vocabulary
    projection to those positions: logits [B, len(positions), V].  `cross_qk` = dict(heads=[(layer, head), ...], out=[]):
    the pre-softmax cross-attention scores q k^T / 8 of those heads are appended to out as [B, n_heads, n, 1500] (what the
    reference collects per forward, /root/reference/mlx_whisper_optimized_final.py:74-96)."""
    qk_here = {}
    B, n = tokens.shape
    pos0 = cache.pos
    x = w["decoder.token_embedding.weight"][tokens] + w["decoder.positional_embedding"][pos0:pos0 + n][None]
    for i in range(dims["n_text_layer"]):
        p = f"decoder.blocks.{i}"
        h = _ln(x, w[p + ".attn_ln.weight"], w[p + ".attn_ln.bias"])
        k_new, v_new = _lin(h, w, p + ".attn.key"), _lin(h, w, p + ".attn.value")
        cache.self_k[i] = k_new if cache.self_k[i] is None else torch.cat([cache.self_k[i], k_new], 1)
        cache.self_v[i] = v_new if cache.self_v[i] is None else torch.cat([cache.self_v[i], v_new], 1)
        a = _mha(_lin(h, w, p + ".attn.query"), cache.self_k[i], cache.self_v[i], dims["n_text_head"],
                 causal_offset=pos0)
        x = x + _lin(a, w, p + ".attn.out")
        h = _ln(x, w[p + ".cross_attn_ln.weight"], w[p + ".cross_attn_ln.bias"])
        if cross_qk is not None and any(l == i for l, _ in cross_qk["heads"]):
            dh = x.shape[-1] // dims["n_text_head"]
            qh = _lin(h, w, p + ".cross_attn.query").view(B, n, dims["n_text_head"], dh).transpose(1, 2)
            kh = cache.cross_k[i].view(B, -1, dims["n_text_head"], dh).transpose(1, 2)
            s_all = (qh @ kh.transpose(-1, -2)) * (dh ** -0.5)
            for l, hd in cross_qk["heads"]:
                if l == i:
                    qk_here[(l, hd)] = s_all[:, hd]
        a = _mha(_lin(h, w, p + ".cross_attn.query"), cache.cross_k[i], cache.cross_v[i], dims["n_text_head"])
        x = x + _lin(a, w, p + ".cross_attn.out")
        h = _ln(x, w[p + ".mlp_ln.weight"], w[p + ".mlp_ln.bias"])
        x = x + _lin(F.gelu(_lin(h, w, p + ".mlp.0")), w, p + ".mlp.2")
    cache.pos += n
    if cross_qk is not None:
        cross_qk["out"].append(torch.stack([qk_here[(l, hd)] for l, hd in cross_qk["heads"]], 1))
    if positions is not None:
        x = x[:, list(positions)]
    x = _ln(x, w["decoder.ln.weight"], w["decoder.ln.bias"])
    return (x @ w["decoder.token_embedding.weight"].t()).float()


def apply_timestamp_rules(logits: torch.Tensor, sampled: torch.Tensor, eot: int, ts_begin: int, no_timestamps: int,
                          max_initial_timestamp_index: Optional[int] = 50) -> torch.Tensor:
    """ApplyTimestampRules of the decoding module the reference uses (mlx_whisper.decoding, a port of OpenAI
    whisper/decoding.py; un-vendored, restated from the published algorithm, SURVEY A.3 filter 3).  logits f32 [B, V]
    (modified in place and returned), sampled int64 [B, n] = the tokens sampled so far (prompt excluded).  The last clause
    is the one the reference patches to be batch-safe, /root/reference/mlx_ultra_optimized_batch.py:38-71 (keepdims)."""
    B, n = sampled.shape
    if no_timestamps is not None and no_timestamps >= 0:
        logits[:, no_timestamps] = -float("inf")
    for k in range(B):
        seq = sampled[k].tolist()
        last_was_ts = len(seq) >= 1 and seq[-1] >= ts_begin
        pen_was_ts = len(seq) < 2 or seq[-2] >= ts_begin
        if last_was_ts:
            if pen_was_ts:
                logits[k, ts_begin:] = -float("inf")
            else:
                logits[k, :eot] = -float("inf")
        ts = [t for t in seq if t >= ts_begin]
        if ts:
            last = ts[-1] if (last_was_ts and not pen_was_ts) else ts[-1] + 1
            logits[k, ts_begin:last] = -float("inf")
    if n == 0:
        logits[:, :ts_begin] = -float("inf")
        if max_initial_timestamp_index is not None:
            logits[:, ts_begin + max_initial_timestamp_index + 1:] = -float("inf")
    logprobs = logits - torch.logsumexp(logits, -1, keepdim=True)
    ts_lp = torch.logsumexp(logprobs[:, ts_begin:], -1, keepdim=True)
    max_text = logprobs[:, :ts_begin].max(-1, keepdim=True).values
    logits[:, :ts_begin] = torch.where(ts_lp > max_text, torch.full_like(logits[:, :ts_begin], -float("inf")), logits[:, :ts_begin])
    return logits


def greedy_update(last: torch.Tensor, logits: torch.Tensor, sum_lp: torch.Tensor, eot: int):
    """BatchGreedyDecoder.update, mlx_whisper_batch_decoder.py:267-303: argmax, logprob of the chosen token added unless the
    row's last token is EOT, rows at EOT keep emitting EOT.  Returns (next [B], completed [B], sum_logprob [B])."""
    nxt = logits.argmax(-1)
    lp = logits - torch.logsumexp(logits, -1, keepdim=True)
    cur = lp[torch.arange(logits.shape[0]), nxt]
    not_eot = last != eot
    sum_lp = sum_lp + torch.where(not_eot, cur, torch.zeros_like(cur))
    nxt = torch.where(last == eot, torch.full_like(nxt, eot), nxt)
    return nxt, nxt == eot, sum_lp


def greedy_loop(logits_fn, prompt: torch.Tensor, eot: int, no_speech: int, sample_len: int, n_ctx: int = 448):
    """BatchDecodingTask._main_loop_batch, mlx_whisper_batch_decoder.py:317-384, over `logits_fn(step, tokens, active)`
    -> f32 [B, V] (rows of inactive sequences are zeros there, :91-98).  Filters are the caller's business.  Note the
    in-tree loop takes no_speech_prob from the FIRST SAMPLING step's logits (:347-352)."""
    tokens = prompt.clone()
    B = tokens.shape[0]
    sum_lp = torch.zeros(B)
    logits = logits_fn(0, tokens, torch.ones(B, dtype=torch.bool))
    nxt, done, sum_lp = greedy_update(tokens[:, -1], logits, sum_lp, eot)
    tokens = torch.cat([tokens, nxt[:, None]], 1)
    nsp = torch.softmax(logits, -1)[:, no_speech] if no_speech >= 0 else torch.full((B,), float("nan"))
    steps = 1
    for i in range(1, sample_len):
        if bool(done.all()) or tokens.shape[-1] > n_ctx:
            break
        logits = logits_fn(i, tokens, ~done)
        steps += 1
        nxt, done, sum_lp = greedy_update(tokens[:, -1], logits, sum_lp, eot)
        tokens = torch.cat([tokens, nxt[:, None]], 1)
    return tokens, sum_lp, nsp, steps


def greedy_decode(w, dims, enc_out: torch.Tensor, prompt: List[int], eot: int, no_speech: int = -1,
                  sample_len: int = 224, suppress_blank: bool = False, blank_token: int = 220,
                  suppress_tokens=(), return_logits: bool = False, timestamp_rules: Optional[dict] = None):
    """Batched greedy loop, mlx_whisper_batch_decoder.py:317-384 + :267-303, with the SuppressBlank /
    SuppressTokens filters and, with `timestamp_rules` (a dict like the kernel binding takes), ApplyTimestampRules.

    Returns dict(tokens=[B][...] up to first EOT, sum_logprob[B], avg_logprob[B], no_speech_prob[B],
    all_tokens int64 [B, n_sampled], step_logits (optional, filtered f32 logits per step)).
    no_speech_prob follows upstream Whisper (softmax at the SOT position, unfiltered); the in-tree
    batch loop takes it from the filtered last-prompt logits (:347-352) - see DESIGN.md."""
    B = enc_out.shape[0]
    cache = DecoderCache(w, dims, enc_out)
    toks = torch.tensor(prompt, dtype=torch.long)[None].expand(B, -1).contiguous()
    logits_all = decoder_forward(w, dims, toks, cache)
    if no_speech >= 0:
        no_speech_prob = torch.softmax(logits_all[:, 0].float(), -1)[:, no_speech]
    else:
        no_speech_prob = torch.full((B,), float("nan"))
    logits = logits_all[:, -1].clone()
    sum_lp = torch.zeros(B)
    last = toks[:, -1]
    sampled, step_logits = [], []
    suppress = torch.tensor(list(suppress_tokens), dtype=torch.long)
    n_ctx = dims["n_text_ctx"]
    for i in range(sample_len):
        if i == 0 and suppress_blank:
            logits[:, blank_token] = -float("inf")
            logits[:, eot] = -float("inf")
        if len(suppress):
            logits[:, suppress] = -float("inf")
        if timestamp_rules:
            hist = torch.stack(sampled, 1) if sampled else torch.zeros((B, 0), dtype=torch.long)
            apply_timestamp_rules(logits, hist, eot, timestamp_rules["timestamp_begin"], timestamp_rules.get("no_timestamps", -1),
                                  timestamp_rules.get("max_initial_timestamp_index", 50))
        if return_logits:
            step_logits.append(logits.clone())
        nxt, _, sum_lp = greedy_update(last, logits, sum_lp, eot)
        sampled.append(nxt)
        last = nxt
        if bool((last == eot).all()):
            break
        if len(prompt) + len(sampled) > n_ctx:
            break
        if i + 1 < sample_len:
            logits = decoder_forward(w, dims, nxt[:, None], cache)[:, -1].clone()
    all_tokens = torch.stack(sampled, 1)
    out_tokens = []
    for b in range(B):
        row = all_tokens[b].tolist()
        if eot in row:
            row = row[:row.index(eot)]
        out_tokens.append(row)
    avg = [float(s) / (len(t) + 1) for s, t in zip(sum_lp.tolist(), out_tokens)]
    res = dict(tokens=out_tokens, sum_logprob=sum_lp, avg_logprob=avg, no_speech_prob=no_speech_prob,
               all_tokens=all_tokens)
    if return_logits:
        res["step_logits"] = step_logits
    return res
