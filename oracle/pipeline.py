"""
ORACLE — test infrastructure only (bench.py's cpu_baseline / --impl reference legs and tests).

The reference's CPU path for the hot path, end to end on host cores: log-mel (oracle/mel.py, a
restatement of whisperx/audio.py) -> Whisper encoder + batched greedy decode (oracle/whisper.py:
torch CPU fp32; stands in for the faster-whisper / CTranslate2 CPU backend BASELINE.json names, which
is a 15-line stub in the reference and not installable here) -> CTC trellis + beam-2 backtrack
(oracle/ctc.py, a restatement of whisperx/alignment.py).
"""
import time
from typing import Dict, List

import numpy as np
import torch

from . import ctc as octc
from . import mel as omel
from . import whisper as ow


def cpu_hot_path(chunks: List[np.ndarray], dims: Dict[str, int], w: Dict[str, torch.Tensor], prompt: List[int], eot: int,
                 no_speech: int, sample_len: int, emissions: List[np.ndarray], token_lists: List[List[int]],
                 suppress_blank: bool = True, blank_token: int = 220):
    """Returns (results, timings_seconds)."""
    t = {}
    t0 = time.perf_counter()
    mel = torch.from_numpy(omel.log_mel_chunks(chunks, dims["n_mels"]))
    t["mel"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    with torch.no_grad():
        enc = ow.encoder_forward(w, dims, mel)
        t["encoder"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        dec = ow.greedy_decode(w, dims, enc, prompt, eot, no_speech=no_speech, sample_len=sample_len,
                               suppress_blank=suppress_blank, blank_token=blank_token)
    t["decoder"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    paths = []
    for e, toks in zip(emissions, token_lists):
        em = torch.log_softmax(torch.from_numpy(e), -1).numpy()
        tr = octc.get_trellis(em, toks, 0)
        paths.append(octc.backtrack_beam(tr, em, toks, 0, beam_width=2))
    t["ctc"] = time.perf_counter() - t0
    return dict(decode=dec, paths=paths), t
