"""A deterministic stand-in for the wav2vec2 CTC acoustic model (which is library code upstream of
kernel K4): logits are a seeded function of the waveform length, so the reference's align() (run in
tests/golden/make_golden.py) and ours see identical emissions without shipping 94 M parameters."""
import torch

# torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H.get_labels()
LABELS = ('-', '|', 'E', 'T', 'A', 'O', 'N', 'I', 'H', 'S', 'R', 'D', 'L', 'U', 'M', 'W', 'C', 'F', 'G', 'Y',
          'P', 'B', 'V', 'K', "'", 'X', 'J', 'Q', 'Z')
DICTIONARY = {c.lower(): i for i, c in enumerate(LABELS)}
METADATA = {"language": "en", "dictionary": DICTIONARY, "type": "torchaudio"}


class FakeCTCModel(torch.nn.Module):
    def __init__(self, n_labels: int = len(LABELS), seed: int = 0, scale: float = 3.0):
        super().__init__()
        self.n_labels, self.seed, self.scale = n_labels, seed, scale

    def forward(self, waveform, lengths=None):
        n = waveform.shape[-1]
        frames = (n - 400) // 320 + 1
        g = torch.Generator().manual_seed(self.seed * 1000003 + n)
        logits = torch.randn(1, frames, self.n_labels, generator=g) * self.scale
        return logits.to(waveform.device), None


def synthetic_speech(seconds: float, seed: int = 1234, sr: int = 16000):
    """Deterministic speech-like test signal: AM-modulated harmonics + noise with
    near-silent gaps, peak-normalised to 0.8."""
    import numpy as np
    rng = np.random.RandomState(seed)
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    x = np.zeros(n)
    for f0, a in ((220.0, 1.0), (440.0, 0.6), (880.0, 0.3)):
        x += a * np.sin(2 * np.pi * f0 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 5.0 * t))
    x += rng.normal(0.0, 0.02, n)
    # near-silent gaps of 0.3-1.0 s every few seconds
    pos = 0.0
    while pos < seconds:
        pos += rng.uniform(2.0, 6.0)
        gap = rng.uniform(0.3, 1.0)
        a, b = int(pos * sr), int(min(seconds, pos + gap) * sr)
        x[a:b] *= 1e-3
        pos += gap
    x = 0.8 * x / np.abs(x).max()
    return x.astype(np.float32)
