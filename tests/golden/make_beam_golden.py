"""
Golden paths of the reference's backtrack_beam for beam widths other than the 2 that align() uses
(/root/reference/whisperx/alignment.py:500-579, default beam_width = 5), on the emissions / tokens already stored in
ctc_golden.npz.  Run:  python tests/golden/make_beam_golden.py  ->  tests/golden/beam_width_golden.npz
"""
import os

import numpy as np
import torch

import make_golden as mg  # installs the nltk stub and imports the reference (whisperx.alignment as ref_align)

HERE = os.path.dirname(os.path.abspath(__file__))
g = np.load(os.path.join(HERE, "ctc_golden.npz"))
out = {"widths": np.array([1, 3, 5, 8], np.int32)}
for name in g["names"].tolist():
    em = torch.from_numpy(g[f"{name}_emission"])
    tokens = g[f"{name}_tokens"].tolist()
    blank = int(g[f"{name}_blank"])
    trellis = mg.ref_align.get_trellis(em, tokens, blank)
    for w in out["widths"].tolist():
        p = mg.ref_align.backtrack_beam(trellis, em, tokens, blank, beam_width=w)
        out[f"{name}_w{w}_ok"] = np.array(0 if p is None else 1)
        if p is not None:
            out[f"{name}_w{w}_tok"] = np.array([q.token_index for q in p], np.int32)
            out[f"{name}_w{w}_score"] = np.array([q.score for q in p], np.float64)
    print(name, [int(out[f"{name}_w{w}_ok"]) for w in out["widths"].tolist()])
np.savez_compressed(os.path.join(HERE, "beam_width_golden.npz"), **out)
