"""One K4 launch for an ncu capture: python tools/probe/ctc_one.py [N] [n_seg] [mode: 0 backtrack, 1 beam2, 2 trellis only]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'whisperx-mlx_b200')): sys.path.insert(0, p)
from whisperx._native import get_context
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1040
n_seg = int(sys.argv[2]) if len(sys.argv) > 2 else 15
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = get_context(0); T, V = 1499, 29
rng = np.random.RandomState(1)
em = torch.log_softmax(torch.randn(n_seg * T, V, device='cuda') * 3.0, -1)
tok = torch.from_numpy(rng.randint(1, V, size=n_seg * N).astype(np.int32)).cuda()
ctx.ctc_align(em, np.arange(n_seg + 1) * T, tok, np.arange(n_seg + 1) * N, 0, mode)
torch.cuda.synchronize()
