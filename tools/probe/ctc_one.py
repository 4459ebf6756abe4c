import os, sys
import numpy as np, torch
ROOT='/root/repo'
for p in (ROOT, os.path.join(ROOT,'whisperx-mlx_b200')): sys.path.insert(0,p)
from whisperx._native import CTC_BEAM2, get_context
ctx=get_context(0); T,V,N,n_seg=1499,29,1040,15
rng=np.random.RandomState(1)
em=torch.log_softmax(torch.randn(n_seg*T,V,device='cuda')*3.0,-1)
tok=torch.from_numpy(rng.randint(1,V,size=n_seg*N).astype(np.int32)).cuda()
ctx.ctc_align(em,np.arange(n_seg+1)*T,tok,np.arange(n_seg+1)*N,0,CTC_BEAM2)
torch.cuda.synchronize()
