import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspectrogram_b200 import engine
nfft, ncol = 4096, 40000
iq = torch.empty(nfft * ncol + 8, dtype=torch.complex64, device="cuda")
torch.view_as_real(iq).normal_(0, 1e-2)
starts = torch.arange(ncol, device="cuda", dtype=torch.int64) * nfft
plan = engine.StiPlan(nfft)
out = torch.empty((1, ncol, nfft), dtype=torch.float32, device="cuda")
for _ in range(3):
    plan.run(iq, starts, 1, nfft, want_lin=False, want_db=True, out_db=out)
torch.cuda.synchronize()
print("ok")
