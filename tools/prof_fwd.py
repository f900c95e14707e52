"""ncu target: a few Clifford forward (device RNG, row-scalar kappa) launches.  python tools/prof_fwd.py [B] [d]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "clifford-vae_b200")]
import torch
from clifford_b200 import _lib
dev = torch.device("cuda:0"); _lib.ensure_device(dev); lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
loc = torch.randn(B, d, device=dev); kap = torch.rand(B, device=dev) * 9.87 + 0.13
z = torch.empty(B, 2 * d, device=dev); kl = torch.empty(B, device=dev)
for i in range(4):
    lib.cvb_clifford_ps_rsample(loc.data_ptr(), kap.data_ptr(), 1, 0, B, None, None, 7, i, z.data_ptr(), None, None, kl.data_ptr(), None, B, d, st)
torch.cuda.synchronize()
print("ok", B, d)
