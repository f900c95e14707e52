"""ncu target: a few bind launches at one size.  python tools/prof_bind.py [d]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "clifford-vae_b200")]
import torch
from clifford_b200 import _lib
dev = torch.device("cuda:0"); _lib.ensure_device(dev); lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
dd = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = (1 << 29) // (12 * dd)
a = torch.randn(N, dd, device=dev); b = torch.randn(N, dd, device=dev); o = torch.empty(N, dd, device=dev)
for _ in range(4):
    lib.cvb_vsa_bind(a.data_ptr(), b.data_ptr(), o.data_ptr(), N, N, N, dd, 0, st)
torch.cuda.synchronize()
print("ok", N)
