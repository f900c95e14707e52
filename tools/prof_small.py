"""ncu target: the short-row kernels at the reference's default MNIST latent (d = 20 circles, n = 40; bind at 40 and 21).
python tools/prof_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "clifford-vae_b200")]
import torch
from clifford_b200 import _lib
dev = torch.device("cuda:0"); _lib.ensure_device(dev); lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
B, d = 1 << 16, 20
loc = torch.randn(B, d, device=dev); kap = torch.rand(B, device=dev) * 9.87 + 0.13
z = torch.empty(B, 2 * d, device=dev); kl = torch.empty(B, device=dev); lp = torch.empty(B, device=dev)
gz = torch.randn(B, 2 * d, device=dev); tps = torch.rand(B, d, device=dev) * 0.98 + 0.01; dl = torch.empty(B, d, device=dev); dk = torch.empty(B, device=dev)
for i in range(2):
    assert lib.cvb_clifford_ps_rsample(loc.data_ptr(), kap.data_ptr(), 1, 0, B, None, None, 7, i, z.data_ptr(), None, None, kl.data_ptr(), None, B, d, st) == 0
    assert lib.cvb_clifford_ps_rsample_backward(gz.data_ptr(), loc.data_ptr(), kap.data_ptr(), 1, 0, B, None, None, tps.data_ptr(), dl.data_ptr(), dk.data_ptr(), B, d, st) == 0
    assert lib.cvb_clifford_ps_log_prob(z.data_ptr(), loc.data_ptr(), kap.data_ptr(), 1, 0, B, lp.data_ptr(), None, None, None, B, d, st) == 0
for dd in (40, 21):
    N = 1 << 17
    a = torch.randn(N, dd, device=dev); b = torch.randn(N, dd, device=dev); o = torch.empty(N, dd, device=dev)
    for mode in (0, 2):
        assert lib.cvb_vsa_bind(a.data_ptr(), b.data_ptr(), o.data_ptr(), N, N, N, dd, mode, st) == 0
torch.cuda.synchronize()
print("ok")
