import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "clifford-vae_b200")]
import torch
from clifford_b200 import _lib
dev = torch.device("cuda:0"); _lib.ensure_device(dev); lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
B, d = 4096, 2048
loc = torch.randn(B, d, device=dev); kap = torch.rand(B, device=dev) * 9.87 + 0.13
z = torch.empty(B, 2 * d, device=dev); tps = torch.empty(B, d, device=dev)
gz = torch.randn(B, 2 * d, device=dev); dloc = torch.empty(B, d, device=dev); dk = torch.empty(B, device=dev); lp = torch.empty(B, device=dev)
for _ in range(3):
    lib.cvb_clifford_ps_rsample(loc.data_ptr(), kap.data_ptr(), 1, 0, B, None, None, 7, 0, z.data_ptr(), tps.data_ptr(), None, None, None, B, d, st)
    lib.cvb_clifford_ps_rsample_backward(gz.data_ptr(), loc.data_ptr(), kap.data_ptr(), 1, 0, B, None, None, tps.data_ptr(), dloc.data_ptr(), dk.data_ptr(), B, d, st)
    lib.cvb_clifford_ps_log_prob(z.data_ptr(), loc.data_ptr(), kap.data_ptr(), 1, 0, B, lp.data_ptr(), None, None, None, B, d, st)
torch.cuda.synchronize()
print("ok")
