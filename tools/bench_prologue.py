"""Cost of the fp64 entropy / KL prologue inside the fused sampler at the headline shape: the same launch with and without
the row-scalar outputs (CUDA events, inputs rotated over > L2 of buffers)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "clifford-vae_b200")]
import torch
from clifford_b200 import _lib
dev = torch.device("cuda:0"); _lib.ensure_device(dev); lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
for B, d in ((4096, 2048), (65536, 2048), (4096, 512)):
    nset = max(2, int(300e6 // (B * d * 12)) + 1)
    locs = [torch.randn(B, d, device=dev) for _ in range(nset)]
    zs = [torch.empty(B, 2 * d, device=dev) for _ in range(nset)]
    kap = torch.rand(B, device=dev) * 9.97 + 0.03; kl = torch.empty(B, device=dev)
    def run(with_kl, reps=40):
        for i in range(5):
            lib.cvb_clifford_ps_rsample(locs[i % nset].data_ptr(), kap.data_ptr(), 1, 0, B, None, None, 7, i, zs[i % nset].data_ptr(), None, None, kl.data_ptr() if with_kl else None, None, B, d, st)
        torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
        for i in range(reps):
            lib.cvb_clifford_ps_rsample(locs[i % nset].data_ptr(), kap.data_ptr(), 1, 0, B, None, None, 7, i, zs[i % nset].data_ptr(), None, None, kl.data_ptr() if with_kl else None, None, B, d, st)
        b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps * 1e3
    t1 = min(run(True) for _ in range(3)); t0 = min(run(False) for _ in range(3))
    print(f"B={B:6d} d={d:5d}  with KL {t1:8.2f} us   without {t0:8.2f} us   prologue {t1 - t0:6.2f} us ({100 * (t1 - t0) / t1:.1f} %)")
    del locs, zs
