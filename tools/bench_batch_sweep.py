"""Fixed overhead vs per-row cost of the two headline kernels: python tools/bench_batch_sweep.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "clifford-vae_b200")]
import torch
from clifford_b200 import _lib

dev = torch.device("cuda:0")
_lib.ensure_device(dev)
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, reps=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


d = 2048
n = 2 * d
for B in (148, 592, 1184, 2048, 4096, 8192, 16384, 32768):
    loc = torch.randn(B, d, device=dev); kap = torch.rand(B, device=dev) * 9.87 + 0.13
    z = torch.empty(B, n, device=dev); kl = torch.empty(B, device=dev)
    roles = torch.randn(B, n, device=dev) / n ** 0.5; out = torch.empty(B, n, device=dev)
    t_f = timeit(lambda: lib.cvb_clifford_ps_rsample(loc.data_ptr(), kap.data_ptr(), 1, 0, B, None, None, 7, 0, z.data_ptr(), None, None,
                                                      kl.data_ptr(), None, B, d, st))
    t_b = timeit(lambda: lib.cvb_vsa_bind(z.data_ptr(), roles.data_ptr(), out.data_ptr(), B, B, B, n, 0, st))
    print(f"B={B:6d}  rsample+KL {t_f*1e3:8.1f} us ({t_f*1e3/B*1e3:6.2f} ns/row)   bind {t_b*1e3:8.1f} us ({t_b*1e3/B*1e3:6.2f} ns/row)")
