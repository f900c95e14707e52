"""Kernels on the N = 8192 FFT plan (Clifford d = 8192, bind / unbind n = 16384): python tools/bench_n8192.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "clifford-vae_b200")]
import torch
from clifford_b200 import _lib

dev = torch.device("cuda:0"); _lib.ensure_device(dev); lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
PEAK = 6548.8


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for dd in (16384,):
    N = (1 << 30) // (12 * dd)
    a = torch.randn(N, dd, device=dev); b = torch.randn(N, dd, device=dev); o = torch.empty(N, dd, device=dev)
    for mode, name in ((0, "bind"), (1, "unbind inv"), (2, "unbind deconv")):
        ms = timeit(lambda: lib.cvb_vsa_bind(a.data_ptr(), b.data_ptr(), o.data_ptr(), N, N, N, dd, mode, st))
        gb = N * 12 * dd / (ms * 1e-3) / 1e9
        print(f"{name:14s} d={dd} N={N} {ms:8.3f} ms {gb:7.1f} GB/s {100*gb/PEAK:5.1f}%")
    del a, b, o
B, d = 16384, 8192
loc = torch.randn(B, d, device=dev); kap = torch.rand(B, device=dev) * 9.87 + 0.13
z = torch.empty(B, 2 * d, device=dev); kl = torch.empty(B, device=dev); tps = torch.empty(B, d, device=dev)
ms = timeit(lambda: lib.cvb_clifford_ps_rsample(loc.data_ptr(), kap.data_ptr(), 1, 0, B, None, None, 7, 0, z.data_ptr(), None, None, kl.data_ptr(), None, B, d, st))
gb = B * (12 * d + 8) / (ms * 1e-3) / 1e9
print(f"clifford fwd rng  d={d} B={B} {ms:8.3f} ms {gb:7.1f} GB/s {100*gb/PEAK:5.1f}%")
lib.cvb_clifford_ps_rsample(loc.data_ptr(), kap.data_ptr(), 1, 0, B, None, None, 7, 0, z.data_ptr(), tps.data_ptr(), None, None, None, B, d, st)
gz = torch.randn(B, 2 * d, device=dev); dloc = torch.empty(B, d, device=dev); dk = torch.empty(B, device=dev)
ms = timeit(lambda: lib.cvb_clifford_ps_rsample_backward(gz.data_ptr(), loc.data_ptr(), kap.data_ptr(), 1, 0, B, None, None, tps.data_ptr(), dloc.data_ptr(), dk.data_ptr(), B, d, st))
gb = B * (20 * d + 8) / (ms * 1e-3) / 1e9
print(f"clifford bwd rng  d={d} B={B} {ms:8.3f} ms {gb:7.1f} GB/s {100*gb/PEAK:5.1f}%")
lp = torch.empty(B, device=dev)
ms = timeit(lambda: lib.cvb_clifford_ps_log_prob(z.data_ptr(), loc.data_ptr(), kap.data_ptr(), 1, 0, B, lp.data_ptr(), None, None, None, B, d, st))
gb = B * (12 * d + 8) / (ms * 1e-3) / 1e9
print(f"clifford log_prob d={d} B={B} {ms:8.3f} ms {gb:7.1f} GB/s {100*gb/PEAK:5.1f}%")
