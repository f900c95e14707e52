"""bind / unbind(inv) / unbind(deconv) per size: python tools/bench_unbind.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "clifford-vae_b200")]
import torch
from clifford_b200 import _lib
dev = torch.device("cuda:0"); _lib.ensure_device(dev); lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
PEAK = 6548.8
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for dd in (1024, 4096, 16384):
    N = (1 << 30) // (12 * dd)
    a = torch.randn(N, dd, device=dev); b = torch.randn(N, dd, device=dev); o = torch.empty(N, dd, device=dev)
    for mode, name in ((0, "bind"), (1, "unbind inv"), (2, "unbind deconv")):
        ms = timeit(lambda: lib.cvb_vsa_bind(a.data_ptr(), b.data_ptr(), o.data_ptr(), N, N, N, dd, mode, st))
        gb = N * 12 * dd / (ms * 1e-3) / 1e9
        print(f"{name:14s} d={dd:6d} N={N:6d} {ms:8.3f} ms {gb:7.1f} GB/s {100*gb/PEAK:5.1f}%")
    del a, b, o
