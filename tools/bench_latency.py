"""Latency-sized configurations (BASELINE configs 1 and 2): back-to-back launches timed with CUDA events, with and without
the fused row scalars, next to the launch floor of this box (an empty-work kernel of the same library).
python tools/bench_latency.py             (under `ncu --metrics gpu__time_duration.sum` it gives the pure kernel durations)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "clifford-vae_b200")]
import torch
from clifford_b200 import _lib
dev = torch.device("cuda:0"); _lib.ensure_device(dev); raw = _lib.load(); st = torch.cuda.current_stream().cuda_stream


class _Checked:
    def __getattr__(self, name):
        fn = getattr(raw, name)
        def call(*a):
            _lib.check(fn(*a), name)
        return call


lib = _Checked()
REPS = int(os.environ.get("REPS", "200"))


def timeit(fn, reps=REPS):
    for _ in range(10): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps * 1e3


buf = torch.empty(64, dtype=torch.int32, device=dev)
print(f"launch floor (philox_fill of 4 vec4)            {timeit(lambda: lib.cvb_philox_fill(buf.data_ptr(), 4, 1, 0, st)):7.2f} us")
for B1, d1 in ((128, 512), (128, 64), (1, 512), (1024, 512), (128, 2048)):
    loc1 = torch.randn(B1, d1, device=dev); kap1 = torch.rand(B1, device=dev) * 9.97 + 0.03
    z1 = torch.empty(B1, 2 * d1, device=dev); kl1 = torch.empty(B1, device=dev)
    t_kl = timeit(lambda: lib.cvb_clifford_ps_rsample(loc1.data_ptr(), kap1.data_ptr(), 1, 0, B1, None, None, 7, 0, z1.data_ptr(), None, None, kl1.data_ptr(), None, B1, d1, st))
    t_no = timeit(lambda: lib.cvb_clifford_ps_rsample(loc1.data_ptr(), kap1.data_ptr(), 1, 0, B1, None, None, 7, 0, z1.data_ptr(), None, None, None, None, B1, d1, st))
    print(f"clifford rsample B={B1:5d} d={d1:5d}   with KL {t_kl:7.2f} us   without {t_no:7.2f} us")
for B2, D2 in ((1024, 513), (1024, 512), (128, 513), (1, 513)):
    loc = torch.nn.functional.normalize(torch.randn(B2, D2, device=dev), dim=-1); kap2 = torch.rand(B2, device=dev) * 9.2 + 0.8
    z = torch.empty(B2, D2, device=dev); save = torch.empty(B2, 2, device=dev); ent = torch.empty(B2, device=dev); kl = torch.empty(B2, device=dev)
    t1 = timeit(lambda: lib.cvb_powerspherical_rsample_kl(loc.data_ptr(), kap2.data_ptr(), B2, None, None, 3, 0, z.data_ptr(), save.data_ptr(), ent.data_ptr(), kl.data_ptr(), None, B2, D2, st))
    t0 = timeit(lambda: lib.cvb_powerspherical_rsample(loc.data_ptr(), kap2.data_ptr(), B2, None, None, 3, 0, z.data_ptr(), save.data_ptr(), B2, D2, st))
    t3 = timeit(lambda: lib.cvb_vmf_rsample_kl(loc.data_ptr(), kap2.data_ptr(), B2, None, None, 0, None, 3, 0, z.data_ptr(), save.data_ptr(), ent.data_ptr(), kl.data_ptr(), None, None, None, B2, D2, st))
    t2 = timeit(lambda: lib.cvb_vmf_rsample(loc.data_ptr(), kap2.data_ptr(), B2, None, None, 0, None, 3, 0, z.data_ptr(), save.data_ptr(), B2, D2, st))
    print(f"sphere  rsample B={B2:5d} D={D2:5d}   PS with KL {t1:7.2f} us  without {t0:7.2f} us   vMF with KL {t3:7.2f} us  without {t2:7.2f} us")
