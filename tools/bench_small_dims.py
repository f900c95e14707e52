import os, sys
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "clifford-vae_b200")]
import torch
from clifford_b200 import _lib
dev = torch.device("cuda:0"); _lib.ensure_device(dev); lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps
for d in (16, 32, 64, 128, 256):
    B = (1<<28)//(12*d)
    loc=torch.randn(B,d,device=dev); kap=torch.rand(B,device=dev)*9+0.1; z=torch.empty(B,2*d,device=dev); kl=torch.empty(B,device=dev)
    ms=timeit(lambda: lib.cvb_clifford_ps_rsample(loc.data_ptr(),kap.data_ptr(),1,0,B,None,None,7,0,z.data_ptr(),None,None,kl.data_ptr(),None,B,d,st))
    print(f"clifford fwd rng d={d:4d} B={B:8d} {ms:7.3f} ms {B*(12*d+8)/ms/1e6:8.1f} GB/s")
    del loc,z
for d in (32, 64, 128, 256, 512):
    N=(1<<28)//(12*d)
    a=torch.randn(N,d,device=dev); b=torch.randn(N,d,device=dev); o=torch.empty(N,d,device=dev)
    ms=timeit(lambda: lib.cvb_vsa_bind(a.data_ptr(),b.data_ptr(),o.data_ptr(),N,N,N,d,0,st))
    print(f"bind d={d:4d} N={N:8d} {ms:7.3f} ms {N*12*d/ms/1e6:8.1f} GB/s")
    del a,b,o
for d in (20, 100, 513):
    N=4096
    a=torch.randn(N,d,device=dev); b=torch.randn(N,d,device=dev); o=torch.empty(N,d,device=dev)
    ms=timeit(lambda: lib.cvb_vsa_bind(a.data_ptr(),b.data_ptr(),o.data_ptr(),N,N,N,d,0,st))
    print(f"bind generic d={d:4d} N={N:8d} {ms:7.3f} ms {N/ms*1e3:.3e} vec/s")
