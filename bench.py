#!/usr/bin/env python
"""bench.py -- latent hot-path throughput on B200 (contract: one JSON line on stdout).

Workload (BASELINE.json configs[2] latent shape, the one the north_star's 70 %-of-HBM target is
quoted on): per step and per GPU, B = 4096 rows of the Clifford-torus latent at d = 2048
  1. fused Clifford power-spherical rsample + entropy/KL (device Philox RNG, one kappa per row)
     -> z (B, 4096)                                   [reference dists/clifford.py:295-327]
  2. vsa.bind(z, roles) with per-row roles (B, 4096)  [reference utils/vsa.py:43-46]
`value` = samples/s with inputs resident in HBM (two rotating buffer sets, 2 x 224 MB > 126 MB L2);
`e2e` = the same step through the reference-named Python API with HOST (pinned) inputs/outputs.
`--impl reference` times the reference's OWN classes on the host cores (staged copy under oracle/_ref, see
oracle/stage_reference.py; falls back to the oracle port only when no copy of the reference is available).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "clifford-vae_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

B_ROWS, D_LAT = 4096, 2048          # per GPU
N_VEC = 2 * D_LAT
METRIC = "latent rsample+KL+bind samples/s"
UNIT = "samples/s"
WORKLOAD = ("C3 latent: Clifford-PS fused rsample+KL (device RNG, row-scalar kappa) d=2048 -> z (4096 x 4096), "
            "then vsa.bind(z, roles) n=4096; B=4096 rows per GPU per step")
# algorithmic bytes per sample: rsample+KL 12d+8, bind 12n
BYTES_RSAMPLE = 12 * D_LAT + 8
BYTES_BIND = 12 * N_VEC


def load_traffic(kernel_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    ncu --set full capture of this bench (profiles/r01_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)
        for k, v in t.items():
            if kernel_name.startswith(k):
                return v["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own torch CPU path (oracle/_ref), else the oracle port
# --------------------------------------------------------------------------------------------------
REF_ROWS = 1024     # bounded sample of the 4096-row step (the reference runs at ~1e3 samples/s on 16 host threads)


def make_cpu_step(torch):
    """-> (step(loc, kap, roles) -> (bound, kl), kind, description).  kind "reference": the reference's classes,
    imported unmodified (dists/clifford.py:295-327 rsample + kl_divergence, utils/vsa.py:43-46 bind), constructed the
    way its conv VAE does (cnn/models.py:226-231: concentration expanded to loc's shape).  kind "port": the oracle
    restatement, used only when neither /root/reference nor oracle/_ref exists."""
    from oracle import reference_loader as RL
    ref = RL.load()
    if ref is not None:
        C, V = ref.clifford, ref.vsa
        prior = C.CliffordTorusUniform(D_LAT)

        def step(loc, kap, roles):
            q = C.CliffordPowerSphericalDistribution(loc, kap.expand_as(loc))
            z = q.rsample()
            kl = torch.distributions.kl.kl_divergence(q, prior)
            return V.bind(z, roles), kl
        return step, "reference", f"the reference's own classes ({ref.kind} copy), torch {torch.__version__} CPU"

    from oracle import latent_oracle as O

    def step(loc, kap, roles):
        rows, d = loc.shape
        alpha = 0.5 + kap + 1e-7
        tprime = torch.distributions.Beta(alpha.expand(rows, d), torch.full((rows, d), 0.5)).sample()
        g = torch.randn(rows, d)
        z = O.clifford_ps_rsample(loc, kap, tprime, g)
        kl = O.clifford_ps_kl(kap.expand(rows, d))
        return O.bind(z, roles), kl
    return step, "port", f"oracle port of the reference's torch CPU path, torch {torch.__version__} CPU"


def cpu_inputs(torch, rows):
    torch.manual_seed(0)
    loc = torch.randn(rows, D_LAT)
    kap = torch.rand(rows, 1) * 9.87 + 0.13
    roles = torch.randn(rows, N_VEC) / math.sqrt(N_VEC)
    return loc, kap, roles


def cpu_timing(rows, reps):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind, desc = make_cpu_step(torch)
    loc, kap, roles = cpu_inputs(torch, rows)
    with torch.no_grad():
        step(loc, kap, roles)          # warm-up
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            step(loc, kap, roles)
            ts.append(time.perf_counter() - t0)
    return rows / statistics.median(ts), torch.get_num_threads(), ts, kind, desc


def bench_config():
    """`config` of both arms (identical keys and values: same workload, same shapes)."""
    return {"workload": WORKLOAD, "rows_per_gpu": B_ROWS, "d": D_LAT,
            "rng": "b200 arm: Philox4x32-10 on the device; reference arm: torch CPU generator",
            "l2": "b200 arm: inputs rotate over 2 buffer sets of 224 MB (> 126 MB L2)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind, desc = make_cpu_step(torch)
    rows = REF_ROWS
    loc, kap, roles = cpu_inputs(torch, rows)
    with torch.no_grad():
        for _ in range(args.warmup):
            step(loc, kap, roles)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(loc, kap, roles)
        dt = time.perf_counter() - t0
    val = rows * args.steps / dt
    sample = f"{rows} of the {B_ROWS} rows per step (same d={D_LAT}); {desc}"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from clifford_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.ensure_device(dev)
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    B, d, n = B_ROWS, D_LAT, N_VEC
    torch.manual_seed(1234 + rank)
    NSETS = 2
    sets = []
    for _ in range(NSETS):
        sets.append(dict(
            loc=torch.randn(B, d, device=dev), kap=(torch.rand(B, device=dev) * 9.87 + 0.13),
            roles=torch.randn(B, n, device=dev) / math.sqrt(n), z=torch.empty(B, n, device=dev),
            out=torch.empty(B, n, device=dev), kl=torch.empty(B, device=dev), ent=torch.empty(B, device=dev)))
    seed = 1234 + rank

    def step(i, ev=None):
        s = sets[i % NSETS]
        if ev:
            ev[0].record()
        rc = lib.cvb_clifford_ps_rsample(s["loc"].data_ptr(), s["kap"].data_ptr(), 1, 0, B, None, None, seed, i,
                                         s["z"].data_ptr(), None, s["ent"].data_ptr(), s["kl"].data_ptr(), None, B, d, st)
        if ev:
            ev[1].record()
        rc |= lib.cvb_vsa_bind(s["z"].data_ptr(), s["roles"].data_ptr(), s["out"].data_ptr(), B, B, B, n, 0, st)
        if ev:
            ev[2].record()
        if rc:
            raise RuntimeError(lib.cvb_last_error_string().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    l0 = _lib.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for i in range(args.steps):
        step(args.warmup + i, evs[i])
    t_end.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms_total = t_start.elapsed_time(t_end)
    ms_rs = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    ms_bind = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps

    # ---- e2e: reference-named Python API, host (pinned) buffers in and out ------------------------
    from dists.clifford import CliffordPowerSphericalDistribution, CliffordTorusUniform
    from utils import vsa
    h_loc = torch.randn(B, d).pin_memory()
    h_kap = (torch.rand(B, 1) * 9.87 + 0.13).pin_memory()
    h_roles = (torch.randn(B, n) / math.sqrt(n)).pin_memory()
    h_out = torch.empty(B, n).pin_memory()
    h_kl = torch.empty(B).pin_memory()
    prior = CliffordTorusUniform(d, device=dev)

    # Double-buffered pipeline: copy-in (H2D), compute, copy-out (D2H) on three streams, two device buffer
    # sets; every step still moves its own inputs from pinned host memory and its own results back.
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    dbuf = [dict(loc=torch.empty(B, d, device=dev), kap=torch.empty(B, 1, device=dev), roles=torch.empty(B, n, device=dev),
                 ev_in=torch.cuda.Event(), ev_cmp=torch.cuda.Event(), ev_out=torch.cuda.Event(), out=None, kl=None)
            for _ in range(2)]

    def e2e_step(i):
        b = dbuf[i % 2]
        with torch.no_grad():
            with torch.cuda.stream(s_in):
                s_in.wait_event(b["ev_cmp"])                  # the compute that last read this set is done
                b["loc"].copy_(h_loc, non_blocking=True)
                b["kap"].copy_(h_kap, non_blocking=True)
                b["roles"].copy_(h_roles, non_blocking=True)
                b["ev_in"].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(b["ev_in"])
                s_cmp.wait_event(b["ev_out"])                 # the previous outputs of this set were copied out
                q = CliffordPowerSphericalDistribution(b["loc"], b["kap"], validate_args=False)
                z = q.rsample()
                b["kl"] = torch.distributions.kl.kl_divergence(q, prior)
                b["out"] = vsa.bind(z, b["roles"])
                b["ev_cmp"].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(b["ev_cmp"])
                h_out.copy_(b["out"], non_blocking=True)
                h_kl.copy_(b["kl"], non_blocking=True)
                b["out"].record_stream(s_out)
                b["kl"].record_stream(s_out)
                b["ev_out"].record(s_out)

    def e2e_drain():
        s_in.synchronize(); s_cmp.synchronize(); s_out.synchronize()

    e2e_steps = max(3, min(args.steps, 20))
    for i in range(4):
        e2e_step(i)
    e2e_drain()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s_in.wait_event(e0); s_cmp.wait_event(e0); s_out.wait_event(e0)
    for i in range(e2e_steps):
        e2e_step(i)
    e2e_drain()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None

    vae = None if args.no_vae_step else vae_train_leg(torch, dev, world, rank, local)

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_total, ms_e2e, ms_rs, ms_bind], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e, ms_rs, ms_bind = (float(x) for x in t.tolist())

    if rank == 0:
        peak, peak_src = load_peaks()
        value = world * B * args.steps / (ms_total * 1e-3)
        e2e_val = world * B * e2e_steps / (ms_e2e * 1e-3)
        k_rs = {"name": "clifford_fwd_kernel<11,PsRng,rowk>", "ms": ms_rs, "bytes": B * BYTES_RSAMPLE}
        k_bd = {"name": "bind_v3_kernel<11,Mul,direct>", "ms": ms_bind, "bytes": B * BYTES_BIND}
        dom = k_rs if ms_rs >= ms_bind else k_bd
        achieved = dom["bytes"] / (dom["ms"] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(),
            "roofline": {"bound": "hbm", "kernel": dom["name"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": load_traffic(dom["name"]), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom["bytes"], "ms_per_launch": dom["ms"]},
            "kernels": {
                "rsample_kl": {"ms": ms_rs, "GBps": k_rs["bytes"] / (ms_rs * 1e-3) / 1e9,
                               "frac": k_rs["bytes"] / (ms_rs * 1e-3) / 1e9 / peak},
                "bind": {"ms": ms_bind, "GBps": k_bd["bytes"] / (ms_bind * 1e-3) / 1e9,
                         "frac": k_bd["bytes"] / (ms_bind * 1e-3) / 1e9 / peak}},
            "e2e": {"value": e2e_val, "unit": UNIT, "steps": e2e_steps,
                    "h2d_bytes_per_step": world * (h_loc.numel() + h_kap.numel() + h_roles.numel()) * 4,
                    "d2h_bytes_per_step": world * (h_out.numel() + h_kl.numel()) * 4,
                    "api": "dists.clifford.CliffordPowerSphericalDistribution.rsample + kl_divergence + utils.vsa.bind",
                    "pipeline": "double-buffered: H2D / compute / D2H on three streams"},
            "gpu_launches": int(launches), "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, ts, kind, desc = cpu_timing(REF_ROWS, 5)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{REF_ROWS} of the {B} rows (same d={d}), median of 5 passes "
                                              f"({sum(ts):.1f} s CPU); {desc}"}
        if vae is not None:
            line["vae_train_step"] = vae
        if world == 1:
            line["other_configs"] = extras(torch, lib, dev, st, peak, full=args.extras)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def vae_train_leg(torch, dev, world, rank, local, steps=8, warmup=3, compare_reference_dists=True):
    """Second half of BASELINE.json's metric: the C3 training step of the reference's OWN model -- `cnn.models.VAE(
    latent_dim=2048, in_channels=3, distribution="clifford", device, recon_loss_type="l1")` (cnn/models.py:134-315,
    18.4 M parameters), loaded unmodified from the staged copy under oracle/_ref with `dists.clifford` resolved to the
    drop-in classes -- run the way cnn/cifar10_train.py:62-77 runs it (zero_grad, forward, compute_loss, backward,
    clip_grad_norm_ 1.0, AdamW lr 3e-4 step; the per-step .item() logging left out), batch 4096 per GPU, synthetic
    U(-1,1) 3x32x32 inputs, DistributedDataParallel over NCCL when world > 1.  Every rank runs it; the time is the
    max over ranks.  At world == 1 the same step is also timed with the reference's own distribution classes on the
    GPU (eager PyTorch + cuFFT): the "reference on cuda" comparator of BASELINE.md.  Falls back to the stand-in conv
    VAE of examples/train_vae_ddp.py only when no copy of the reference is available."""
    import torch.distributed as dist
    import torch.nn as nn
    import importlib.util
    spec = importlib.util.spec_from_file_location("train_vae_ddp", os.path.join(ROOT, "examples", "train_vae_ddp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from oracle import reference_loader as RL
    ref = RL.load()
    batch = 4096
    torch.manual_seed(1234 + rank)

    def reference_model_step(models_mod):
        model = models_mod.VAE(latent_dim=D_LAT, in_channels=3, distribution="clifford", device=dev, recon_loss_type="l1")
        net = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
        opt = torch.optim.AdamW(net.parameters(), lr=3e-4)
        x = torch.rand(batch, 3, 32, 32, device=dev) * 2 - 1

        def step():
            opt.zero_grad()
            x_recon, q_z, p_z, _ = net(x)
            losses = model.compute_loss(x, x_recon, q_z, p_z, 1.0)
            losses["total_loss"].backward()
            torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
            opt.step()
            return losses["total_loss"]
        return step, sum(p.numel() for p in model.parameters())

    out = {"batch_per_gpu": batch, "n_gpus": world, "parallelism": f"ddp{world}" if world > 1 else "single",
           "steps": steps, "warmup": warmup}
    if ref is not None:
        step, nparam = reference_model_step(RL.models(ref, "cnn"))
        out["model"] = (f"reference cnn.models.VAE(latent_dim={D_LAT}, in_channels=3, distribution='clifford', "
                        f"recon_loss_type='l1') [{ref.kind} copy, unmodified], {nparam} parameters, AdamW lr 3e-4, clip 1.0; "
                        "dists.clifford = drop-in kernels")
    else:
        step = mod.make_training_step("conv", "clifford", D_LAT, batch, dev, world, local)
        out["model"] = "stand-in conv VAE (examples/train_vae_ddp.py): no copy of the reference available"
    ms, loss = mod.time_training_steps(step, steps, warmup, dev, world)
    out.update({"ms_per_step": ms, "steps_per_s": 1e3 / ms, "samples_per_s": world * batch * 1e3 / ms, "loss": loss})
    del step
    torch.cuda.empty_cache()
    if ref is not None and world == 1 and compare_reference_dists:
        try:
            step, _ = reference_model_step(RL.models(ref, "cnn", ref.clifford))
            ms_r, loss_r = mod.time_training_steps(step, max(3, steps // 2), 2, dev, world)
            out["reference_on_cuda"] = {"what": "same model and step with the reference's own dists/clifford.py classes "
                                                "on the GPU (eager PyTorch, cuFFT, torch._sample_dirichlet)",
                                        "ms_per_step": ms_r, "steps_per_s": 1e3 / ms_r, "loss": loss_r,
                                        "speedup_of_drop_in": ms_r / ms}
            del step
        except Exception as e:          # the comparator must never take the bench line down
            out["reference_on_cuda"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()
    return out


def extras(torch, lib, dev, st, peak, full=False):
    """Per-op throughput at the other BASELINE shapes (context for the headline, not bench lines):
    C1 Clifford B=128 d=512, C2 PowerSpherical / vMF B=1024 D=513, C4 bind sweep (~1 GiB per launch)."""
    out = {}

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

    for dd in ((1024, 4096, 16384) if not full else (1024, 2048, 4096, 8192, 16384)):
        N = (1 << 30) // (12 * dd)
        a = torch.randn(N, dd, device=dev)
        b = torch.randn(N, dd, device=dev)
        o = torch.empty(N, dd, device=dev)
        ms = timeit(lambda: lib.cvb_vsa_bind(a.data_ptr(), b.data_ptr(), o.data_ptr(), N, N, N, dd, 0, st))
        gb = N * 12 * dd / (ms * 1e-3) / 1e9
        out[f"C4_bind_d{dd}"] = {"vectors": N, "ms": ms, "vec_per_s": N / (ms * 1e-3), "GBps": gb, "frac": gb / peak}
        del a, b, o
    for name, B, d in (("C1_clifford_rsample_kl_B128_d512", 128, 512), ("clifford_rsample_kl_B65536_d2048", 65536, 2048)):
        loc = torch.randn(B, d, device=dev)
        kap = torch.rand(B, device=dev) * 9.87 + 0.13
        z = torch.empty(B, 2 * d, device=dev)
        kl = torch.empty(B, device=dev)
        ms = timeit(lambda: lib.cvb_clifford_ps_rsample(loc.data_ptr(), kap.data_ptr(), 1, 0, B, None, None, 7, 0,
                                                        z.data_ptr(), None, None, kl.data_ptr(), None, B, d, st))
        gb = B * (12 * d + 8) / (ms * 1e-3) / 1e9
        out[name] = {"rows": B, "d": d, "ms": ms, "samples_per_s": B / (ms * 1e-3), "GBps": gb, "frac": gb / peak}
        del loc, z
    B, D = 1024, 513
    loc = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=-1)
    kap = torch.rand(B, device=dev) * 9.2 + 0.8
    z = torch.empty(B, D, device=dev)
    save = torch.empty(B, 2, device=dev)
    ent = torch.empty(B, device=dev)
    kl = torch.empty(B, device=dev)

    def ps_step():
        lib.cvb_powerspherical_rsample(loc.data_ptr(), kap.data_ptr(), B, None, None, 3, 0, z.data_ptr(), save.data_ptr(), B, D, st)
        lib.cvb_ps_entropy_kl(kap.data_ptr(), 1, 0, B, D, (D - 1) / 2, 0, 0.0, ent.data_ptr(), kl.data_ptr(), None, st)

    def vmf_step():
        lib.cvb_vmf_rsample(loc.data_ptr(), kap.data_ptr(), B, None, None, 0, None, 3, 0, z.data_ptr(), save.data_ptr(), B, D, st)
        lib.cvb_vmf_entropy_lognorm(kap.data_ptr(), B, D, ent.data_ptr(), None, None, None, st)

    for name, fn in (("C2_powerspherical_rsample_kl_B1024_D513", ps_step), ("C2_vmf_rsample_kl_B1024_D513", vmf_step)):
        ms = timeit(fn, reps=50)
        gb = B * (8 * D + 8) / (ms * 1e-3) / 1e9
        out[name] = {"rows": B, "D": D, "ms": ms, "samples_per_s": B / (ms * 1e-3), "GBps": gb, "frac": gb / peak,
                     "note": "4 MB per launch: launch-latency bound at this batch"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--extras", action="store_true", help="time the full bind sweep in other_configs")
    ap.add_argument("--no-vae-step", action="store_true", help="skip the C3 VAE training-step leg (vae_train_step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
