#!/usr/bin/env python
"""Data-parallel VAE training-step harness on top of the B200 latent kernels.

Restates the shape of the reference's training loops (mnist/mnist_clifpws.py:268-280,
cnn/cifar10_train.py:62-121): encode -> q_z = distribution(loc, kappa) -> z = q_z.rsample() ->
decode -> recon + beta * KL(q_z || prior) -> backward -> clip_grad_norm_ -> optimiser step, with
synthetic inputs.  The encoder/decoder are plain cuBLAS/cuDNN modules (not the target); the latent
distribution classes are the drop-in ones from dists.clifford / hyperspherical_vae.

    python examples/train_vae_ddp.py --arch mlp --latent clifford --dim 512 --batch 128
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_vae_ddp.py \
        --arch conv --latent clifford --dim 2048 --batch 4096

Prints one JSON line (rank 0): steps/s (max over ranks), samples/s, ms/step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "clifford-vae_b200")]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from dists.clifford import (CliffordPowerSphericalDistribution, CliffordTorusUniform, HypersphericalUniform,  # noqa: E402
                            PowerSpherical)


class LatentVAE(nn.Module):
    def __init__(self, arch: str, latent: str, dim: int):
        super().__init__()
        self.arch, self.latent, self.dim = arch, latent, dim
        zdim = 2 * dim if latent == "clifford" else dim
        if arch == "mlp":                       # 784 -> 256 -> 128 -> latent -> 128 -> 256 -> 784
            self.enc = nn.Sequential(nn.Flatten(), nn.Linear(784, 256), nn.ReLU(), nn.Linear(256, 128), nn.ReLU())
            feat = 128
            self.dec = nn.Sequential(nn.Linear(zdim, 128), nn.ReLU(), nn.Linear(128, 256), nn.ReLU(), nn.Linear(256, 784))
        else:                                   # 3x32x32 conv stack
            self.enc = nn.Sequential(
                nn.Conv2d(3, 64, 4, 2, 1), nn.SiLU(), nn.Conv2d(64, 128, 4, 2, 1), nn.SiLU(),
                nn.Conv2d(128, 256, 4, 2, 1), nn.SiLU(), nn.Flatten())
            feat = 256 * 4 * 4
            self.dec_fc = nn.Linear(zdim, feat)
            self.dec = nn.Sequential(
                nn.ConvTranspose2d(256, 128, 4, 2, 1), nn.SiLU(), nn.ConvTranspose2d(128, 64, 4, 2, 1), nn.SiLU(),
                nn.ConvTranspose2d(64, 3, 4, 2, 1))
        self.fc_loc = nn.Linear(feat, dim)
        self.fc_kappa = nn.Linear(feat, 1)
        self.floor = 0.13 if latent == "clifford" else 0.8

    def forward(self, x):
        h = self.enc(x)
        loc = self.fc_loc(h)
        kappa = torch.clamp(F.softplus(self.fc_kappa(h)) + self.floor, max=10.0)
        dev = x.device
        if self.latent == "clifford":
            q = CliffordPowerSphericalDistribution(loc, kappa, validate_args=False)
            p = CliffordTorusUniform(self.dim, device=dev, validate_args=False)
        elif self.latent == "powerspherical":
            q = PowerSpherical(F.normalize(loc, dim=-1), kappa.squeeze(-1))
            p = HypersphericalUniform(self.dim, device=dev, validate_args=False)
        else:
            from hyperspherical_vae.distributions import VonMisesFisher
            from hyperspherical_vae.distributions.hyperspherical_uniform import HypersphericalUniform as VU
            q = VonMisesFisher(F.normalize(loc, dim=-1), kappa)
            p = VU(self.dim - 1, device=dev, validate_args=False)
        z = q.rsample()
        if self.arch == "mlp":
            recon = self.dec(z)
        else:
            recon = self.dec(self.dec_fc(z).view(-1, 256, 4, 4))
        kl = torch.distributions.kl.kl_divergence(q, p).mean()
        return recon, kl


def make_training_step(arch, latent, dim, batch, dev, world, local):
    """Model + optimiser + synthetic batch for one rank; returns the closure that runs one full training step."""
    torch.backends.cudnn.benchmark = True       # cuDNN picks its conv algorithms by timing (same fp32/TF32 math)
    model = LatentVAE(arch, latent, dim).to(dev)
    if arch == "conv":                            # NHWC convolutions: +6 % on B200 (cuDNN), identical math
        model = model.to(memory_format=torch.channels_last)
    net = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.AdamW(net.parameters(), lr=3e-4)
    if arch == "mlp":
        x = (torch.rand(batch, 1, 28, 28, device=dev) > torch.rand(batch, 1, 28, 28, device=dev)).float()
    else:
        x = torch.rand(batch, 3, 32, 32, device=dev) * 2 - 1
        x = x.contiguous(memory_format=torch.channels_last)

    def step():
        recon, kl = net(x)
        if arch == "mlp":
            rec = F.binary_cross_entropy_with_logits(recon, x.view(-1, 784), reduction="sum") / x.size(0)
        else:
            rec = F.l1_loss(recon, x, reduction="sum") / x.size(0)
        loss = rec + kl
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        return loss

    return step


def time_training_steps(step, steps, warmup, dev, world):
    """ms per step over `steps` steps after `warmup`, CUDA events, max over ranks."""
    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss = step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps, float(loss.detach())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="mlp", choices=["mlp", "conv"])
    ap.add_argument("--latent", default="clifford", choices=["clifford", "powerspherical", "vmf"])
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--batch", type=int, default=128, help="per-GPU batch")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)
    step = make_training_step(args.arch, args.latent, args.dim, args.batch, dev, world, local)
    ms_step, loss = time_training_steps(step, args.steps, args.warmup, dev, world)
    ms = ms_step * args.steps
    if rank == 0:
        print(json.dumps({
            "harness": "train_vae_ddp", "arch": args.arch, "latent": args.latent, "dim": args.dim, "n_gpus": world,
            "batch_per_gpu": args.batch, "steps": args.steps, "ms_per_step": ms / args.steps,
            "steps_per_s": args.steps / (ms * 1e-3), "samples_per_s": world * args.batch * args.steps / (ms * 1e-3),
            "final_loss": loss}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
