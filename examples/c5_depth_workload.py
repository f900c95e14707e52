#!/usr/bin/env python
"""BASELINE config 5: binding-depth workload scaled up (reference scripts/binding_depth_heatmap.py:16-39 and
scripts/rolefiller_heatmap.py:17-44), d = 8192, depths 1..32, trials batched on the device and sharded over ranks.

    python examples/c5_depth_workload.py --trials 4096
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/c5_depth_workload.py --trials 4096

Per depth m every trial does m binds and m unbinds of 8192-vectors, then a cosine to the target; the last cell also
runs the cleanup against an item memory of 1000 vectors.  Trials are independent: the only collective is the final
gather of the per-depth means (plus max-over-ranks timing).  Prints one JSON line on rank 0.
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

from clifford_b200 import distributed as D  # noqa: E402
from clifford_b200 import harness  # noqa: E402
from utils import vsa  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=8192)
    ap.add_argument("--max-depth", type=int, default=32)
    ap.add_argument("--trials", type=int, default=4096, help="trials per depth over all ranks")
    ap.add_argument("--init", default="unitary", choices=["unitary", "hrr", "clifford"])
    ap.add_argument("--unfused", action="store_true", help="2m bind/unbind launches per cell instead of the fused chain kernel")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(99)
    init = {"unitary": vsa.unitary_init, "hrr": vsa.hrr_init,
            "clifford": lambda n, d, device: harness.clifford_init(n, d // 2, device=device)}[args.init]
    t0, t1 = D.shard_rows(args.trials)
    T = t1 - t0
    depths = list(range(1, args.max_depth + 1))

    def sweep():
        cells = []
        for m in depths:
            # chunk the trials so a (T, m+1, d) block stays below ~8 GiB
            chunk = max(1, min(T, (1 << 33) // (4 * args.d * (m + 1))))
            acc, n = torch.zeros((), device=dev), 0
            for s in range(0, T, chunk):
                c = min(chunk, T - s)
                vecs = vsa.normalize_vectors(init(c * (m + 1), args.d, device=dev)).view(c, m + 1, args.d)
                acc = acc + (harness.binding_depth_cell(vecs) if args.unfused else harness.binding_depth_cell_fused(vecs)).sum()
                n += c
            cells.append(acc)
        return torch.stack(cells), n

    sweep()                                    # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    sums, n = sweep()
    b.record()
    torch.cuda.synchronize()
    ms = D.max_over_ranks(a.elapsed_time(b), device=dev)
    if world > 1:
        dist.all_reduce(sums)
    sims = (sums / args.trials).cpu().tolist()
    if rank == 0:
        bind_ops = args.trials * sum(2 * m for m in depths)
        line = {"workload": "C5 binding depth", "d": args.d, "depths": [1, args.max_depth], "trials_per_depth": args.trials,
                "init": args.init, "chain": "unfused" if args.unfused else "fused", "n_gpus": world, "ms": ms, "bind_unbind_ops_per_s": bind_ops / (ms * 1e-3),
                "trial_depth_cells_per_s": args.trials * len(depths) / (ms * 1e-3),
                "similarity_depth_1_8_32": [sims[0], sims[min(7, len(sims) - 1)], sims[-1]]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
