"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row sharding, rank-disjoint RNG seeds,
sharded bundle (all-reduce) and sharded cleanup (all-gather of (max, argmax)), max-over-ranks timing.
The local compute is injected (torch ops) because the product kernels have no CPU fallback."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world_size, init_file, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "clifford-vae_b200")]
    from clifford_b200 import distributed as D
    from clifford_b200 import _lib
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world_size)
    try:
        torch.manual_seed(0)                              # same data on every rank, then shard
        k, d, M, Q = 37, 64, 101, 9
        stack = torch.randn(k, d)
        items = torch.randn(M, d)
        query = items[[3, 50, 100, 7, 64, 99, 0, 51, 52]] + 0.05 * torch.randn(Q, d)
        s0, s1 = D.shard_rows(k)
        got = D.sharded_bundle(stack[s0:s1], k, normalize=True, local_sum=lambda v: v.sum(0))
        ref = stack.sum(0) / k ** 0.5
        assert torch.allclose(got, ref, atol=1e-5), "sharded bundle"
        m0, m1 = D.shard_rows(M)
        cos = lambda q, m: torch.nn.functional.cosine_similarity(q.unsqueeze(1), m.unsqueeze(0), dim=-1)  # noqa: E731
        best, idx = D.sharded_cleanup(query, items[m0:m1], m0, local_similarity=cos)
        full = cos(query, items)
        assert torch.equal(idx, full.argmax(1)), "sharded cleanup argmax"
        assert torch.allclose(best, full.max(1).values)
        # every row is owned exactly once
        owned = torch.zeros(k)
        owned[s0:s1] += 1
        dist.all_reduce(owned)
        assert torch.equal(owned, torch.ones(k))
        # rank-disjoint Philox streams: (seed, offset) differ across ranks for the same torch seed
        seed, off = _lib.next_rng()
        seeds = [torch.zeros(1, dtype=torch.int64) for _ in range(world_size)]
        dist.all_gather(seeds, torch.tensor([seed & 0x7FFFFFFFFFFFFFFF]))
        assert len({int(s) for s in seeds}) == world_size
        assert seed == D.rank_seed(torch.initial_seed() & 0xFFFFFFFFFFFFFFFF)
        assert D.max_over_ranks(float(rank + 1)) == float(world_size)
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, f"fail: {e!r}"))
    finally:
        dist.destroy_process_group()


def test_shard_rows_partition():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "clifford-vae_b200")]
    from clifford_b200 import distributed as D
    for total in (0, 1, 7, 8, 4096, 4099):
        for w in (1, 2, 3, 8):
            spans = [D.shard_rows(total, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    with tempfile.TemporaryDirectory() as tmp:
        init_file = os.path.join(tmp, "init")
        procs = [ctx.Process(target=_worker, args=(r, 2, init_file, out)) for r in range(2)]
        for p in procs:
            p.start()
        results = [out.get(timeout=100) for _ in procs]
        for p in procs:
            p.join(timeout=30)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
