"""Host-side logic of the range-sharded Bloom filter on CPU: ShardPlan ownership math and the
counts / all-to-all-v exchange helpers of pyprobables_b200/sharded.py, run as 2 and 3 gloo ranks.
The CUDA kernels (route / apply) are replaced here by numpy stand-ins fed with oracle hashes, so what is
under test is exactly the part that also runs around NCCL on the GPU box: who owns which bit, what gets
sent where, and that the concatenated shards equal the single-filter bit array of the oracle."""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_shard_plan_math():
    from pyprobables_b200.sharded import ShardPlan

    for m, g in ((9585058424, 1), (9585058424, 2), (143775874672, 8), (63, 2), (63, 8), (1000, 3), (32, 4)):
        p = ShardPlan.make(m, g)
        assert p.shard_bits % 32 == 0 and p.shard_bits * g >= m
        covered = 0
        for r in range(g):
            lo, hi = p.bounds(r)
            assert lo == covered or (lo == m and hi == m)
            assert lo % 32 == 0 or lo == m
            covered = hi
            assert p.shard_nbytes(r) == (hi - lo + 7) // 8
        assert covered == m
        for idx in (0, m - 1, m // 2, min(p.shard_bits - 1, m - 1), min(p.shard_bits, m - 1)):
            r = p.owner(idx)
            lo, hi = p.bounds(r)
            assert lo <= idx < hi
    with pytest.raises(ValueError):
        ShardPlan.make(0, 2)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_keys: int, est: int, fpr: float, q):
    try:
        sys.path.insert(0, str(ROOT))
        import torch
        import torch.distributed as dist

        from oracle import oracle as orc
        from pyprobables_b200.bloom import optimized_params
        from pyprobables_b200.sharded import ShardPlan, exchange_counts, exchange_indices

        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        _, k, m = optimized_params(est, fpr)
        plan = ShardPlan.make(m, world)
        lo, hi = plan.bounds(rank)
        shard = np.zeros(plan.shard_nbytes(rank), dtype=np.uint8)
        per = n_keys // world
        # ragged on purpose: the last rank gets the remainder, one chunk is empty for rank 0
        first = rank * per
        count = per if rank < world - 1 else n_keys - first
        chunk = 700
        n_chunks = torch.tensor([-(-count // chunk)], dtype=torch.int64)
        dist.all_reduce(n_chunks, op=dist.ReduceOp.MAX)
        for ci in range(int(n_chunks.item()) + 1):  # +1: an all-empty round must work too
            c0 = min(ci * chunk, count)
            c1 = min(c0 + chunk, count)
            if c1 > c0:
                keys = orc.uniform_keys(first + c0, c1 - c0)
                idx = (orc.default_fnv_1a_many(orc.pack(keys), k) % np.uint64(m)).reshape(-1).astype(np.int64)
            else:
                idx = np.zeros(0, dtype=np.int64)
            owner = plan.owner(idx)
            segs = [torch.from_numpy(np.ascontiguousarray(idx[owner == d])) for d in range(world)]
            recv_counts = exchange_counts([s.numel() for s in segs])
            recv, offs = exchange_indices(segs, recv_counts.tolist())
            got = recv.numpy()
            assert len(offs) == world + 1 and offs[-1] == got.size
            assert ((got >= lo) & (got < hi)).all(), "an index reached a rank that does not own it"
            l = got - lo
            np.bitwise_or.at(shard, l >> 3, (1 << (l & 7)).astype(np.uint8))
        # the whole filter from the oracle, sliced to this rank's byte range
        ob = orc.Bloom(m, k)
        ob.add(orc.pack(orc.uniform_keys(0, n_keys)))
        want = ob.bloom[lo // 8 : lo // 8 + shard.size]
        ok = bool((shard == want).all())
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, ok, int(shard.sum())))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, False, traceback.format_exc()))
        raise e


@pytest.mark.parametrize("world,est,fpr", [(2, 5000, 0.01), (3, 1000, 0.05)])
def test_exchange_reconstructs_the_single_filter(world, est, fpr):
    import torch.multiprocessing as mp

    from oracle import oracle as orc

    orc.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 4001, est, fpr, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, info in sorted(results):
        assert ok, f"rank {rank}: {info}"
