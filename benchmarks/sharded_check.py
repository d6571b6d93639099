#!/usr/bin/env python
"""Parity of the range-sharded multi-GPU Bloom filter / data-parallel Count-Min against the CPU oracle.
Run one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      benchmarks/sharded_check.py [--keys 2000000]

or, on a box with ONE GPU, several ranks sharing cuda:0 (control plane over gloo, data path over CUDA IPC exactly as
between GPUs; NCCL refuses two ranks on one device):

  python -m torch.distributed.run ... --nproc-per-node 2 benchmarks/sharded_check.py --same-gpu --keys 300000

Every rank checks its own shard against the oracle: byte for byte against the oracle's single-filter bit array, or
(--verify bits, for filters whose bit array does not fit the host, e.g. BASELINE config 5) by the set of expected bit
positions -- every expected bit of the shard is set and the shard's popcount equals the number of distinct expected
positions, which together mean the shard is identical.  Rank 0 prints "SHARDED PARITY OK" when all ranks agree.
(Test infrastructure: this script may use oracle/.)
"""

import argparse
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    import torch.distributed as dist

    import pyprobables_b200 as pb  # noqa: F401
    from oracle import oracle as orc
    from pyprobables_b200.sharded import ShardedBloomFilter, ShardedCountMinSketch

    ap = argparse.ArgumentParser()
    ap.add_argument("--keys", type=int, default=2_000_000, help="keys per rank")
    ap.add_argument("--est", type=float, default=50_000_000)
    ap.add_argument("--fpr", type=float, default=0.01)
    ap.add_argument("--chunk", type=int, default=700_000)
    ap.add_argument("--modes", default="p2p,route,gather")
    ap.add_argument("--verify", default="bytes", choices=["bytes", "bits"])
    ap.add_argument("--same-gpu", action="store_true", help="all ranks on cuda:0, gloo control plane")
    ap.add_argument("--no-skew", action="store_true")
    ap.add_argument("--no-cms", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if a.same_gpu:
        local = 0
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    est = int(a.est)
    ok = True
    n = a.keys + 1000 * rank  # ragged on purpose
    first = sum(a.keys + 1000 * r for r in range(rank))
    total = sum(a.keys + 1000 * r for r in range(world))
    host_keys = orc.uniform_keys(first, n)
    dkeys = torch.from_numpy(host_keys).cuda()
    all_keys = orc.pack(orc.uniform_keys(0, total))
    skew_n = 2_100_000 if not a.same_gpu else 300_000
    for mode in a.modes.split(","):
        f = ShardedBloomFilter(est, a.fpr, mode=mode, chunk_keys=a.chunk, device=local)
        m, k = f.number_bits, f.number_hashes
        f.add_many(dkeys)
        skewed = mode == "p2p" and not a.no_skew
        if skewed:  # a heavily duplicated batch must take the exact overflow path
            skew = torch.from_numpy(np.repeat(orc.uniform_keys(5 * 10**9 + rank, 1), skew_n, axis=0)).cuda()
            f.add_many(skew)
        if a.verify == "bytes":
            ob = orc.Bloom(m, k)
            ob.add(all_keys)
            if skewed:
                for r in range(world):
                    ob.add(orc.pack(orc.uniform_keys(5 * 10**9 + r, 1)))
            mine = f.shard_numpy()
            want = ob.bloom[f.lo // 8 : f.lo // 8 + mine.size]
            same = bool((mine == want).all())
            detail = f"bytes={mine.size}"
        else:
            # expected global bit positions of every key of every rank, from the oracle's hashes
            hashes = [orc.default_fnv_1a_many(all_keys, k).reshape(-1)]
            if skewed:
                hashes += [orc.default_fnv_1a_many(orc.pack(orc.uniform_keys(5 * 10**9 + r, 1)), k).reshape(-1) for r in range(world)]
            pos = np.unique(np.concatenate(hashes) % np.uint64(m))
            pos = pos[(pos >= f.lo) & (pos < f.hi)].astype(np.int64)
            bits = f.test_bit_indices(torch.from_numpy(pos).cuda())
            pc = f.popcount_local()
            same = bool(bits.all().item()) and pc == pos.size
            detail = f"expected_bits={pos.size} popcount={pc}"
        # membership: my keys are present, foreign probes match the oracle
        hit = f.check_many(dkeys)
        probes = orc.uniform_keys(10**9 + first, n)
        got = f.check_many(torch.from_numpy(probes).cuda()).cpu().numpy()
        if a.verify == "bytes":
            exp = ob.check(orc.pack(probes))
            probe_ok = bool((got == exp).all())
        else:
            probe_ok = got.mean() < max(4 * a.fpr, 0.01)  # (no host bit array to ask: a sanity bound only)
        good = same and bool(hit.all()) and probe_ok
        print(f"[rank {rank}] bloom mode={mode} m={m} k={k} shard bits [{f.lo},{f.hi}) {detail} same={same} "
              f"members={bool(hit.all())} probe_match={probe_ok}", flush=True)
        ok &= good
        f.close()
    if not a.no_cms:
        # Count-Min: private tables + merge == one sketch over the whole stream; merging again after more adds
        # must not count anything twice
        ranks_all = np.random.default_rng(11).zipf(1.1, 2 * total).astype(np.uint64)
        rk = orc.rank_keys(ranks_all)
        c = ShardedCountMinSketch(1 << 16, 5, device=local)
        oc = orc.CMS(1 << 16, 5)
        same = True
        top = orc.rank_keys(np.arange(1, 1001, dtype=np.uint64))
        for part in range(2):
            base = part * total
            c.add_many(torch.from_numpy(rk[base + first : base + first + n]).cuda())
            c.merge()
            c.merge()  # idempotent
            oc.add(orc.pack(rk[base : base + total]))
            same &= bool((c.merged.bins_numpy() == oc.bins).all()) and c.elements_added == oc.elements_added
            same &= bool((c.check_many(top) == oc.check(orc.pack(top))).all())
        print(f"[rank {rank}] cms merged == oracle (two rounds, repeated merge): {same}", flush=True)
        ok &= same
    t = torch.tensor([1 if ok else 0])
    if not a.same_gpu:
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED PARITY OK" if int(t.item()) == 1 else "SHARDED PARITY FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
