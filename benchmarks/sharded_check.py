#!/usr/bin/env python
"""Parity of the range-sharded multi-GPU Bloom filter / data-parallel Count-Min against the CPU oracle.
Run one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      benchmarks/sharded_check.py [--keys 2000000]

Every rank checks its own shard byte for byte against the oracle's single-filter bit array; rank 0 prints
"SHARDED PARITY OK" when all ranks agree.  (Test infrastructure: this script may use oracle/.)
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

    import pyprobables_b200 as pb
    from oracle import oracle as orc
    from pyprobables_b200.sharded import ShardedBloomFilter, ShardedCountMinSketch

    ap = argparse.ArgumentParser()
    ap.add_argument("--keys", type=int, default=2_000_000, help="keys per rank")
    ap.add_argument("--est", type=int, default=50_000_000)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ok = True
    n = a.keys + 1000 * rank  # ragged on purpose
    first = sum(a.keys + 1000 * r for r in range(rank))
    total = sum(a.keys + 1000 * r for r in range(world))
    host_keys = orc.uniform_keys(first, n)
    dkeys = torch.from_numpy(host_keys).cuda()
    all_keys = orc.pack(orc.uniform_keys(0, total))
    skew = torch.from_numpy(np.repeat(orc.uniform_keys(5 * 10**9 + rank, 1), 2_100_000, axis=0)).cuda()  # overflows its windows
    for mode in ("p2p", "p2p_direct", "fused", "route", "gather"):
        f = ShardedBloomFilter(a.est, 0.01, mode=mode, chunk_keys=700_000)
        f.add_many(dkeys)
        ob = orc.Bloom(f.number_bits, f.number_hashes)
        ob.add(all_keys)
        if mode in ("fused", "p2p", "p2p_direct"):  # a heavily duplicated batch must take the exact overflow path
            f.add_many(skew)
            for r in range(world):
                ob.add(orc.pack(orc.uniform_keys(5 * 10**9 + r, 1)))
        mine = f.shard_numpy()
        want = ob.bloom[f.lo // 8 : f.lo // 8 + mine.size]
        same = bool((mine == want).all())
        # membership: my keys are present, foreign probes match the oracle
        hit = f.check_many(dkeys)
        probes = orc.uniform_keys(10**9 + first, n)
        got = f.check_many(torch.from_numpy(probes).cuda()).cpu().numpy()
        exp = ob.check(orc.pack(probes))
        good = same and bool(hit.all()) and bool((got == exp).all())
        print(f"[rank {rank}] bloom mode={mode} shard bits [{f.lo},{f.hi}) bytes={mine.size} same={same} members={bool(hit.all())} "
              f"probe_match={bool((got == exp).all())}", flush=True)
        ok &= good
        f.close()
    # Count-Min: private tables + merge == one sketch over the whole stream
    ranks_all = np.random.default_rng(11).zipf(1.1, total).astype(np.uint64)
    rk = orc.rank_keys(ranks_all)
    c = ShardedCountMinSketch(1 << 16, 5)
    c.add_many(torch.from_numpy(rk[first : first + n]).cuda())
    c.merge()
    oc = orc.CMS(1 << 16, 5)
    oc.add(orc.pack(rk))
    same = bool((c.local.bins_numpy() == oc.bins).all()) and c.local.elements_added == oc.elements_added
    top = orc.rank_keys(np.arange(1, 1001, dtype=np.uint64))
    same &= bool((c.check_many(top) == oc.check(orc.pack(top))).all())
    print(f"[rank {rank}] cms merged == oracle: {same}", flush=True)
    ok &= same
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED PARITY OK" if int(t.item()) == 1 else "SHARDED PARITY FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
