#!/usr/bin/env python
"""Throughput of the stateful wrappers on one GPU, each with a parity check: one JSON line per part.

  python benchmarks/wrappers_bench.py [--keys 100000000] [--prefix 200000]

Parts: expanding_bloom (check-then-add batches with growth), rotating_bloom, counting_cuckoo (add / check / remove),
heavy_hitters + stream_threshold (ordered Count-Min batches with per-key return values).  Keys are generated on the
device; every part first replays a `--prefix` of the same stream through the oracle's one-key-at-a-time loop and
compares state, then times the full stream (CUDA events on the context's stream) and checks size-independent
properties.  Not part of bench.py's headline; numbers land in profiles/.
"""

from __future__ import annotations

import argparse
import hashlib
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def md5(b) -> str:
    return hashlib.md5(bytes(b)).hexdigest()


class Timer:
    def __init__(self, torch, stream):
        self.torch, self.stream = torch, stream
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        self.stream.synchronize()
        self.e0.record(self.stream)
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *a):
        self.e1.record(self.stream)
        self.stream.synchronize()
        self.ms = self.e0.elapsed_time(self.e1)
        self.wall_ms = (time.perf_counter() - self.t0) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--keys", type=int, default=100_000_000)
    ap.add_argument("--prefix", type=int, default=200_000)
    ap.add_argument("--parts", default="expanding_bloom,rotating_bloom,counting_cuckoo,heavy_hitters")
    args = ap.parse_args()
    import numpy as np
    import torch

    import pyprobables_b200 as pb
    from oracle import oracle as orc

    orc.build()
    n, npre = args.keys, args.prefix
    stream = torch.cuda.Stream()
    ctx = pb.Context(0, stream=stream.cuda_stream)
    parts = args.parts.split(",")
    with torch.cuda.stream(stream):
        keys = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
        ctx.gen_uniform_keys(0, n, keys.data_ptr())
        ctx.synchronize()
        pre_host = orc.uniform_keys(0, npre)
        assert (keys[:npre].cpu().numpy() == pre_host).all()

        def stack_state(f):
            return [(b.elements_added, md5(b.bloom_numpy().tobytes())) for b in f._blooms], f.elements_added

        def oracle_state(o):
            return [(b.elements_added, md5(b.bloom.tobytes())) for b in o.blooms], o.elements_added

        for part in parts:
            if part in ("expanding_bloom", "rotating_bloom"):
                rot = part == "rotating_bloom"
                # prefix parity: many growth steps, 2 % false positives
                kw = {"max_queue_size": 4} if rot else {}
                small = (pb.RotatingBloomFilter if rot else pb.ExpandingBloomFilter)(est_elements=npre // 16, false_positive_rate=0.02, context=ctx, **kw)
                small.add_many(keys[:npre])
                o = orc.ExpandingBloom(npre // 16, 0.02, max_queue_size=4 if rot else None)
                o.add(orc.pack(pre_host))
                parity = stack_state(small) == oracle_state(o)
                small.close()
                est = max(n // 8, 1)
                f = (pb.RotatingBloomFilter if rot else pb.ExpandingBloomFilter)(est_elements=est, false_positive_rate=0.01, context=ctx, **kw)
                f.add_many(keys[: 1 << 16])  # warm-up (allocations, module load)
                f.close()
                f = (pb.RotatingBloomFilter if rot else pb.ExpandingBloomFilter)(est_elements=est, false_positive_rate=0.01, context=ctx, **kw)
                with Timer(torch, stream) as t:
                    f.add_many(keys)
                per = [b.elements_added for b in f._blooms]
                with Timer(torch, stream) as tc:
                    found = f.check_many(keys)
                n_found = int(found.sum())
                # properties: every filter but the newest is exactly full; keys are either stored or were found when their turn came
                props = all(p == est for p in per[:-1]) and f.elements_added == n and (rot or n_found == n)
                print(json.dumps({"part": part, "keys": n, "est_elements_per_filter": est, "filters": len(per), "stored": sum(per),
                                  "skipped_as_found": n - sum(per) if not rot else None, "add_keys_per_s": n / t.ms * 1e3,
                                  "add_ms": t.ms, "check_keys_per_s": n / tc.ms * 1e3, "prefix_parity_vs_oracle": parity,
                                  "properties": props, "parity": bool(parity and props)}), flush=True)
                f.close()
            elif part == "counting_cuckoo":
                cap = 1 << 13
                small = pb.CountingCuckooFilter(capacity=cap, bucket_size=4, max_swaps=500, auto_expand=False, context=ctx)
                rng = np.random.default_rng(1)
                draws = np.minimum(rng.zipf(1.3, npre) - 1, 27_000)
                stream_host = pre_host[draws]
                small.add_many(stream_host)
                o = orc.CountingCuckoo(cap, 4, 500)
                o.add(orc.pack(stream_host))
                parity = small.bins() == o.bins() and (small.elements_added, small.unique_elements) == (o.elements_added, o.unique_elements)
                got = small.remove_many(stream_host[: npre // 2])
                parity = parity and (got == o.remove(orc.pack(stream_host[: npre // 2]))).all() and small.bins() == o.bins()
                small.close()
                log2 = max(10, int(np.ceil(np.log2(n / 4 / 0.5))))  # 50 % load with distinct keys
                f = pb.CountingCuckooFilter(capacity=1 << log2, bucket_size=4, max_swaps=500, auto_expand=False, context=ctx)
                f.add_many(keys[: 1 << 16])
                f.remove_many(keys[: 1 << 16])
                with Timer(torch, stream) as t:
                    f.add_many(keys)
                with Timer(torch, stream) as t2:
                    f.add_many(keys[: n // 2])  # second sighting: counts go to 2, nothing new is stored
                with Timer(torch, stream) as tc:
                    counts = f.check_many(keys)
                c = np.asarray(counts)
                with Timer(torch, stream) as tr:
                    removed = f.remove_many(keys[: n // 2])
                # properties (fingerprint collisions only ever raise a count)
                props = bool((c[: n // 2] >= 2).all() and (c[n // 2 :] >= 1).all() and f.elements_added == n and removed.all()
                             and c.astype(np.int64).sum() >= n + n // 2)
                print(json.dumps({"part": part, "keys": n, "capacity_log2": log2, "unique_elements": f.unique_elements,
                                  "add_new_keys_per_s": n / t.ms * 1e3, "add_seen_keys_per_s": (n // 2) / t2.ms * 1e3,
                                  "check_keys_per_s_to_host": n / tc.ms * 1e3, "remove_keys_per_s_to_host": (n // 2) / tr.ms * 1e3,
                                  "prefix_parity_vs_oracle": bool(parity), "properties": props, "parity": bool(parity and props)}), flush=True)
                f.close()
            elif part == "heavy_hitters":
                n_all, n = n, min(n, 20_000_000)  # the dictionary replay walks ~1/3 of a Zipf(1.1) stream on the host
                ranks = torch.empty(n, dtype=torch.int64, device="cuda")
                ctx.gen_zipf_ranks(0, n, ranks.data_ptr(), 1.1)
                rkeys = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
                ctx.gen_rank_keys(ranks.data_ptr(), n, rkeys.data_ptr())
                ctx.synchronize()
                pre = rkeys[:npre].cpu().numpy()
                names = [k.tobytes() for k in pre]
                hh = pb.HeavyHitters(num_hitters=100, width=1 << 14, depth=5, context=ctx)
                got = hh.add_many(pre)
                o = orc.HeavyHitters(100, 1 << 14, 5)
                want = o.add_tracked(names, orc.pack(pre))
                parity = bool((got == want).all() and hh.heavy_hitters == o.top_x)
                st = pb.StreamThreshold(threshold=200, width=1 << 14, depth=5, context=ctx)
                os_ = orc.StreamThreshold(200, 1 << 14, 5)
                st.add_many(pre), os_.add_tracked(names, orc.pack(pre))
                parity = parity and st.meets_threshold == os_.meets
                hh.close(), st.close()
                c = pb.CountMinSketch(width=1 << 20, depth=5, context=ctx)
                c.add_many_returns(rkeys[: 1 << 16])
                c.clear()
                with Timer(torch, stream) as t:
                    rets = c.add_many_returns(rkeys)
                plain = pb.CountMinSketch(width=1 << 20, depth=5, context=ctx)
                plain.add_many(rkeys)
                same_table = bool((c.bins_numpy() == plain.bins_numpy()).all())
                final = torch.from_numpy(c.check_many(rkeys[: 1 << 20])).cuda()
                props = same_table and bool((rets[: 1 << 20] <= final).all()) and int(rets.min()) >= 1
                hh = pb.HeavyHitters(num_hitters=100, width=1 << 20, depth=5, context=ctx)
                with Timer(torch, stream) as th:
                    hh.add_many(rkeys)
                top = sorted(hh.heavy_hitters.values(), reverse=True)[:3]
                st = pb.StreamThreshold(threshold=max(n // 1000, 10), width=1 << 20, depth=5, context=ctx)
                with Timer(torch, stream) as ts:
                    st.add_many(rkeys)
                print(json.dumps({"part": part, "keys": n, "add_many_returns_keys_per_s": n / t.ms * 1e3,
                                  "heavy_hitters_add_many_keys_per_s": n / th.wall_ms * 1e3, "heavy_hitters_top3": top,
                                  "stream_threshold_add_many_keys_per_s": n / ts.wall_ms * 1e3, "keys_over_threshold": len(st.meets_threshold),
                                  "prefix_parity_vs_oracle": parity, "table_equals_add_many": same_table, "properties": props,
                                  "parity": bool(parity and props)}), flush=True)
            else:
                raise SystemExit(f"unknown part {part}")


if __name__ == "__main__":
    main()
