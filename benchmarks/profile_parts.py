#!/usr/bin/env python
"""One or two launches of each hot kernel at BASELINE-size state, for `ncu --set full` captures (profiles/):

  ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 2 -o gpurun_out/prof_<part> \
      python benchmarks/profile_parts.py <part>

parts: bloom_insert (bloom_part4 + bloom_apply2), bloom_insert_w286 (the same with 286 windows: the 8-GPU geometry), bloom_check (bloom_check_fixed16), bloom_query (bloom_part4<IDS> +
bloom_probe2), cms (cms_add_fixed16, cms_check_fixed16), cuckoo (cuckoo_claim_fixed16, cuckoo_insert_kernel,
cuckoo_check_fixed16), cbloom (cbloom_add_fixed16).  Not a benchmark: numbers taken under a profiler are never reported.
"""

import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch

    import pyprobables_b200 as pb
    from pyprobables_b200 import _native
    from pyprobables_b200.keys import pack_keys

    part = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 125_000_000
    stream = torch.cuda.Stream()
    ctx = pb.Context(0, stream=stream.cuda_stream)
    with torch.cuda.stream(stream):
        keys = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
        ctx.gen_uniform_keys(0, n, keys.data_ptr())
        kb = pack_keys(keys)
        if part in ("bloom_insert", "bloom_insert_w286", "bloom_check", "bloom_query"):
            f = pb.BloomFilter(10**9, 0.01, context=ctx)
            ctx.set_option("bloom_insert_mode", 1 if part == "bloom_query" else 2)  # (query capture: no pass-1 launches before it)
            if part == "bloom_insert_w286":  # pass 1 with the window count of an 8-GPU run (286 windows, 512-key tiles)
                ctx.set_option("bloom_window_log2_bits", 25)
            _native.call("pb_bloom_add_keys", f._h, kb.ref())
            if part != "bloom_insert":
                res = torch.empty(n, dtype=torch.uint8, device="cuda")
                ctx.set_option("bloom_check_mode", 1 if part == "bloom_check" else 2)
                _native.call("pb_bloom_check_keys", f._h, kb.ref(), C.c_void_p(res.data_ptr()), 1)
        elif part == "cms":
            ranks = torch.empty(n, dtype=torch.int64, device="cuda")
            ctx.gen_zipf_ranks(0, n, ranks.data_ptr(), 1.1)
            ctx.gen_rank_keys(ranks.data_ptr(), n, keys.data_ptr())
            c = pb.CountMinSketch(width=1 << 20, depth=5, context=ctx)
            c.add_many(keys)
            est = torch.empty(n, dtype=torch.int64, device="cuda")
            _native.call("pb_cms_check_keys", c._h, kb.ref(), 0, n, C.c_void_p(est.data_ptr()), 1)
        elif part == "cuckoo":
            cap = 1 << 25  # 125 M keys -> 93 % load: the second batch shows the eviction regime
            f = pb.CuckooFilter(capacity=cap, bucket_size=4, max_swaps=500, auto_expand=False, context=ctx)
            step = 1 << 26
            for lo in range(0, n, step):
                f.add_many(keys[lo : min(lo + step, n)])
            f.check_many(keys[:step])
        elif part == "cbloom":
            f = pb.CountingBloomFilter(2 * 10**8, 0.01, context=ctx)
            f.add_many(keys)
            f.check_many(keys[: 1 << 24])
        else:
            raise SystemExit(f"unknown part {part}")
        ctx.synchronize()
    print("done", part, n)


if __name__ == "__main__":
    main()
