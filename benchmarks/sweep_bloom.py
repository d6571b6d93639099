#!/usr/bin/env python
"""Sweep the tuning knobs of the partitioned Bloom insert on one B200 and print per-kernel device times.

  python benchmarks/sweep_bloom.py [--keys 250000000]
"""

import argparse
import itertools
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch

    import pyprobables_b200 as pb
    from pyprobables_b200 import _native
    from pyprobables_b200.keys import pack_keys

    ap = argparse.ArgumentParser()
    ap.add_argument("--keys", type=int, default=250_000_000)
    ap.add_argument("--windows", default="24,25,26,27,28")
    ap.add_argument("--cpw", default="2,4,8")
    ap.add_argument("--modes", default="2")
    ap.add_argument("--versions", default="0", help="pass-1 CTAs per SM to sweep (0 = auto)")
    ap.add_argument("--overlap", default="0,1")
    ap.add_argument("--tiles", default="0")
    a = ap.parse_args()
    stream = torch.cuda.Stream()
    ctx = pb.Context(0, stream=stream.cuda_stream)
    with torch.cuda.stream(stream):
        keys = torch.empty((a.keys, 16), dtype=torch.uint8, device="cuda")
        ctx.gen_uniform_keys(0, a.keys, keys.data_ptr())
        filt = pb.BloomFilter(10**9, 0.01, context=ctx)
        kb = pack_keys(keys)
        ctx.set_option("kernel_timing", 1)
        ref_bits = None
        rows = []
        combos = [(int(m), int(w), int(c), int(v), int(o), int(t)) for m in a.modes.split(",") for v in a.versions.split(",")
                  for o in a.overlap.split(",") for t in a.tiles.split(",")
                  for w, c in itertools.product(a.windows.split(","), a.cpw.split(","))]
        if "1" not in a.modes.split(","):
            combos.insert(0, (1, 28, 4, 0, 0, 0))
        for mode, wl, cpw, ver, ov, tile in combos:
            ctx.set_option("bloom_insert_mode", mode)
            ctx.set_option("bloom_part_tile", tile)
            ctx.set_option("bloom_overlap", ov)
            ctx.set_option("bloom_part_ctas_per_sm", ver)  # (the round-1 "version" axis is gone: this axis is now CTAs per SM, 0 = auto)
            ctx.set_option("bloom_window_log2_bits", wl)
            ctx.set_option("bloom_apply_cpw_per_sm", cpw)
            for rep in range(2):
                _native.call("pb_bloom_clear", filt._h)
                ctx.kernel_times()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                _native.call("pb_bloom_add_keys", filt._h, kb.ref())
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                kt = ctx.kernel_times()
            bits = filt._cnt_number_bits_set()
            if ref_bits is None:
                ref_bits = bits
            row = {"mode": mode, "part_ctas_per_sm": ver, "overlap": ov, "tile": tile, "window_log2": wl, "cpw_per_sm": cpw, "ms": round(ms, 3), "Gkeys_s": round(a.keys / ms / 1e6, 3),
                   "kernels_ms": {k: round(v[1], 3) for k, v in kt.items()}, "bits_ok": bits == ref_bits}
            rows.append(row)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
