#!/usr/bin/env python
"""Roofline micro-benchmarks of one B200: what a random 32-bit atomic / load / shared-memory
atomic / plain copy costs with no hashing in the way.  Prints one JSON object.

  python benchmarks/micro.py [--n 268435456]
"""

import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import pyprobables_b200 as pb

    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 28)
    a = ap.parse_args()
    ctx = pb.default_context()
    out = {}
    sizes = {"bloom_cfg2_1198MB": 1198132304 // 4, "l2_window_32MB": (32 << 20) // 4, "cms_20MB": (20 << 20) // 4, "cuckoo_4GiB": (1 << 32) // 4 - 4}
    for name, words in sizes.items():
        row = {}
        for op, label in ((0, "red_or"), (1, "red_add"), (2, "load32")):
            ms = ctx.microbench(words, a.n, op, 3)
            row[label + "_per_s"] = a.n / (ms * 1e-3)
        ms = ctx.microbench(words, 0, 3, 3)
        row["copy_GBps"] = 2 * words * 4 / (ms * 1e-3) / 1e9
        out[name] = row
    ms = ctx.microbench(1 << 20, a.n, 4, 3)
    out["smem_tile_128KB_atomic_or_per_s"] = a.n / (ms * 1e-3)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
