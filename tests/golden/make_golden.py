"""Generate tests/golden/golden.json by running the UNMODIFIED pure-Python reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py            # ~3 min (config 1 is 3 M reference hash calls)

Everything the parity tests compare against that is not already a literal in the reference's own
test-suite comes from this script, so the oracle (oracle/pb_oracle.c) and the CUDA path are both
pinned to outputs of the real reference.  Key generators are the SURVEY.md 8(d) ones, restated
here in plain Python so the fixtures do not depend on the oracle either.
"""

from __future__ import annotations

import hashlib
import io
import json
import random
import struct
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference")
from probables import BloomFilter, CountMinSketch, CuckooFilter  # noqa: E402
from probables.hashes import default_fnv_1a, fnv_1a  # noqa: E402

M64 = (1 << 64) - 1
SEED = 0xB200


def sm64(x: int) -> int:
    z = (x + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def ukey(i: int) -> bytes:
    return struct.pack("<QQ", sm64((SEED + 2 * i) & M64), sm64((SEED + 2 * i + 1) & M64))


def rkey(r: int) -> bytes:
    return struct.pack("<QQ", r & M64, sm64(r & M64))


def md5(b) -> str:
    return hashlib.md5(bytes(b)).hexdigest()


def main():
    out: dict = {"reference_version": "0.7.0", "generator": "tests/golden/make_golden.py"}
    t0 = time.time()

    # ---- key generator + hash KATs
    out["keys"] = {str(i): ukey(i).hex() for i in (0, 1, 1000000)}
    out["fnv_key0_k7"] = default_fnv_1a(ukey(0), 7)
    misc = ["", "a", "this is a test", "é", "naïve café", "日本語のキー", "\U0001f600 emoji", "x" * 300]
    out["fnv_str"] = {s: default_fnv_1a(s, 4) for s in misc}
    out["fnv_bytes"] = {b.hex(): default_fnv_1a(b, 4) for b in (b"", b"\x00", b"\xff" * 7, bytes(range(256)))}
    out["fnv_seed_big"] = {str(s): fnv_1a("seed test", s) for s in (0, 1, 7, 1000, 2**32 + 5)}

    # ---- 64-bit modulo KATs (configs 2 and 5) for key_0 .. key_7
    kat = {}
    for name, (est, fpr) in {"cfg2": (10**9, 0.01), "cfg5": (10**10, 0.001), "cfg1": (10**6, 0.01)}.items():
        t_fpr, k, m = BloomFilter._get_optimized_params(est, fpr)
        kat[name] = {
            "est": est,
            "fpr": fpr,
            "t_fpr": t_fpr,
            "k": k,
            "m": m,
            "bits": [[h % m for h in default_fnv_1a(ukey(i), k)] for i in range(8)],
        }
    out["bloom_index_kat"] = kat
    out["bloom_sizing"] = {
        f"{e}/{f}": list(BloomFilter._get_optimized_params(e, f))
        for e, f in [(10, 0.05), (16_000_000, 0.001), (10**6, 0.01), (10**9, 0.01), (10**10, 0.001), (1, 0.5), (3, 0.3)]
    }

    # ---- small Bloom cases: variable-length keys, non-ASCII strs, tiny filters
    rnd = random.Random(1234)
    var_keys = [bytes(rnd.getrandbits(8) for _ in range(rnd.randrange(0, 41))) for _ in range(500)]
    probe_keys = [bytes(rnd.getrandbits(8) for _ in range(rnd.randrange(0, 41))) for _ in range(500)]
    blm = BloomFilter(est_elements=600, false_positive_rate=0.02)
    for k_ in var_keys:
        blm.add(k_)
    out["bloom_var"] = {
        "est": 600,
        "fpr": 0.02,
        "keys": [k_.hex() for k_ in var_keys],
        "probes": [k_.hex() for k_ in probe_keys],
        "bitmap_hex": bytes(blm.bloom).hex(),
        "export_hex": blm.export_hex(),
        "check_probes": [bool(blm.check(k_)) for k_ in probe_keys],
        "bits_set": blm._cnt_number_bits_set(),
    }
    ustr = ["ключ-%d" % i for i in range(50)] + ["キー%d" % i for i in range(50)] + ["clé-%d" % i for i in range(50)]
    blm = BloomFilter(est_elements=200, false_positive_rate=0.01)
    for s in ustr:
        blm.add(s)
    out["bloom_unicode"] = {
        "est": 200,
        "fpr": 0.01,
        "keys": ustr,
        "bitmap_hex": bytes(blm.bloom).hex(),
        "check_utf8_bytes": [bool(blm.check(s.encode("utf-8"))) for s in ustr],
    }
    blm = BloomFilter(est_elements=10, false_positive_rate=0.05)
    for i in range(10):
        blm.add(f"this is a test {i}")
    out["bloom_10"] = {"export_hex": blm.export_hex(), "export_md5": md5(bytes(blm))}

    # ---- config 1: Bloom 1e6 / 0.01, insert uniform keys 0..999999, probe 1e6..2e6-1
    print("config 1 (about 100 s)...", flush=True)
    blm = BloomFilter(est_elements=10**6, false_positive_rate=0.01)
    for i in range(10**6):
        blm.add(ukey(i))
    present = sum(1 for i in range(0, 10**6, 50) if blm.check(ukey(i)))
    fp_idx = [i for i in range(10**6, 2 * 10**6) if blm.check(ukey(i))]
    out["config1"] = {
        "num_bits": blm.number_bits,
        "k": blm.number_hashes,
        "bloom_length": blm.bloom_length,
        "bits_set": blm._cnt_number_bits_set(),
        "bitmap_md5": md5(blm.bloom),
        "export_md5": md5(bytes(blm)),
        "present_sampled": present,
        "present_sample_size": len(range(0, 10**6, 50)),
        "false_positives": len(fp_idx),
        "false_positive_first20": fp_idx[:20],
        "false_positive_idx_md5": md5(np.asarray(fp_idx, dtype="<u8").tobytes()),
    }
    print("  done", round(time.time() - t0), "s", flush=True)

    # ---- CMS: 2^20 x 5, 200 000 zipf(1.1) adds
    ranks = np.random.default_rng(SEED).zipf(1.1, 200_000).astype(np.int64)
    cms = CountMinSketch(width=2**20, depth=5)
    for r in ranks.tolist():
        cms.add(rkey(r))
    est = {}
    for qt in ("min", "mean", "mean-min"):
        cms.query_type = qt
        est[qt] = [cms.check(rkey(r)) for r in range(1, 1001)]
    cms.query_type = "min"
    out["cms_zipf"] = {
        "width": 2**20,
        "depth": 5,
        "n": 200_000,
        "ranks_first5": ranks[:5].tolist(),
        "ranks_md5": md5(ranks.astype("<i8").tobytes()),
        "key0": rkey(int(ranks[0])).hex(),
        "bins_md5": md5(cms._bins),
        "export_md5": md5(bytes(cms)),
        "elements_added": cms.elements_added,
        "nonzero_bins": int(np.count_nonzero(np.frombuffer(bytes(cms._bins), dtype=np.int32))),
        "estimates_1_1000": est,
    }
    # small CMS with weights, returns, all query types, odd/even depth, saturation
    small = {}
    for depth in (4, 5):
        for qt in ("min", "mean", "mean-min"):
            c = CountMinSketch(width=97, depth=depth)
            c.query_type = qt
            rr = random.Random(depth)
            seq = [(f"key-{rr.randrange(40)}", rr.randrange(1, 300)) for _ in range(400)]
            rets = [c.add(k_, n) for k_, n in seq]
            small[f"d{depth}-{qt}"] = {
                "seq": seq,
                "returns": rets,
                "bins": list(c._bins),
                "elements_added": c.elements_added,
                "checks": [c.check(f"key-{i}") for i in range(50)],
            }
    out["cms_small"] = small
    c = CountMinSketch(width=1000, depth=5)
    r1 = c.add("this is a test", 2**31 - 10)
    r2 = c.add("this is a test", 100)
    r3 = c.add("other", 7)
    out["cms_saturation"] = {"returns": [r1, r2, r3], "bins_md5": md5(c._bins), "elements_added": c.elements_added,
                             "check": c.check("this is a test")}
    c = CountMinSketch(width=1000, depth=5)
    c.add("this is a test", 100)
    out["cms_export_md5"] = md5(bytes(c))

    # ---- Cuckoo: 2^16 x 4 to 95 % load with uniform keys (membership is placement invariant)
    print("cuckoo...", flush=True)
    random.seed(0)
    cko = CuckooFilter(capacity=2**16, bucket_size=4, max_swaps=500, auto_expand=False)
    target = int(0.95 * 2**16 * 4)
    i = 0
    while cko.elements_added < target:
        cko.add(ukey(i))
        i += 1
    fps = sorted(fp for b in cko.buckets for fp in b)
    probes = [j for j in range(10_000_000, 11_000_000) if cko.check(ukey(j))]
    out["cuckoo_95"] = {
        "capacity": 2**16,
        "keys_consumed": i,
        "elements_added": cko.elements_added,
        "sorted_fp_md5": md5(np.asarray(fps, dtype="<u4").tobytes()),
        "all_present_sampled": all(cko.check(ukey(j)) for j in range(0, i, 97)),
        "probe_positives": len(probes),
        "probe_positive_idx": probes,
    }
    cko = CuckooFilter(capacity=1000, bucket_size=4, max_swaps=5)
    for j in range(1000):
        cko.add(str(j))
    out["cuckoo_1000"] = {"export_md5": md5(bytes(cko)), "elements_added": cko.elements_added}
    cko = CuckooFilter.init_error_rate(0.00001)
    for j in range(1000):
        cko.add(str(j))
    out["cuckoo_1000_err"] = {
        "export_md5": md5(bytes(cko)),
        "elements_added": cko.elements_added,
        "fp_bits": cko.fingerprint_size_bits,
        "capacity": cko.capacity,
    }
    info = {}
    for cap in (2**28, 2**10, 10000):
        c2 = CuckooFilter.__new__(CuckooFilter)  # index math only; no 2^28 Python lists
        c2._cuckoo_capacity = cap
        c2._fingerprint_size = 32
        c2._CuckooFilter__hash_func = fnv_1a
        info[str(cap)] = [list(c2._generate_fingerprint_info(ukey(j))) for j in range(8)]
    out["cuckoo_info"] = info
    # full filter
    cko = CuckooFilter(capacity=100, bucket_size=2, max_swaps=100, auto_expand=False)
    n_ok = 0
    try:
        for j in range(400):
            cko.add(ukey(j))
            n_ok += 1
    except Exception as exc:  # CuckooFilterFullError
        out["cuckoo_full"] = {"type": type(exc).__name__, "msg": str(exc), "added_before_fail": n_ok}

    path = Path(__file__).with_name("golden.json")
    path.write_text(json.dumps(out, indent=1, ensure_ascii=True))
    print("wrote", path, path.stat().st_size, "bytes in", round(time.time() - t0), "s")


if __name__ == "__main__":
    main()
