#!/usr/bin/env python
"""BASELINE.json configs 2, 3 and 4 at FULL size on one B200: device-timed throughput plus bit-exact parity
against the CPU oracle (test infrastructure -- this script is a checker, not the product).

  python benchmarks/full_configs.py bloom  [--keys 1000000000]      # cfg 2: Bloom 1e9/0.01, 1 B inserts + checks
  python benchmarks/full_configs.py cms    [--keys 1000000000]      # cfg 3: CMS 2^20 x 5, 1 B Zipf(1.1) adds + top-1M query
  python benchmarks/full_configs.py cuckoo [--capacity-log2 28]     # cfg 4: Cuckoo 2^28 x 4 to 95 % load

Each prints one JSON object (also handy under gpurun: redirect into gpurun_out/).
"""

import argparse
import ctypes as C
import hashlib
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


OPTS: list = []


def md5(a) -> str:
    return hashlib.md5(memoryview(np.ascontiguousarray(a))).hexdigest()


def setup():
    import torch

    import pyprobables_b200 as pb

    stream = torch.cuda.Stream()
    ctx = pb.Context(0, stream=stream.cuda_stream)
    for kv in OPTS:
        name, value = kv.split("=")
        ctx.set_option(name, int(value))
    ctx.set_option("kernel_timing", 1)
    return torch, pb, stream, ctx


def timed(torch, stream, fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def run_bloom(a):
    torch, pb, stream, ctx = setup()
    from oracle import oracle as orc
    from pyprobables_b200 import _native
    from pyprobables_b200.keys import pack_keys

    n = a.keys
    out = {"config": f"BloomFilter est_elements=1e9 fpr=0.01, {n} inserts + {n} present checks + {a.absent} absent checks"}
    with torch.cuda.stream(stream):
        keys = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
        ctx.gen_uniform_keys(0, n, keys.data_ptr())
        f = pb.BloomFilter(10**9, 0.01, context=ctx)
        kb = pack_keys(keys)
        _native.call("pb_bloom_add_keys", f._h, kb.ref())  # warm-up (allocates staging)
        _native.call("pb_bloom_clear", f._h)
        ms = timed(torch, stream, lambda: _native.call("pb_bloom_add_keys", f._h, kb.ref()))
        out["insert_keys_per_s"] = n / ms * 1e3
        res = torch.empty(n, dtype=torch.uint8, device="cuda")
        ms = timed(torch, stream, lambda: _native.call("pb_bloom_check_keys", f._h, kb.ref(), C.c_void_p(res.data_ptr()), 1))
        out["check_present_keys_per_s"] = n / ms * 1e3
        out["all_present"] = bool(res.all())
        absent = torch.empty((a.absent, 16), dtype=torch.uint8, device="cuda")
        ctx.gen_uniform_keys(n, a.absent, absent.data_ptr())
        res2 = torch.empty(a.absent, dtype=torch.uint8, device="cuda")
        kb2 = pack_keys(absent)
        ms = timed(torch, stream, lambda: _native.call("pb_bloom_check_keys", f._h, kb2.ref(), C.c_void_p(res2.data_ptr()), 1))
        out["check_absent_keys_per_s"] = a.absent / ms * 1e3
        out["false_positives"] = int(res2.sum())
        out["false_positive_rate"] = out["false_positives"] / a.absent
        out["bits_set"] = f._cnt_number_bits_set()
        bitmap = f.bloom_numpy()
        out["bitmap_md5"] = md5(bitmap)
    # oracle at full size (all host cores), chunked key generation to bound host memory
    t0 = time.perf_counter()
    ob = orc.Bloom(f.number_bits, f.number_hashes)
    step = 50_000_000
    for lo in range(0, n, step):
        ob.add(orc.pack(orc.uniform_keys(lo, min(step, n - lo))))
    out["oracle_seconds"] = time.perf_counter() - t0
    out["oracle_threads"] = orc.num_threads()
    out["oracle_bits_set"] = ob.popcount()
    out["oracle_bitmap_md5"] = md5(ob.bloom)
    fp_gpu = res2.cpu().numpy().astype(bool)
    fp_orc = np.concatenate([ob.check(orc.pack(orc.uniform_keys(n + lo, min(step, a.absent - lo)))) for lo in range(0, a.absent, step)])
    out["absent_checks_equal_oracle"] = bool((fp_gpu == fp_orc).all())
    out["parity"] = bool(out["bitmap_md5"] == out["oracle_bitmap_md5"] and out["bits_set"] == out["oracle_bits_set"]
                         and out["all_present"] and out["absent_checks_equal_oracle"])
    print(json.dumps(out), flush=True)
    return out["parity"]


def run_cms(a):
    torch, pb, stream, ctx = setup()
    from oracle import oracle as orc
    from pyprobables_b200 import _native
    from pyprobables_b200.keys import pack_keys

    n = a.keys
    width, depth = 1 << 20, 5
    out = {"config": f"CountMinSketch width=2^20 depth=5, {n} Zipf(1.1) adds + query of ranks 1..10^6"}
    from concurrent.futures import ThreadPoolExecutor

    pool = ThreadPoolExecutor(max_workers=16)
    sub = 5_000_000

    def zipf_chunk(first: int, m: int) -> np.ndarray:
        """ranks [first, first+m) of the stream: piece j is default_rng([0xB200, j]).zipf(1.1, sub) (numpy
        releases the GIL while sampling, so the pieces are drawn on all host cores)"""
        assert first % sub == 0
        pieces = list(pool.map(lambda j: np.random.default_rng([0xB200, j]).zipf(1.1, min(sub, first + m - j * sub)).astype(np.uint64),
                               range(first // sub, (first + m + sub - 1) // sub)))
        return np.concatenate(pieces)

    c = pb.CountMinSketch(width=width, depth=depth, context=ctx)
    oc = orc.CMS(width, depth)
    step = 100_000_000
    add_ms = 0.0
    t_or = 0.0
    with torch.cuda.stream(stream):
        dk = torch.empty((step, 16), dtype=torch.uint8, device="cuda")
        for lo in range(0, n, step):
            m = min(step, n - lo)
            ranks = zipf_chunk(lo, m)
            dr = torch.from_numpy(ranks.view(np.int64)).cuda()
            ctx.gen_rank_keys(dr.data_ptr(), m, dk.data_ptr())
            kb = pack_keys(dk[:m])
            ea = C.c_int64(c._elements_added)
            add_ms += timed(torch, stream, lambda: _native.call("pb_cms_add_keys", c._h, kb.ref(), None, 1, C.byref(ea)))
            c._elements_added = ea.value
            t0 = time.perf_counter()
            oc.add_parallel(orc.pack(orc.rank_keys(ranks)))
            t_or += time.perf_counter() - t0
        out["add_keys_per_s"] = n / add_ms * 1e3
        top = torch.from_numpy(np.arange(1, 10**6 + 1, dtype=np.int64)).cuda()
        tk = torch.empty((10**6, 16), dtype=torch.uint8, device="cuda")
        ctx.gen_rank_keys(top.data_ptr(), 10**6, tk.data_ptr())
        est = torch.empty(10**6, dtype=torch.int64, device="cuda")
        kbt = pack_keys(tk)
        q = lambda: _native.call("pb_cms_check_keys", c._h, kbt.ref(), 0, c._elements_added, C.c_void_p(est.data_ptr()), 1)
        q()
        ms = timed(torch, stream, q)
        out["query_keys_per_s"] = 10**6 / ms * 1e3
        bins = c.bins_numpy()
    out["elements_added"] = c.elements_added
    out["bins_md5"], out["oracle_bins_md5"] = md5(bins), md5(oc.bins)
    out["oracle_seconds"], out["oracle_threads"] = t_or, orc.num_threads()
    want = oc.check(orc.pack(orc.rank_keys(np.arange(1, 10**6 + 1, dtype=np.uint64))))
    out["estimates_equal_oracle"] = bool((est.cpu().numpy() == want).all())
    out["estimate_rank1"] = int(want[0])
    out["max_bin"] = int(bins.max())
    out["kernels_ms"] = {k: {"launches": v[0], "total_ms": round(v[1], 3)} for k, v in ctx.kernel_times().items()}
    out["parity"] = bool(out["bins_md5"] == out["oracle_bins_md5"] and out["estimates_equal_oracle"]
                         and out["elements_added"] == oc.elements_added == n)
    print(json.dumps(out), flush=True)
    return out["parity"]


def run_cuckoo(a):
    torch, pb, stream, ctx = setup()
    from oracle import oracle as orc
    from pyprobables_b200 import _native
    from pyprobables_b200.keys import pack_keys

    cap = 1 << a.capacity_log2
    target = int(0.95 * cap * 4)
    out = {"config": f"CuckooFilter capacity=2^{a.capacity_log2} bucket_size=4 max_swaps=500, insert to 95 % load ({target} fingerprints)"}
    with torch.cuda.stream(stream):
        f = pb.CuckooFilter(capacity=cap, bucket_size=4, max_swaps=500, auto_expand=False, context=ctx)
        step = 1 << 26
        dk = torch.empty((step, 16), dtype=torch.uint8, device="cuda")
        consumed, added, failed_total, ins_ms = 0, 0, 0, 0.0
        n_added, n_failed = C.c_uint64(0), C.c_uint64(0)
        failed = np.empty(1 << 20, dtype=np.uint32)
        load_curve = []
        while added < target:
            # distinct 32-bit fingerprints among x keys ~ x, shrink the last batches to land on the target
            m = int(min(step, max((target - added) * 1.02 + 64, 1024)))
            ctx.gen_uniform_keys(consumed, m, dk.data_ptr())
            kb = pack_keys(dk[:m])
            holder = {}

            def go():
                holder["st"] = _native.lib().pb_cuckoo_add_keys(f._h, kb.ref(), C.byref(n_added), C.byref(n_failed),
                                                                C.c_void_p(failed.ctypes.data), failed.size)

            ms = timed(torch, stream, go)
            ins_ms += ms
            consumed += m
            added += n_added.value
            failed_total += n_failed.value
            load_curve.append({"load": added / (cap * 4), "batch_keys": m, "Mkeys_per_s": m / ms / 1e3})
            if holder["st"] not in (0, _native.PB_ERR_CUCKOO_FULL):
                _native.check(holder["st"])
            if n_failed.value:
                break
        f._inserted = added
        out.update(keys_consumed=consumed, elements_added=added, n_failed=failed_total, load_factor=added / (cap * 4),
                   insert_keys_per_s=consumed / ins_ms * 1e3, load_curve=load_curve[:: max(1, len(load_curve) // 12)] + load_curve[-1:])
        # membership of everything inserted + absent probes, timed
        res = torch.empty(step, dtype=torch.uint8, device="cuda")
        chk_ms, all_present = 0.0, True
        for lo in range(0, consumed, step):
            m = min(step, consumed - lo)
            ctx.gen_uniform_keys(lo, m, dk.data_ptr())
            kb = pack_keys(dk[:m])
            chk_ms += timed(torch, stream, lambda: _native.call("pb_cuckoo_check_keys", f._h, kb.ref(), C.c_void_p(res.data_ptr()), 1))
            all_present &= bool(res[:m].all())
        out["check_present_keys_per_s"] = consumed / chk_ms * 1e3
        out["all_inserted_present"] = all_present
        # the stored fingerprint set, sorted on the device
        p, cnt = C.c_void_p(), C.c_uint64()
        _native.call("pb_cuckoo_device_ptr", f._h, C.byref(p), C.byref(cnt))
        from pyprobables_b200.sharded import device_view

        tab = device_view(p.value, cap * 4, "<i4", 0)
        stored = tab[tab != 0]
        stored_sorted = torch.sort(stored.to(torch.int64) & 0xFFFFFFFF).values
        z = C.c_int(0)
        _native.call("pb_cuckoo_download", f._h, None, 0, C.byref(z))
        out["stored_count"] = int(stored_sorted.numel()) + int(z.value)
        # expectation from the oracle's fingerprint function: the distinct fingerprints of the consumed keys
        uniq = None
        ofp = orc.Cuckoo(16, 4, 5, 32)
        t0 = time.perf_counter()
        pieces = []
        for lo in range(0, consumed, 50_000_000):
            m = min(50_000_000, consumed - lo)
            fp = ofp.fingerprint_info(orc.pack(orc.uniform_keys(lo, m)))[2]
            pieces.append(torch.unique(torch.from_numpy(fp.astype(np.int64)).cuda()))
        uniq = torch.unique(torch.cat(pieces))
        out["oracle_seconds"] = time.perf_counter() - t0
        has_zero = bool((uniq == 0).any())
        uniq_nz = uniq[uniq != 0]
        out["distinct_fingerprints_oracle"] = int(uniq.numel())
        same_set = (uniq_nz.numel() == stored_sorted.numel()) and bool((uniq_nz == stored_sorted).all()) and (has_zero == bool(z.value))
        out["fingerprint_set_equals_oracle"] = bool(same_set) if failed_total == 0 else None
        # absent probes: expected = fingerprint of the probe is in the stored set
        pm = 100_000_000 if a.capacity_log2 >= 26 else 10_000_000
        fps_probe_all_equal = True
        positives = 0
        for lo in range(0, pm, step):
            m = min(step, pm - lo)
            ctx.gen_uniform_keys(10**10 + lo, m, dk.data_ptr())
            kb = pack_keys(dk[:m])
            _native.call("pb_cuckoo_check_keys", f._h, kb.ref(), C.c_void_p(res.data_ptr()), 1)
            pf = ofp.fingerprint_info(orc.pack(orc.uniform_keys(10**10 + lo, m)))[2]
            exp = torch.isin(torch.from_numpy(pf.astype(np.int64)).cuda(), uniq)
            torch.cuda.synchronize()
            fps_probe_all_equal &= bool((res[:m].bool() == exp).all())
            positives += int(res[:m].sum())
        out["absent_probe_positives"] = positives
        out["absent_probe_fpr"] = positives / pm
        out["absent_probes_equal_oracle"] = fps_probe_all_equal
    out["kernels_ms"] = {k: {"launches": v[0], "total_ms": round(v[1], 3)} for k, v in ctx.kernel_times().items()}
    out["parity"] = bool(failed_total == 0 and all_present and same_set and fps_probe_all_equal
                         and out["stored_count"] == added == out["distinct_fingerprints_oracle"])
    print(json.dumps(out), flush=True)
    return out["parity"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["bloom", "cms", "cuckoo"])
    ap.add_argument("--keys", type=int, default=10**9)
    ap.add_argument("--absent", type=int, default=10**8)
    ap.add_argument("--capacity-log2", type=int, default=28)
    ap.add_argument("--opt", action="append", default=[], help="context option name=value (repeatable)")
    a = ap.parse_args()
    OPTS.extend(a.opt)
    ok = {"bloom": run_bloom, "cms": run_cms, "cuckoo": run_cuckoo}[a.which](a)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
