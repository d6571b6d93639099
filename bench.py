#!/usr/bin/env python
"""bench.py -- headline benchmark of the hash-then-scatter hot path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm   (CUDA engine)
  python bench.py --impl reference [--gpus N] [--steps K] ...    reference arm (pure-Python pyprobables on host cores)

Headline (`value`): N = 1: BASELINE configs[1] -- BloomFilter(est_elements=1e9, fpr=0.01) (m = 9 585 058 424 bits,
k = 7), one step = clear the filter and batch-insert 1e9 synthetic 16-byte keys that already sit in HBM.
N > 1 (torchrun, one rank per GPU): weak scaling of the same workload -- ONE logical filter of N*1e9 elements, bit
array range-sharded over the ranks, every rank inserts its own 1e9 keys (partition -> exchange over NVLink peer
memory -> apply).  The metric is "insert + query (Bloom, CMS)", so the same JSON line carries `parts`: Bloom check,
Count-Min add / check (configs[2]), Cuckoo add / check (configs[3]), a variable-length-key Bloom insert, each with its
own throughput, roofline and parity bit; at N = 8 `config5` is BASELINE configs[4] exactly (1e10 / 0.001, k = 10,
8 x 1e9 inserts) with per-shard parity.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 0xB200
EST_PER_GPU = 10**9
FPR = 0.01
METRIC = "bloom_batch_insert_keys_per_sec"
UNIT = "keys/s"


# ----------------------------------------------------------------------------- helpers
def sm64(x: np.ndarray) -> np.ndarray:
    """splitmix64 (SURVEY 8d synthetic key generator), vectorised; used only to make host keys for the
    reference arm -- the GPU arm generates the same keys on the device (pb_gen_uniform_keys)"""
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def host_uniform_keys(first: int, n: int) -> np.ndarray:
    g = np.arange(first, first + n, dtype=np.uint64)
    out = np.empty((n, 2), dtype=np.uint64)
    with np.errstate(over="ignore"):
        out[:, 0] = sm64(np.uint64(SEED) + np.uint64(2) * g)
        out[:, 1] = sm64(np.uint64(SEED) + np.uint64(2) * g + np.uint64(1))
    return out.view(np.uint8).reshape(n, 16)


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        try:
            for line in Path(self.path).read_text().splitlines():
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                    pw.append(float(p[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=float(max(pw)))
        return out


def peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic() -> dict:
    """DRAM bytes per key of the hot kernels from the committed `ncu --set full` capture (profiles/r2_traffic.json)"""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            return json.loads((ROOT / "profiles" / name).read_text()) | {"_file": f"profiles/{name}"}
        except Exception:
            continue
    return {}


# ----------------------------------------------------------------------------- reference arm
def run_reference(args) -> None:
    """the reference's own implementation of the path on the host: pure-Python pyprobables BloomFilter.add from
    baseline/_ref (unmodified, its public API; __graft_entry__.build() installs it).  The reference is single-threaded
    by construction, so cores = 1; the compiled oracle port on all cores is reported beside it as `port_all_cores`.
    Only when the reference cannot be imported is the oracle port itself the timed arm (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_gpus = args.gpus
    est = EST_PER_GPU * n_gpus
    sample = args.ref_keys
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic"}
    ref_dir = ROOT / "baseline" / "_ref"
    kind = "reference"
    try:
        if not (ref_dir / "probables" / "__init__.py").exists():
            import __graft_entry__ as entry

            entry.install_reference()  # only possible where /root/reference exists (the build container)
        sys.path.insert(0, str(ref_dir))
        from probables import BloomFilter  # type: ignore

        blm = BloomFilter(est_elements=est, false_positive_rate=FPR)
        m, k = blm.number_bits, blm.number_hashes

        def step(i):
            keys = [bytes(r) for r in host_uniform_keys(i * sample, sample)]
            t0 = time.perf_counter()
            for key in keys:
                blm.add(key)
            return time.perf_counter() - t0

        cores = 1
    except Exception as e:  # baseline/_ref missing: time the oracle port instead (the oracle always exists)
        kind = "port"
        from oracle import oracle as orc

        orc.set_threads(host_threads())  # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
        fpr32, k, m, _ = orc.bloom_params(est, FPR)
        ob = orc.Bloom(m, k)
        sample = max(sample, 20_000_000)
        cores = orc.num_threads()

        def step(i):
            keys = orc.pack(orc.uniform_keys(i * sample, sample))
            t0 = time.perf_counter()
            ob.add(keys)
            return time.perf_counter() - t0

        line["note"] = f"baseline/_ref unavailable ({type(e).__name__}: {str(e)[:80]}); timed the oracle port"
    for i in range(args.warmup):
        step(i)
    total = sum(step(args.warmup + i) for i in range(args.steps))
    value = sample * args.steps / total
    line.update(value=value, ms_per_step=1e3 * total / args.steps,
                config={"workload": f"BloomFilter est_elements={est:.0e} fpr={FPR} (m={m} bits, k={k}); batch insert of 16-byte keys; "
                                    f"each step = a bounded sample of {sample} keys of that workload on the host",
                        "key_bytes": 16, "l2_policy": "bit array (>= 1.2 GB) far larger than any cache"},
                cpu_baseline={"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                              "sample": f"{sample} keys/step x {args.steps} steps, host threads available={host_threads()}"},
                e2e={"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                gpu_launches=0)
    if kind == "reference" and not args.no_port:
        try:
            from oracle import oracle as orc

            orc.set_threads(host_threads())
            _, ok, om, _ = orc.bloom_params(est, FPR)
            ob = orc.Bloom(om, ok)
            ns = 20_000_000
            kk = orc.pack(orc.uniform_keys(0, ns))
            ob.add(orc.pack(orc.uniform_keys(ns, 1_000_000)))  # touch
            t0 = time.perf_counter()
            ob.add(kk)
            dt = time.perf_counter() - t0
            line["port_all_cores"] = {"value": ns / dt, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                                      "sample": f"{ns} keys, C oracle with OpenMP"}
        except Exception as e:
            line["port_all_cores"] = {"unavailable": str(e)[:200]}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def cpu_baseline_port(m: int, k: int, seconds: float = 12.0) -> dict:
    """the oracle (C port of the reference algorithm) on all host cores, bounded sample of the same workload"""
    from oracle import oracle as orc

    orc.set_threads(host_threads())
    ob = orc.Bloom(m, k)
    probe = 4_000_000
    t0 = time.perf_counter()
    ob.add(orc.pack(orc.uniform_keys(0, probe)))
    rate = probe / (time.perf_counter() - t0)
    n = int(min(max(rate * seconds, probe), 400_000_000))
    keys = orc.pack(orc.uniform_keys(probe, n))
    t0 = time.perf_counter()
    ob.add(keys)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
            "sample": f"{n} of the workload's keys into the full-size {m}-bit array, C oracle + OpenMP, host threads={host_threads()}"}


class Timer:
    """device time of a region on the engine's stream (CUDA events on that stream)"""

    def __init__(self, torch, stream):
        self.torch, self.stream = torch, stream

    def ms(self, fn, reps: int = 1) -> float:
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(reps):
            fn()
        e1.record(self.stream)
        self.torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps


def expected_positions(orc, world: int, keys_per_rank: int, prefix: int, m: int, k: int) -> np.ndarray:
    """distinct global bit positions of the first `prefix` keys of every rank (oracle hashes)"""
    parts = []
    for r in range(world):
        h = orc.default_fnv_1a_many(orc.pack(orc.uniform_keys(r * keys_per_rank, prefix)), k).reshape(-1)
        parts.append(h % np.uint64(m))
    return np.unique(np.concatenate(parts))


def shard_parity(torch, dist, orc, filt, keys, n_keys: int, prefix: int, dev) -> dict:
    """per-rank parity of a sharded filter against the oracle on a bounded key prefix: the filter is cleared, every
    rank inserts its first `prefix` keys, and each rank checks that every expected bit of its shard is set and that the
    shard's popcount equals the number of distinct expected positions inside it (together: the shard is identical
    to that slice of the reference's bit array)."""
    world = filt.world
    filt.clear()
    filt.add_many(keys[:prefix])
    pos = expected_positions(orc, world, n_keys, prefix, filt.number_bits, filt.number_hashes)
    mine = pos[(pos >= filt.lo) & (pos < filt.hi)].astype(np.int64)
    bits = filt.test_bit_indices(torch.from_numpy(mine).to(dev))
    pc = filt.popcount_local()
    ok = bool(bits.all().item()) and pc == mine.size
    flags = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(flags, torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev))
    pcs = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(pcs, torch.tensor([pc], dtype=torch.int64, device=dev))
    return {"parity": all(int(f.item()) == 1 for f in flags), "per_rank": [bool(int(f.item())) for f in flags],
            "prefix_keys_per_rank": prefix, "shard_popcounts": [int(p.item()) for p in pcs],
            "distinct_expected_positions": int(pos.size),
            "how": "expected bit positions from the oracle's hashes of every rank's key prefix: all set, and popcount == their count, per shard"}


def run_parts(args, torch, pb, ctx, stream, timer, filt, keys, hbm_peak) -> dict:
    """the rest of the metric on one GPU: Bloom query, Count-Min add/query (configs[2]), Cuckoo add/query (configs[3]),
    variable-length-key Bloom insert.  Every part: device-timed throughput, roofline in SURVEY 8(d)'s bytes, parity."""
    from oracle import oracle as orc
    from pyprobables_b200 import _native
    from pyprobables_b200.keys import KeyBatch, pack_keys

    orc.set_threads(host_threads())
    parts: dict = {}
    n = int(keys.shape[0])
    dev = keys.device
    m, k = filt.number_bits, filt.number_hashes

    def guarded(name):
        def deco(fn):
            try:
                parts[name] = fn()
            except Exception as e:  # a failing part must not take the headline down
                parts[name] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
            return fn
        return deco

    # ---- Bloom query (bloom.py:252-272): the filter holds the n keys of the timed steps
    @guarded("bloom_check")
    def _():
        res = torch.empty(n, dtype=torch.uint8, device=dev)
        kb = pack_keys(keys)
        call = lambda: _native.call("pb_bloom_check_keys", filt._h, kb.ref(), C.c_void_p(res.data_ptr()), 1)
        call()
        ms_p = timer.ms(call, 2)  # auto mode: samples the hit rate, members -> partitioned query
        all_present = bool(res.all().item())
        ctx.set_option("bloom_check_mode", 1)
        try:
            call()
            ms_direct = timer.ms(call, 1)
            all_present &= bool(res.all().item())
        finally:
            ctx.set_option("bloom_check_mode", 0)
        na = min(n, 250_000_000)
        absent = torch.empty((na, 16), dtype=torch.uint8, device=dev)
        ctx.gen_uniform_keys(10 * 10**9, na, absent.data_ptr(), seed=SEED)
        kba = pack_keys(absent)
        resa = res[:na]
        calla = lambda: _native.call("pb_bloom_check_keys", filt._h, kba.ref(), C.c_void_p(resa.data_ptr()), 1)
        calla()
        ms_a = timer.ms(calla, 2)
        # parity: the oracle's check on the very same bit array (downloaded) for a sample of absent probes
        ob = orc.Bloom(m, k)
        ob.bloom[:] = filt.bloom_numpy()
        ns = min(na, 2_000_000)
        want = ob.check(orc.pack(orc.uniform_keys(10 * 10**9, ns)))
        same = bool((resa[:ns].cpu().numpy().astype(bool) == want).all())
        words = (filt.bloom_length + 3) // 4
        ms_ld = ctx.microbench(words, 1 << 28, 2, 3)
        load_ceiling = (1 << 28) / (ms_ld * 1e-3)
        v = n / (ms_p * 1e-3)
        algo = 16 + 32 * k + 1
        fpr_seen = float(resa.float().mean().item())
        del absent
        return {"value": v, "unit": UNIT, "keys": n, "present_keys_per_s": v, "absent_keys_per_s": na / (ms_a * 1e-3),
                "present_keys_per_s_direct_kernel": n / (ms_direct * 1e-3), "false_positive_rate": fpr_seen,
                "roofline": {"bound": "hbm", "kernel": "bloom_part4<IDS> + bloom_probe2 (members, partitioned); bloom_check_fixed16 (absent keys, direct)",
                             "algorithmic_bytes_per_key": algo,
                             "achieved": v * algo / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": v * algo / 1e9 / hbm_peak,
                             "random_sector_loads_per_s": v * k, "random_load_ceiling_per_s": load_ceiling,
                             "frac_of_random_load_ceiling": v * k / load_ceiling},
                "parity": all_present and same, "all_inserted_present": all_present, "absent_sample_equals_oracle": same}

    # ---- Count-Min add + query, BASELINE configs[2]: 2^20 x 5, n Zipf(1.1) adds, query ranks 1..10^6
    @guarded("cms")
    def _():
        width, depth = 1 << 20, 5
        ranks = torch.empty(n, dtype=torch.int64, device=dev)
        ctx.gen_zipf_ranks(0, n, ranks.data_ptr(), 1.1, SEED)
        ctx.gen_rank_keys(ranks.data_ptr(), n, keys.data_ptr())  # the Bloom keys are not needed any more
        c = pb.CountMinSketch(width=width, depth=depth, context=ctx)
        kb = pack_keys(keys)
        ea = C.c_int64(0)
        add = lambda: _native.call("pb_cms_add_keys", c._h, kb.ref(), None, 1, C.byref(ea))
        add()
        times = []
        for _ in range(3):
            _native.call("pb_cms_clear", c._h)
            ea.value = 0
            times.append(timer.ms(add))
        ms_add = float(np.median(times))
        c._elements_added = n
        nq = 10**6
        top = torch.arange(1, nq + 1, dtype=torch.int64, device=dev)
        tk = torch.empty((nq, 16), dtype=torch.uint8, device=dev)
        ctx.gen_rank_keys(top.data_ptr(), nq, tk.data_ptr())
        est = torch.empty(nq, dtype=torch.int64, device=dev)
        kbt = pack_keys(tk)
        q = lambda: _native.call("pb_cms_check_keys", c._h, kbt.ref(), 0, n, C.c_void_p(est.data_ptr()), 1)
        q()
        ms_q = timer.ms(q, 5)
        # parity of the query: the oracle's query on the very same table; of the add: a fresh sketch over a 2e7-key
        # prefix of the same stream against the oracle, plus linearity of the full-size table
        bins = c.bins_numpy()
        oc = orc.CMS(width, depth)
        oc.bins[:] = bins
        oc.elements_added = n
        q_same = bool((est.cpu().numpy() == oc.check(orc.pack(orc.rank_keys(np.arange(1, nq + 1, dtype=np.uint64))))).all())
        npre = min(n, 20_000_000)
        c2 = pb.CountMinSketch(width=width, depth=depth, context=ctx)
        c2.add_many(keys[:npre])
        o2 = orc.CMS(width, depth)
        o2.add_parallel(orc.pack(orc.rank_keys(ranks[:npre].cpu().numpy().astype(np.uint64))))
        a_same = bool((c2.bins_numpy() == o2.bins).all())
        rows_sum_ok = bool((bins.reshape(depth, -1).astype(np.int64).sum(axis=1) == n).all())
        ms_at = ctx.microbench(width * depth, 1 << 28, 1, 3)
        ceiling = (1 << 28) / (ms_at * 1e-3)
        v = n / (ms_add * 1e-3)
        ctx.gen_uniform_keys(0, n, keys.data_ptr(), seed=SEED)  # restore the uniform keys for the parts below
        return {"add": {"value": v, "unit": UNIT, "keys": n, "stream": "Zipf(1.1) ranks drawn on the device (pb_gen_zipf_ranks)",
                        "roofline": {"bound": "l2-atomic", "kernel": "cms_add_fixed16", "algorithmic_bytes_per_key": 16,
                                     "achieved": v * 16 / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": v * 16 / 1e9 / hbm_peak,
                                     "counter_updates_per_s": v * depth, "random_atomic_add_ceiling_per_s": ceiling,
                                     "frac_of_atomic_ceiling": v * depth / ceiling,
                                     "note": "20 MiB table is L2 resident: HBM only carries the 16 B key; the yardstick is the "
                                             "same-run random atomicAdd ceiling on an array of the same size"},
                        "parity": a_same and rows_sum_ok, "prefix_table_equals_oracle": a_same, "row_sums_equal_adds": rows_sum_ok},
                "check": {"value": nq / (ms_q * 1e-3), "unit": UNIT, "keys": nq,
                          "roofline": {"bound": "l2-gather", "kernel": "cms_check_fixed16", "algorithmic_bytes_per_key": 24,
                                       "achieved": nq / (ms_q * 1e-3) * 24 / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                       "note": "1e6 keys is one launch of a few tens of microseconds: launch-latency bound"},
                          "parity": q_same}}

    # ---- Cuckoo add + query, BASELINE configs[3]: 2^28 x 4, max_swaps 500, insert to 95 % load
    @guarded("cuckoo")
    def _():
        cap = 1 << args.cuckoo_log2
        target = int(0.95 * cap * 4)
        step = 1 << 26
        # warm-up: a throw-away filter takes one full-size and one small batch so that the context's scratch buffers
        # (fingerprint lists, claim bitmap / claim set) exist before anything is timed
        warm = pb.CuckooFilter(capacity=1 << min(25, args.cuckoo_log2), bucket_size=4, max_swaps=500, auto_expand=False, context=ctx)
        warm.add_many(keys[: min(n, step)])
        warm.add_many(keys[: min(n, 1 << 22)])
        warm.close()
        f = pb.CuckooFilter(capacity=cap, bucket_size=4, max_swaps=500, auto_expand=False, context=ctx)
        spare = torch.empty((step, 16), dtype=torch.uint8, device=dev)
        consumed = added = failed_total = 0
        ins_ms = 0.0
        n_added, n_failed = C.c_uint64(0), C.c_uint64(0)
        failed = np.empty(1 << 20, dtype=np.uint32)
        curve = []
        while added < target:
            mkeys = int(min(step, max((target - added) * 1.02 + 64, 1024)))
            if consumed + mkeys <= n:
                batch = keys[consumed : consumed + mkeys]
            else:  # beyond the resident keys: generate the continuation of the same key stream
                ctx.gen_uniform_keys(consumed, mkeys, spare.data_ptr(), seed=SEED)
                batch = spare[:mkeys]
            kb = pack_keys(batch)
            holder = {}

            def go():
                holder["st"] = _native.lib().pb_cuckoo_add_keys(f._h, kb.ref(), C.byref(n_added), C.byref(n_failed),
                                                                C.c_void_p(failed.ctypes.data), failed.size)

            ms = timer.ms(go)
            ins_ms += ms
            consumed += mkeys
            added += n_added.value
            failed_total += n_failed.value
            curve.append({"load": round(added / (cap * 4), 4), "batch_keys": mkeys, "Mkeys_per_s": round(mkeys / ms / 1e3, 1)})
            if holder["st"] not in (0, _native.PB_ERR_CUCKOO_FULL):
                _native.check(holder["st"])
            if n_failed.value:
                break
        f._inserted = added
        nchk = min(consumed, n, 1 << 28)
        res = torch.empty(nchk, dtype=torch.uint8, device=dev)
        kbc = pack_keys(keys[:nchk])
        chk = lambda: _native.call("pb_cuckoo_check_keys", f._h, kbc.ref(), C.c_void_p(res.data_ptr()), 1)
        chk()
        ms_c = timer.ms(chk, 2)
        present = bool(res.all().item())
        before = added
        f.add_many(keys[: min(n, 1 << 24)])  # re-adding is a no-op (dedupe, cuckoo.py:300-302)
        noop = f.elements_added == before
        cnt = C.c_uint64()
        _native.call("pb_cuckoo_count", f._h, C.byref(cnt))
        f.close()
        # exact parity at a size the oracle finishes in seconds: 2^20 x 4 to 95 %, stored fingerprint set == oracle's
        cap2 = 1 << min(20, args.cuckoo_log2)
        n2 = min(n, int(cap2 * 4 * 0.95))
        f2 = pb.CuckooFilter(capacity=cap2, bucket_size=4, max_swaps=500, auto_expand=False, context=ctx)
        f2.add_many(keys[:n2])
        of = orc.Cuckoo(cap2, 4, 500, 32)
        of.add(orc.pack(orc.uniform_keys(0, n2)), failed_cap=1 << 16)
        small_same = bool((f2.fingerprints() == of.fingerprints()).all()) and f2.elements_added == of.elements_added
        f2.close()
        v = consumed / (ins_ms * 1e-3)
        vc = nchk / (ms_c * 1e-3)
        return {"add": {"value": v, "unit": UNIT, "keys": consumed, "elements_added": added, "n_failed": failed_total,
                        "load_factor": added / (cap * 4), "load_curve": curve[:: max(1, len(curve) // 10)] + curve[-1:],
                        "roofline": {"bound": "hbm-latency", "kernel": "cuckoo_claim_fixed16 + cuckoo_insert_kernel",
                                     "algorithmic_bytes_per_key": 112, "achieved": v * 112 / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": v * 112 / 1e9 / hbm_peak},
                        "parity": failed_total == 0 and present and noop and cnt.value == added and small_same,
                        "stored_slots_equal_elements_added": cnt.value == added, "readd_is_noop": noop,
                        "fingerprint_set_equals_oracle_at_2^20": small_same},
                "check": {"value": vc, "unit": UNIT, "keys": nchk,
                          "roofline": {"bound": "hbm", "kernel": "cuckoo_check_fixed16", "algorithmic_bytes_per_key": 81,
                                       "achieved": vc * 81 / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": vc * 81 / 1e9 / hbm_peak},
                          "parity": present}}

    # ---- Bloom insert of variable-length keys (8..40 bytes, device resident): the same partitioned path through
    # TMA-staged tiles (the reference's normal KeyT is str / bytes of any length, hashes.py:10)
    @guarded("bloom_insert_ragged")
    def _():
        nr = min(n // 2, args.ragged_keys)
        lens = 8 + (torch.arange(nr, dtype=torch.int64, device=dev) * 7) % 33
        offs = torch.zeros(nr + 1, dtype=torch.int64, device=dev)
        torch.cumsum(lens, 0, out=offs[1:])
        total = int(offs[-1].item())
        data = keys.reshape(-1)[:total]  # the uniform key bytes, re-cut into ragged keys
        kb = KeyBatch(data.data_ptr(), offs.data_ptr(), nr, 0, 1, True, (data, offs))
        f = pb.BloomFilter(EST_PER_GPU, FPR, context=ctx)
        ins = lambda: _native.call("pb_bloom_add_keys", f._h, kb.ref())
        ins()
        times = []
        for _ in range(3):
            _native.call("pb_bloom_clear", f._h)
            times.append(timer.ms(ins))
        ms_r = float(np.median(times))
        npre = min(nr, 3_000_000)
        _native.call("pb_bloom_clear", f._h)
        kbp = KeyBatch(data.data_ptr(), offs.data_ptr(), npre, 0, 1, True, (data, offs))
        ctx.set_option("bloom_insert_mode", 2)
        try:
            _native.call("pb_bloom_add_keys", f._h, kbp.ref())
        finally:
            ctx.set_option("bloom_insert_mode", args.insert_mode)
        hoffs = offs[: npre + 1].cpu().numpy().astype(np.uint64)
        hdata = data[: int(hoffs[-1])].cpu().numpy()
        ob = orc.Bloom(m, k)
        ob.add(orc.pack((hdata, hoffs)))
        same = bool((f.bloom_numpy() == ob.bloom).all())
        f.close()
        v = nr / (ms_r * 1e-3)
        avg = total / nr
        return {"value": v, "unit": UNIT, "keys": nr, "key_bytes": "8..40 (mean %.1f)" % avg,
                "roofline": {"bound": "hbm", "kernel": "bloom_part4<7,256,KeySrcStaged<1>> + bloom_apply2",
                             "algorithmic_bytes_per_key": avg + 8 + 64 * k, "achieved": v * (avg + 8 + 64 * k) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": v * (avg + 8 + 64 * k) / 1e9 / hbm_peak},
                "parity": same, "how": f"fresh filter, first {npre} keys through the partitioned path, bit array == oracle's"}

    return parts


def run_ours(args) -> None:
    import torch

    import pyprobables_b200 as pb
    from pyprobables_b200 import _native
    from pyprobables_b200.keys import pack_keys

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = args.gpus
    if world != n_gpus:
        if world == 1 and n_gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
        n_gpus = world
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    stream = torch.cuda.Stream(device=dev)
    ctx = pb.Context(local, stream=stream.cuda_stream)
    n_keys = args.keys
    hbm_peak, peak_src = peaks()
    ctx.set_option("bloom_insert_mode", args.insert_mode)
    if args.window_log2:
        ctx.set_option("bloom_window_log2_bits", args.window_log2)
    for kv in args.opt:
        name, value = kv.split("=")
        ctx.set_option(name, int(value))
    timer = Timer(torch, stream)
    parity = None
    config5 = None
    parts = None

    with torch.cuda.stream(stream):
        keys = torch.empty((n_keys, 16), dtype=torch.uint8, device=dev)
        ctx.gen_uniform_keys(rank * n_keys, n_keys, keys.data_ptr(), seed=SEED)
        if world == 1:
            filt = pb.BloomFilter(EST_PER_GPU, FPR, context=ctx)
            m, k = filt.number_bits, filt.number_hashes
            kb = pack_keys(keys)

            def step():
                _native.call("pb_bloom_clear", filt._h)
                _native.call("pb_bloom_add_keys", filt._h, kb.ref())

            bitmap_bytes = filt.bloom_length
        else:
            from pyprobables_b200.sharded import ShardedBloomFilter

            filt = ShardedBloomFilter(EST_PER_GPU * world, FPR, device=local, context=ctx, chunk_keys=args.chunk_keys,
                                      mode=args.shard_mode, window_log2=args.window_log2 or 27)
            m, k = filt.number_bits, filt.number_hashes
            bitmap_bytes = filt.plan.shard_nbytes(rank)

            def step():
                filt.clear()
                filt.add_many(keys)

        def barrier():
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()  # before the warm-up: nvidia-smi needs up to a second before its first sample
        for _ in range(max(args.warmup, 0)):
            step()
        barrier()
        ctxs = [ctx] + [c for c in (getattr(filt, "_ctx_part", None), getattr(filt, "_ctx_state", None)) if c is not None]
        for c in ctxs[1:]:
            for kv in args.opt:
                name, value = kv.split("=")
                c.set_option(name, int(value))
        for c in ctxs:
            c.set_option("kernel_timing", 1)
            c.kernel_times()
        launches0 = sum(c.launch_count for c in ctxs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        if world > 1:
            filt.join()  # the sharded insert returns after its pass 1: order the stream after the exchange and pass 2 too
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else {}
        launches = sum(c.launch_count for c in ctxs) - launches0
        ktimes = {}
        for c in ctxs:
            for kn, (cnt, tot) in c.kernel_times().items():
                old = ktimes.get(kn, (0, 0.0))
                ktimes[kn] = (old[0] + cnt, old[1] + tot)
            c.set_option("kernel_timing", 0)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        total_keys = n_keys * world * args.steps
        value = total_keys / (ms * 1e-3)

        # size-independent sanity of the timed state: inserted keys are members
        probe_n = min(n_keys, 1 << 20)
        ok = bool(filt.check_many(keys[:probe_n]).all())
        if world == 1:
            setbits = filt._cnt_number_bits_set()
        else:
            pc = torch.tensor([filt.popcount_local()], dtype=torch.int64, device=dev)
            dist.all_reduce(pc)
            setbits = int(pc.item())
        assert ok, "inserted keys are not members: the timed state is wrong"

        # ---- roofline.  The insert is two kernels that run CONCURRENTLY (pass 2 of chunk i beside pass 1 of chunk i+1),
        # so the hot path's "launch" is the pair: achieved = SURVEY 8(d)'s algorithmic bytes of the whole insert
        # (16 + 64k per key, the random-atomic model) / the step's duration.  `kernels` charges each pass the bytes IT has
        # to move in this design over its own event-timed duration (they overlap, so the shares add to more than 1).
        roof = None
        atomic = None
        if ktimes:
            algo_per_key = 16 + 64 * k
            n_part = max(1, ktimes.get("bloom_part", (1, 0))[0])
            chunks_per_step = max(1, n_part // max(args.steps, 1))
            per = {}
            design = {"bloom_part": 16 + 4 * k, "bloom_apply_windows": 4 * k + 2.0 * bitmap_bytes * chunks_per_step / n_keys,
                      "bloom_add": algo_per_key}
            for kn, (cnt, tot) in ktimes.items():
                d = design.get(kn)
                per[kn] = {"launches": cnt, "total_ms": tot, "avg_launch_ms": tot / max(cnt, 1), "share_of_step": tot / ms}
                if d is not None and tot > 0:
                    per[kn].update(design_bytes_per_key=d, achieved_GBps=d * n_keys * args.steps / (tot * 1e-3) / 1e9)
            step_gbps = value / world * algo_per_key / 1e9
            tj = ncu_traffic()
            traffic, traffic_src = None, None
            try:
                per_key = tj["bloom_part"]["dram_bytes_per_key"] + tj["bloom_apply_windows"]["list_dram_bytes_per_key"]
                sweep = tj["bloom_apply_windows"]["bitmap_sweep_bytes_per_launch"] * chunks_per_step if world == 1 else 2.0 * bitmap_bytes * chunks_per_step
                traffic = per_key * n_keys + sweep
                traffic_src = (f"{tj['_file']} (ncu --set full, dram__bytes_read+write): {per_key:.1f} DRAM B/key for keys + staged lists "
                               f"x {n_keys} keys + one read+write of the bit array per chunk x {chunks_per_step} chunks")
            except Exception:
                pass
            dom = max(ktimes.items(), key=lambda kv: kv[1][1])[0]
            roof = {"bound": "hbm", "kernel": "bloom_part4 + bloom_apply2 (concurrent; dominant by device time: %s)" % dom,
                    "achieved": step_gbps, "peak": hbm_peak, "unit": "GB/s", "frac": step_gbps / hbm_peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_key": algo_per_key, "algorithmic_bytes_per_step": algo_per_key * n_keys,
                    "keys_per_step_per_gpu": n_keys, "chunks_per_step": chunks_per_step, "kernels": per,
                    "note": ("achieved = SURVEY 8(d) algorithmic bytes of a Bloom insert (16 + 64k per key: the key + k random "
                             "32-B-sector read-modify-writes) x keys per step / step time.  The partitioned design moves ~5x fewer "
                             "DRAM bytes than that model (`traffic`), which is how it passes the random-atomic ceiling "
                             "(`atomic_roofline`).")}
        if rank == 0 and world == 1 and not args.no_micro:
            words = (bitmap_bytes + 3) // 4
            n_at = 1 << 28
            ms_or = ctx.microbench(words, n_at, 0, 3)
            ms_cp = ctx.microbench(words, 0, 3, 3)
            ms_l2 = ctx.microbench(1 << 22, n_at, 0, 3)  # 16 MiB: one L2-resident window
            ceil_rate = n_at / (ms_or * 1e-3)
            atomic = {"random_red_or_per_s": ceil_rate, "bloom_bit_updates_per_s": value * k,
                      "frac_of_random_atomic_ceiling": value * k / ceil_rate,
                      "l2_resident_red_or_per_s": n_at / (ms_l2 * 1e-3),
                      "frac_of_l2_atomic_ceiling": value * k / (n_at / (ms_l2 * 1e-3)),
                      "copy_GBps_same_run": 2 * words * 4 / (ms_cp * 1e-3) / 1e9,
                      "how": f"{n_at} RED.OR.b32 at pre-generated uniform indices over a {words * 4}-byte array (DRAM resident) and over "
                             "a 16 MiB array (L2 resident: what pass 2 is bound by), no hashing"}

        # ---- parity on a bounded prefix, in the bench itself
        if not args.no_parity:
            try:
                from oracle import oracle as orc

                orc.set_threads(host_threads())
                if world == 1:
                    npre = min(n_keys, 20_000_000)
                    _native.call("pb_bloom_clear", filt._h)
                    ctx.set_option("bloom_insert_mode", 2)  # the partitioned path, as in the timed steps
                    try:
                        _native.call("pb_bloom_add_keys", filt._h, pack_keys(keys[:npre]).ref())
                    finally:
                        ctx.set_option("bloom_insert_mode", args.insert_mode)
                    ob = orc.Bloom(m, k)
                    ob.add(orc.pack(orc.uniform_keys(0, npre)))
                    same = bool((filt.bloom_numpy() == ob.bloom).all())
                    parity = {"parity": same, "prefix_keys": npre, "bits_set": filt._cnt_number_bits_set(), "oracle_bits_set": ob.popcount(),
                              "how": "fresh state, key prefix through the timed (partitioned) path, whole 1.2 GB bit array == oracle's"}
                    del ob
                    step()  # back to the full state for the parts below
                else:
                    parity = shard_parity(torch, dist, orc, filt, keys, n_keys, min(n_keys, args.parity_keys), dev)
            except Exception as e:  # noqa: BLE001
                parity = {"parity": None, "error": f"{type(e).__name__}: {str(e)[:300]}"}

        # ---- end to end at N > 1 (before the sharded filter is replaced by config 5's)
        e2e = None
        if world > 1 and not args.no_e2e:
            # every rank: pinned host keys -> H2D -> sharded insert (collective) -> read back its shard's popcount.
            # All ranks first agree that the buffers exist, so nobody is left alone inside a collective.
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
            n_e = min(args.e2e_keys, n_keys)
            host = dbuf = None
            try:
                host = torch.empty((n_e, 16), dtype=torch.uint8, pin_memory=True)
                host.copy_(keys[:n_e])
                dbuf = torch.empty((n_e, 16), dtype=torch.uint8, device=dev)
                ready = 1
            except Exception as ex:  # noqa: BLE001
                ready = 0
                e2e["note"] = f"host staging failed on rank {rank}: {str(ex)[:120]}"
            flag = torch.tensor([ready], dtype=torch.int64, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                try:
                    def e2e_step():
                        filt.clear()
                        dbuf.copy_(host, non_blocking=True)  # H2D of the step's keys on the engine's stream
                        filt.add_many(dbuf)
                        return filt.popcount_local()         # D2H of the step's result

                    e2e_step()
                    barrier()
                    t0 = time.perf_counter()
                    for _ in range(args.steps):
                        bits = e2e_step()
                    barrier()
                    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                    e2e.update(value=n_e * world * args.steps / float(dt.item()), h2d_bytes_per_step=n_e * 16 * world,
                               d2h_bytes_per_step=8 * world, keys_per_step=n_e * world,
                               api="ShardedBloomFilter.add_many per rank from a pinned host array + shard popcount read back",
                               bits_set_rank0_shard=int(bits))
                except Exception as ex:  # noqa: BLE001
                    e2e["note"] = f"end-to-end leg failed: {str(ex)[:160]}"
            elif "note" not in e2e:
                e2e["note"] = "host staging failed on another rank"

        # ---- Count-Min across the ranks: a private 2^20 x 5 table per rank for its own 1e9 Zipf(1.1) keys, then
        # ONE all-reduce into the merged sketch; parity = merged table of a key prefix per rank == the oracle's sketch of
        # all those keys
        if world > 1 and not args.no_parts:
            try:
                from oracle import oracle as orc
                from pyprobables_b200.sharded import ShardedCountMinSketch

                orc.set_threads(host_threads())
                width, depth = 1 << 20, 5
                ranks_t = torch.empty(n_keys, dtype=torch.int64, device=dev)
                ctx.gen_zipf_ranks(rank * n_keys, n_keys, ranks_t.data_ptr(), 1.1, SEED)
                ckeys = torch.empty((n_keys, 16), dtype=torch.uint8, device=dev)
                ctx.gen_rank_keys(ranks_t.data_ptr(), n_keys, ckeys.data_ptr())
                cs = ShardedCountMinSketch(width, depth, device=local, context=ctx)
                cs.add_many(ckeys[: 1 << 24])  # warm-up
                cs.merge()
                cs.local.clear()
                barrier()
                c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                c0.record(stream)
                cs.add_many(ckeys)
                c1.record(stream)
                cs.merge()
                c2.record(stream)
                barrier()
                tt = torch.tensor([c0.elapsed_time(c1), c1.elapsed_time(c2)], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                add_ms, merge_ms = float(tt[0].item()), float(tt[1].item())
                total_ok = cs.elements_added == n_keys * world
                rows_ok = bool((cs.merged.bins_numpy().reshape(depth, -1).astype(np.int64).sum(axis=1) == n_keys * world).all())
                npre = min(n_keys, args.parity_keys)
                cp = ShardedCountMinSketch(width, depth, device=local, context=ctx)
                cp.add_many(ckeys[:npre])
                cp.merge()
                pre = [torch.empty(npre, dtype=torch.int64, device=dev) for _ in range(world)]
                dist.all_gather(pre, ranks_t[:npre].contiguous())
                oc = orc.CMS(width, depth)
                oc.add_parallel(orc.pack(orc.rank_keys(torch.cat(pre).cpu().numpy().astype(np.uint64))))
                same = bool((cp.merged.bins_numpy() == oc.bins).all()) and cp.elements_added == npre * world
                flag = torch.tensor([1 if (same and total_ok and rows_ok) else 0], dtype=torch.int64, device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                v = n_keys * world / ((add_ms + merge_ms) * 1e-3)
                parts = {"cms_sharded": {"value": v, "unit": UNIT, "keys": n_keys * world, "add_ms": add_ms, "merge_ms": merge_ms,
                                         "workload": f"CountMinSketch 2^20 x 5 per rank, {n_keys} Zipf(1.1) adds per rank, one all-reduce merge",
                                         "roofline": {"bound": "l2-atomic", "kernel": "cms_add_fixed16 + NCCL all-reduce of 5 Mi int64",
                                                      "algorithmic_bytes_per_key": 16, "achieved": n_keys / ((add_ms + merge_ms) * 1e-3) * 16 / 1e9,
                                                      "peak": hbm_peak, "unit": "GB/s", "per": "GPU"},
                                         "parity": bool(int(flag.item())), "prefix_keys_per_rank": npre,
                                         "how": "merged table of every rank's key prefix == oracle sketch of all those keys (on every rank); "
                                                "full run: row sums == total adds, elements_added == total"}}
                del ranks_t, ckeys
            except Exception as e:  # noqa: BLE001
                parts = {"cms_sharded": {"error": f"{type(e).__name__}: {str(e)[:300]}"}}

        # ---- BASELINE configs[4]: BloomFilter(1e10, 0.001) range-sharded over 8 GPUs, 8 x 1e9 inserts
        if world > 1 and (world == 8 or args.config5):
            try:
                from oracle import oracle as orc
                from pyprobables_b200.sharded import ShardedBloomFilter

                filt.close()
                f5 = ShardedBloomFilter(10**10, 0.001, device=local, context=ctx, chunk_keys=args.chunk_keys, mode=args.shard_mode,
                                        window_log2=args.window_log2 or 27)

                def step5():
                    f5.clear()
                    f5.add_many(keys)

                step5()
                barrier()
                t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0e.record(stream)
                for _ in range(2):
                    step5()
                f5.join()
                t1e.record(stream)
                barrier()
                t5 = torch.tensor([t0e.elapsed_time(t1e) / 2], dtype=torch.float64, device=dev)
                dist.all_reduce(t5, op=dist.ReduceOp.MAX)
                ms5 = float(t5.item())
                pc = torch.tensor([f5.popcount_local()], dtype=torch.int64, device=dev)
                dist.all_reduce(pc)
                par5 = shard_parity(torch, dist, orc, f5, keys, n_keys, min(n_keys, args.parity_keys), dev)
                k5 = f5.number_hashes
                config5 = {"workload": f"BloomFilter est_elements=1e10 fpr=0.001 (m={f5.number_bits} bits, k={k5}) range-sharded "
                                       f"over {world} GPUs ({f5.plan.windows_per_rank} windows of 2^{f5.plan.window_log2} bits per rank), "
                                       f"{world} x {n_keys} inserts per step",
                           "value": n_keys * world / (ms5 * 1e-3), "unit": UNIT, "ms_per_step": ms5, "bits_set_after_full_insert": int(pc.item()),
                           "roofline": {"bound": "hbm", "algorithmic_bytes_per_key": 16 + 64 * k5,
                                        "achieved": n_keys / (ms5 * 1e-3) * (16 + 64 * k5) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                        "frac": n_keys / (ms5 * 1e-3) * (16 + 64 * k5) / 1e9 / hbm_peak, "per": "GPU"},
                           **par5}
                f5.close()
            except Exception as e:  # noqa: BLE001
                config5 = {"error": f"{type(e).__name__}: {str(e)[:300]}"}

        # ---- end to end through the public API with HOST keys (H2D inside the timed region)
        if world == 1 and not args.no_e2e:
            n_e = min(args.e2e_keys, n_keys)
            hp = _native._vp()
            _native.call("pb_host_alloc", n_e * 16, _native.C.byref(hp))
            host = np.ctypeslib.as_array((_native.C.c_uint8 * (n_e * 16)).from_address(hp.value)).reshape(n_e, 16)
            torch.cuda.synchronize()
            ctx.d2h(hp.value, keys.data_ptr(), n_e * 16)
            filt.clear()
            filt.add_many(host)  # warm the staging buffers
            filt._cnt_number_bits_set()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                filt.clear()
                filt.add_many(host)                # H2D of the step's keys + insert
                bits = filt._cnt_number_bits_set()  # D2H of the step's result (device popcount)
            dt = time.perf_counter() - t0
            e2e = {"value": n_e * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": n_e * 16, "d2h_bytes_per_step": 8,
                   "keys_per_step": n_e, "api": "BloomFilter.add_many(uint8[n,16] pinned host array) + number-of-bits-set read back",
                   "bits_set": int(bits)}
            del host
            _native.call("pb_host_free", hp)
            step()  # the parts below expect the filter to hold all the keys again
        if world == 1 and not args.no_parts:
            parts = run_parts(args, torch, pb, ctx, stream, timer, filt, keys, hbm_peak)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline_port(m, k)
        except Exception as e:
            cpu = {"unavailable": str(e)[:200]}

    if rank == 0:
        exchange = ("" if world == 1 else
                    f"; bit array range-sharded over {world} GPUs, window-partitioned bit indices exchanged "
                    + {"p2p": "through CUDA-IPC mailboxes in peer memory (copy-engine pushes over NVLink, flag protocol; NCCL is control plane only)",
                       "route": "as u64 indices with NCCL all-to-all-v", "gather": "by all-gathering the keys (NCCL)"}[args.shard_mode])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": (f"BloomFilter est_elements={EST_PER_GPU * world:.0e} fpr={FPR} (m={m} bits, k={k}); "
                                    f"step = clear + batch insert of {n_keys} 16-byte keys per GPU resident in HBM" + exchange),
                       "keys_per_gpu_per_step": n_keys, "key_bytes": 16, "bitmap_bytes_per_gpu": int(bitmap_bytes),
                       "l2_policy": f"inputs larger than L2: {n_keys * 16} key bytes + {int(bitmap_bytes)} bitmap bytes per step vs 126 MB L2",
                       "insert_mode": int(ctx.get_option("bloom_insert_mode"))},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "atomic_roofline": atomic,
            "cpu_baseline": cpu, "bits_set_after_timed_region": int(setbits), "parity": parity, "parts": parts,
        }
        if config5 is not None:
            line["config5"] = config5
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--keys", type=int, default=10**9, help="keys per GPU per step")
    ap.add_argument("--e2e-keys", type=int, default=1 << 27)
    ap.add_argument("--ref-keys", type=int, default=100_000, help="keys per step of the pure-Python reference arm")
    ap.add_argument("--parity-keys", type=int, default=1_000_000, help="N > 1: key prefix per rank of the in-bench shard parity")
    ap.add_argument("--ragged-keys", type=int, default=250_000_000)
    ap.add_argument("--cuckoo-log2", type=int, default=28, help="log2 of the Cuckoo capacity of the parts (BASELINE config: 28)")
    ap.add_argument("--chunk-keys", type=int, default=1 << 27)
    ap.add_argument("--shard-mode", default="p2p", choices=["p2p", "route", "gather"])
    ap.add_argument("--insert-mode", type=int, default=0, help="bloom_insert_mode: 0 auto, 1 direct RED, 2 partitioned")
    ap.add_argument("--window-log2", type=int, default=0, help="bloom_window_log2_bits override")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (repeatable), e.g. bloom_part_tile=256")
    ap.add_argument("--config5", action="store_true", help="N > 1: also run BASELINE configs[4] (default at N = 8)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-parts", action="store_true")
    ap.add_argument("--no-port", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
