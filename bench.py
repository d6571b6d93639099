#!/usr/bin/env python
"""bench.py -- headline benchmark of the hash-then-scatter hot path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm   (CUDA engine)
  python bench.py --impl reference [--gpus N] [--steps K] ...    reference arm (pure-Python pyprobables on host cores)

N = 1: BASELINE configs[1] -- BloomFilter(est_elements=1e9, fpr=0.01) (m = 9 585 058 424 bits, k = 7),
one step = clear the filter and batch-insert 1e9 synthetic 16-byte keys that already sit in HBM.
N > 1 (torchrun, one rank per GPU): configs[4]-style weak scaling -- ONE logical filter of N*1e9 elements,
bit array range-sharded over the ranks, every rank inserts its own 1e9 keys through route -> NCCL
all-to-all -> apply.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""

from __future__ import annotations

import argparse
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


# ----------------------------------------------------------------------------- reference arm
def run_reference(args) -> None:
    """the reference's own implementation of the path on the host: pure-Python pyprobables BloomFilter.add
    from baseline/_ref (unmodified, its public API).  The reference is single-threaded by construction, so
    cores = 1; the compiled oracle port on all cores is reported beside it as `port_all_cores`."""
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

        orc.set_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
        fpr32, k, m, _ = orc.bloom_params(est, FPR)
        ob = orc.Bloom(m, k)
        sample = max(sample, 20_000_000)
        cores = orc.num_threads()

        def step(i):
            keys = orc.pack(orc.uniform_keys(i * sample, sample))
            t0 = time.perf_counter()
            ob.add(keys)
            return time.perf_counter() - t0

        line["note"] = f"baseline/_ref unavailable ({type(e).__name__}); timed the oracle port"
    for i in range(args.warmup):
        step(i)
    total = sum(step(args.warmup + i) for i in range(args.steps))
    value = sample * args.steps / total
    line.update(value=value, ms_per_step=1e3 * total / args.steps,
                config={"workload": f"BloomFilter est_elements={est:.0e} fpr={FPR} (m={m} bits, k={k}); batch insert of 16-byte keys; "
                                    f"each step = a bounded sample of {sample} keys of that workload on the host",
                        "key_bytes": 16, "l2_policy": "bit array (>= 1.2 GB) far larger than any cache"},
                cpu_baseline={"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                              "sample": f"{sample} keys/step x {args.steps} steps, host cpu_count={os.cpu_count()}"},
                e2e={"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                gpu_launches=0)
    if kind == "reference" and not args.no_port:
        try:
            from oracle import oracle as orc

            orc.set_threads(os.cpu_count() or 1)
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

    orc.set_threads(os.cpu_count() or 1)
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
            "sample": f"{n} of the workload's keys into the full-size {m}-bit array, C oracle + OpenMP, host cpu_count={os.cpu_count()}"}


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
                                      mode=args.shard_mode)
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
        ctxs = [ctx] + ([filt._ctx_part] if getattr(filt, "_ctx_part", None) is not None else [])
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

        # size-independent sanity of the timed state (full parity lives in tests/): inserted keys are members
        probe_n = min(n_keys, 1 << 20)
        if world == 1:
            ok = bool(filt.check_many(keys[:probe_n]).all())
            setbits = filt._cnt_number_bits_set()
        else:
            ok = bool(filt.check_many(keys[:probe_n]).all())
            pc = torch.tensor([filt.popcount_local()], dtype=torch.int64, device=dev)
            dist.all_reduce(pc)
            setbits = int(pc.item())
        assert ok, "inserted keys are not members: the timed state is wrong"

        # ---- roofline of the dominant kernel (device time from CUDA events on the launching stream)
        roof = None
        atomic = None
        if ktimes:
            name, (n_launch, tot_ms) = max(ktimes.items(), key=lambda kv: kv[1][1])
            keys_per_launch = n_keys * args.steps / max(n_launch, 1)
            # algorithmic bytes of a Bloom insert = key + k random 32-B-sector read-modify-writes
            algo_per_key = 16 + 64 * k
            # what the kernels of this design actually have to move per key: pass 1 reads the key and
            # writes k 4-byte indices, pass 2 reads them back; the bitmap is read+written once per chunk
            n_chunks = max(1, ktimes.get("bloom_apply_windows", (1, 0))[0] // max(args.steps, 1))
            design_per_key = 16 + 8 * k + 2.0 * bitmap_bytes * n_chunks / n_keys if "bloom_part" in ktimes else algo_per_key
            avg_ms = tot_ms / max(n_launch, 1)
            achieved = keys_per_launch * algo_per_key / (avg_ms * 1e-3) / 1e9
            traffic, traffic_src = None, None
            try:  # DRAM bytes per launch of that kernel from the committed ncu --set full capture (scaled to this launch size)
                tj = json.loads((ROOT / "profiles" / "r1_traffic.json").read_text())
                tk = tj.get(name)
                if tk:
                    per_key = (tk["dram_bytes_read"] + tk["dram_bytes_write"]) / tk.get("keys_per_launch", tj["keys_per_launch"])
                    traffic = per_key * keys_per_launch
                    traffic_src = f"profiles/r1_traffic.json ({tk['kernel']}): {per_key:.1f} DRAM B/key x {keys_per_launch:.0f} keys per launch"
            except Exception:
                pass
            roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": keys_per_launch * algo_per_key,
                    "algorithmic_bytes_per_key": algo_per_key, "keys_per_launch": keys_per_launch,
                    "launches": n_launch, "avg_launch_ms": avg_ms, "kernel_share_of_step": tot_ms / ms,
                    "kernels": {kn: {"launches": v[0], "total_ms": v[1]} for kn, v in ktimes.items()},
                    "note": ("achieved = SURVEY 8(d) algorithmic bytes (16+64k per key, the random-atomic model) x keys per launch / "
                             "average launch time of the dominant kernel; pass 2 (bloom_apply_windows) runs concurrently on a second "
                             "stream, so the kernel times overlap.  The partitioned design moves far fewer DRAM bytes than the model "
                             "(design_bytes_per_key), which is how it exceeds the random-atomic ceiling."),
                    "design_bytes_per_key": design_per_key,
                    "design_achieved_GBps": value / world * design_per_key / 1e9,
                    "step_model": {"survey_bytes_per_key": algo_per_key,
                                   "achieved_GBps": value / world * algo_per_key / 1e9,
                                   "frac": value / world * algo_per_key / 1e9 / hbm_peak}}
        if rank == 0 and world == 1 and not args.no_micro:
            words = (bitmap_bytes + 3) // 4
            n_at = 1 << 28
            ms_or = ctx.microbench(words, n_at, 0, 3)
            ms_cp = ctx.microbench(words, 0, 3, 3)
            ceil_rate = n_at / (ms_or * 1e-3)
            atomic = {"random_red_or_per_s": ceil_rate, "bloom_bit_updates_per_s": value * k,
                      "frac_of_random_atomic_ceiling": value * k / ceil_rate,
                      "copy_GBps_same_run": 2 * words * 4 / (ms_cp * 1e-3) / 1e9,
                      "how": f"{n_at} RED.OR.b32 at pre-generated uniform indices over a {words * 4}-byte array, no hashing"}

        # ---- end to end through the public API with HOST keys (H2D inside the timed region)
        e2e = None
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
            _native.call("pb_host_free", hp)
        elif world > 1 and not args.no_e2e:
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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline_port(m, k)
        except Exception as e:
            cpu = {"unavailable": str(e)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": (f"BloomFilter est_elements={EST_PER_GPU * world:.0e} fpr={FPR} (m={m} bits, k={k}); "
                                    f"step = clear + batch insert of {n_keys} 16-byte keys per GPU resident in HBM"
                                    + ("" if world == 1 else f"; bit array range-sharded over {world} GPUs, NCCL all-to-all of window-partitioned bit indices ({args.shard_mode})")),
                       "keys_per_gpu_per_step": n_keys, "key_bytes": 16, "bitmap_bytes_per_gpu": int(bitmap_bytes),
                       "l2_policy": f"inputs larger than L2: {n_keys * 16} key bytes + {int(bitmap_bytes)} bitmap bytes per step vs 126 MB L2",
                       "insert_mode": int(ctx.get_option("bloom_insert_mode"))},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "atomic_roofline": atomic,
            "cpu_baseline": cpu, "bits_set_after_timed_region": int(setbits),
        }
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
    ap.add_argument("--chunk-keys", type=int, default=1 << 27)
    ap.add_argument("--shard-mode", default="p2p", choices=["p2p", "route", "gather"])
    ap.add_argument("--insert-mode", type=int, default=0, help="bloom_insert_mode: 0 auto, 1 direct RED, 2 partitioned")
    ap.add_argument("--window-log2", type=int, default=0, help="bloom_window_log2_bits override")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (repeatable), e.g. bloom_part_tile=256")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-port", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
