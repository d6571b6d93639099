"""Multi-GPU Bloom filter and Count-Min sketch: one process per GPU over torch.distributed (NCCL on the
B200 box; gloo in the CPU tests of the host-side logic).

Bloom: the bit array of ONE logical filter is range-sharded -- rank g owns bits
[g*S, min((g+1)*S, m)) with S = ceil(m/G) rounded up to a multiple of 32 (whole words).  Keys are data
parallel.  add_many per chunk of local keys:
    route    hash k seeds -> global bit index (h % m, same arithmetic as the single-GPU filter) ->
             owner = idx / S -> append idx to the owner's slot              (one CUDA kernel)
    exchange all-to-all-v of the u64 indices over NVLink                      (NCCL send/recv group)
    apply    RED.OR of the received indices into the local shard            (one CUDA kernel)
The concatenation of the shards is bit-identical to the single-GPU (and the reference's) bit array.
check_many all-gathers the probe keys, lets every rank AND the bits it owns (bits of other shards are
neutral) and all-reduces the partial answers with MIN.

Count-Min: addition commutes, so every rank keeps a private full table for its share of the stream
and merge() sums them with one all-reduce -- the reference's join() (countminsketch.py:356-399) across
ranks.
"""

from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _native
from .bloom import optimized_params
from .keys import pack_keys


@dataclass(frozen=True)
class ShardPlan:
    """who owns which bits of an m-bit filter split over `world` ranks.  Shards are whole *windows* of
    2^window_log2 bits (the unit the partitioned insert bins bit indices by), so a window never straddles
    two GPUs: rank g owns windows [g*windows_per_rank, (g+1)*windows_per_rank)."""

    num_bits: int
    world: int
    shard_bits: int
    window_log2: int = 5
    windows_per_rank: int = 0

    MAX_WINDOWS = 512  # kMaxWindows2 in csrc/pb_bloom_part.cuh

    @staticmethod
    def make(num_bits: int, world: int, window_log2: int = 27) -> "ShardPlan":
        if world < 1 or num_bits < 1:
            raise ValueError("world and num_bits must be >= 1")
        wl = max(10, min(int(window_log2), 31))
        nwin = lambda w: -(-num_bits // (1 << w))
        while wl > 10 and nwin(wl) < 4 * world:  # small filters: a few windows per rank keeps the split even
            wl -= 1
        while wl < 31 and -(-nwin(wl) // world) * world > ShardPlan.MAX_WINDOWS:
            wl += 1
        wps = -(-nwin(wl) // world)
        if wps * world > ShardPlan.MAX_WINDOWS:
            raise ValueError("filter too large for the windowed shard plan")
        return ShardPlan(int(num_bits), int(world), int(wps << wl), int(wl), int(wps))

    @property
    def total_windows(self) -> int:
        return self.windows_per_rank * self.world

    def active_windows(self, rank: int) -> int:
        lo, hi = self.bounds(rank)
        return -(-(hi - lo) // (1 << self.window_log2))

    def bounds(self, rank: int) -> tuple[int, int]:
        lo = min(rank * self.shard_bits, self.num_bits)
        hi = min(lo + self.shard_bits, self.num_bits)
        return lo, hi

    def owner(self, idx):
        return idx // self.shard_bits

    def shard_nbytes(self, rank: int) -> int:
        lo, hi = self.bounds(rank)
        return (hi - lo + 7) // 8


def exchange_counts(counts_to_peers, group=None):
    """all-to-all of one int64 per peer: what I send to each rank -> what I receive from each rank"""
    import torch
    import torch.distributed as dist

    send = torch.as_tensor(counts_to_peers, dtype=torch.int64).clone()
    if send.device.type == "cpu" and dist.get_backend(group) == "nccl":
        send = send.cuda()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv


def exchange_indices(send_segments, recv_counts, group=None):
    """all-to-all-v: send_segments[d] (1-D tensor, may be empty) goes to rank d; returns one contiguous
    tensor holding what every rank sent to me (rank order) and the per-source offsets."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    assert len(send_segments) == world
    recv_counts = [int(x) for x in recv_counts]
    proto = send_segments[0]
    recv = torch.empty(sum(recv_counts), dtype=proto.dtype, device=proto.device)
    offs = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
    outs = [recv[offs[r] : offs[r + 1]] for r in range(world)]
    if dist.get_backend(group) == "gloo":
        # gloo has no all_to_all with ragged views: pairwise isend/irecv
        me = dist.get_rank(group)
        outs[me].copy_(send_segments[me])
        reqs = []
        for r in range(world):
            if r == me:
                continue
            if send_segments[r].numel():
                reqs.append(dist.isend(send_segments[r].contiguous(), dst=dist.get_global_rank(group, r) if group else r, group=group))
            if recv_counts[r]:
                reqs.append(dist.irecv(outs[r], src=dist.get_global_rank(group, r) if group else r, group=group))
        for q in reqs:
            q.wait()
    else:
        dist.all_to_all(outs, [s.contiguous() for s in send_segments], group=group)
    return recv, offs


def torch_stream_context(device: int) -> _native.Context:
    """a Context whose kernels run on torch's current stream of `device`"""
    import torch

    handle = torch.cuda.current_stream(device).cuda_stream
    return _native.Context(device, stream=handle if handle else 0x1)


class _DevView:
    """zero-copy torch view of library-owned device memory (via __cuda_array_interface__)"""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_view(ptr: int, n: int, typestr: str, device: int):
    import torch

    return torch.as_tensor(_DevView(ptr, n, typestr), device=f"cuda:{device}")


class ShardedBloomFilter:
    """One logical BloomFilter(est_elements, false_positive_rate) range-sharded over the ranks of `group`.
    Every rank constructs it with the same arguments; `add_many` takes each rank's own keys (device
    resident uint8[n,16] tensors or anything pack_keys accepts)."""

    def __init__(self, est_elements, false_positive_rate, group=None, device=None, context=None, chunk_keys: int = 1 << 26,
                 mode: str = "fused"):
        import torch
        import torch.distributed as dist

        self._dist, self._torch = dist, torch
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._fpr, self._k, self._m = optimized_params(est_elements, false_positive_rate)
        self._est = est_elements
        self.plan = ShardPlan.make(self._m, self.world)
        self.lo, self.hi = self.plan.bounds(self.rank)
        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        # run on torch's current stream so kernels, NCCL collectives and tensor ops are ordered without
        # host synchronization (handle 0 is the legacy default stream = cudaStreamLegacy, 0x1)
        self._ctx = context if context is not None else torch_stream_context(self.device)
        if mode not in ("p2p", "p2p_direct", "fused", "route", "gather"):
            raise ValueError("mode must be 'p2p', 'p2p_direct', 'fused', 'route' or 'gather'")
        self.mode = mode
        self.chunk_keys = int(chunk_keys)
        h = C.c_void_p()
        if self.hi > self.lo:
            _native.call("pb_bloom_create_shard", self._ctx.handle, self._m, self._k, self.lo, self.hi, C.byref(h))
        self._h = h if self.hi > self.lo else None
        self._els_added = 0
        self._send = None
        self._counts = None
        self._fused_bufs = None
        self._s_part = self._s_comm = self._ctx_part = None
        self._p2p = None

    # -- properties in the reference's vocabulary
    @property
    def number_bits(self) -> int:
        return self._m

    @property
    def number_hashes(self) -> int:
        return self._k

    @property
    def false_positive_rate(self) -> float:
        return self._fpr

    @property
    def estimated_elements(self) -> int:
        return self._est

    @property
    def elements_added(self) -> int:
        """keys this rank has contributed (sum over ranks = the logical filter's elements_added)"""
        return self._els_added

    def close(self) -> None:
        if getattr(self, "_p2p", None) is not None and _native._lib is not None:
            _native._lib.pb_p2p_destroy(self._p2p["h"])
            self._p2p = None
        if getattr(self, "_h", None) is not None and _native._lib is not None:
            _native._lib.pb_bloom_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self) -> None:
        self._els_added = 0
        if self._h is not None:
            _native.call("pb_bloom_clear", self._h)

    def shard_numpy(self) -> np.ndarray:
        """host copy of this rank's slice of the bit array (bytes [lo/8, ...))"""
        n = self.plan.shard_nbytes(self.rank)
        out = np.empty(n, dtype=np.uint8)
        if n:
            _native.call("pb_bloom_download", self._h, C.c_void_p(out.ctypes.data), n)
        return out

    def popcount_local(self) -> int:
        if self._h is None:
            return 0
        n = C.c_uint64()
        _native.call("pb_bloom_popcount", self._h, C.byref(n))
        return n.value

    # -- hot path
    def _ensure_buffers(self, chunk: int, worst_case: bool = False):
        torch = self._torch
        # every owner can receive all k indices of a chunk in the worst case; slots are sized for the uniform
        # expectation with generous slack unless worst_case asks for the full bound
        slot = int(chunk * self._k / self.world * 1.25) + 4096
        if self.world == 1 or worst_case:
            slot = chunk * self._k
        if self._send is None or self._send.numel() < slot * self.world:
            self._send = torch.empty(slot * self.world, dtype=torch.int64, device=f"cuda:{self.device}")
            self._counts = torch.zeros(64, dtype=torch.int64, device=f"cuda:{self.device}")
        self._slot = slot

    def _device_keys(self, keys):
        torch = self._torch
        if isinstance(keys, torch.Tensor):
            if not keys.is_cuda:
                keys = keys.to(f"cuda:{self.device}", non_blocking=True)
            return keys
        if isinstance(keys, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(keys)).to(f"cuda:{self.device}")
        raise TypeError("sharded filters take fixed-width uint8[n, L] key arrays (numpy or torch)")

    def add_many(self, keys) -> None:
        """BloomFilter.add (bloom.py:234-250) for this rank's keys; collective: every rank must call it
        (an empty batch is fine)"""
        torch, dist = self._torch, self._dist
        t = self._device_keys(keys)
        n = int(t.shape[0])
        if self.mode == "gather":
            self._add_gather(t)
            self._els_added += n
            return
        if t.shape[1] != 16:
            raise TypeError("fused/route modes take 16-byte keys; use mode='gather' for other widths")
        if self.mode == "fused":
            self._add_fused(t)
            self._els_added += n
            return
        if self.mode in ("p2p", "p2p_direct"):
            self._add_p2p(t)
            self._els_added += n
            return
        self._add_route(t, self.chunk_keys, worst_case=False)
        self._els_added += n

    def _add_route(self, t, chunk_keys: int, worst_case: bool) -> None:
        """u64 global indices binned by owner + all-to-all-v with counts + RED apply.  worst_case sizes every
        owner's slot for ALL indices of a chunk, which makes the path exact for any key distribution."""
        torch, dist = self._torch, self._dist
        n = int(t.shape[0])
        n_chunks = torch.tensor([-(-n // chunk_keys)], dtype=torch.int64, device=t.device)
        dist.all_reduce(n_chunks, op=dist.ReduceOp.MAX, group=self.group)  # all ranks walk the same number of chunks
        for ci in range(int(n_chunks.item())):
            lo = min(ci * chunk_keys, n)
            hi = min(lo + chunk_keys, n)
            self._ensure_buffers(max(hi - lo, 1), worst_case)
            self._counts.zero_()
            if hi > lo:
                kb = pack_keys(t[lo:hi], sync=False)
                _native.call("pb_bloom_route_keys", self._ctx.handle, kb.ref(), self._m, self._k, self.plan.shard_bits,
                             self.world, C.c_void_p(self._send.data_ptr()), self._slot, C.c_void_p(self._counts.data_ptr()))
            counts = self._counts[: self.world].clone()
            if int(counts.max().item()) > self._slot:
                raise RuntimeError("routing slot overflow: keys are too skewed for the slot size; lower chunk_keys")
            recv_counts = exchange_counts(counts, self.group)
            c_host = counts.tolist()
            segs = [self._send[d * self._slot : d * self._slot + c_host[d]] for d in range(self.world)]
            recv, _ = exchange_indices(segs, recv_counts.tolist(), self.group)
            if recv.numel() and self._h is not None:
                _native.call("pb_bloom_add_bit_indices", self._h, C.c_void_p(recv.data_ptr()), recv.numel())

    # -- fused route + partition (default): see pb_bloom_partition_keys in include/pb200.h
    def _fused_buffers(self, chunk: int):
        torch = self._torch
        plan = self.plan
        if self._fused_bufs is not None and self._fused_bufs["chunk"] >= chunk:
            return self._fused_bufs
        W = plan.total_windows
        slack = C.c_uint64()
        _native.call("pb_bloom_partition_slack", self._ctx.handle, chunk, self._k, W, C.byref(slack))
        expect = chunk * self._k * (1 << plan.window_log2) / self._m
        cap = int(expect * 1.03 + 6.0 * math.sqrt(expect + 1.0)) + slack.value
        cap = (cap + 3) // 4 * 4
        if cap * W > 0xFFFFFFF0:
            raise ValueError("chunk_keys too large for the staging layout; lower chunk_keys")
        dev = f"cuda:{self.device}"
        i32 = torch.int32
        b = {"chunk": chunk, "cap": cap,
             # two halves of everything: pass 1 of chunk c+1, the all-to-all of chunk c and pass 2 of chunk c-1 overlap
             "send": [torch.empty(W * cap, dtype=i32, device=dev) for _ in range(2)],
             "recv": [torch.empty(W * cap, dtype=i32, device=dev) for _ in range(2)],
             "scur": [torch.zeros(W, dtype=i32, device=dev) for _ in range(2)],
             "rcur": [torch.zeros(W, dtype=i32, device=dev) for _ in range(2)],
             "ovf": torch.empty(1 << 22, dtype=torch.int64, device=dev), "ovf_n": torch.zeros(1, dtype=torch.int64, device=dev),
             "ev_part": [torch.cuda.Event() for _ in range(2)], "ev_comm": [torch.cuda.Event() for _ in range(2)],
             "ev_apply": [torch.cuda.Event() for _ in range(2)]}
        if self._s_part is None:
            self._s_part = torch.cuda.Stream(device=self.device)
            self._s_comm = torch.cuda.Stream(device=self.device)
            self._ctx_part = _native.Context(self.device, stream=self._s_part.cuda_stream)
        self._fused_bufs = b
        return b

    def _add_fused(self, t) -> None:
        """three-stage pipeline over chunks of keys, one CUDA stream per stage:
             pass 1 (hash + bin by global window)  ->  NCCL all-to-all of the window lists  ->  pass 2 (RED.OR into my shard)
        Pass 2 runs on the stream the filter was created on, so everything the caller does next is ordered after it."""
        torch, dist = self._torch, self._dist
        plan = self.plan
        n = int(t.shape[0])
        n_chunks = torch.tensor([-(-n // self.chunk_keys)], dtype=torch.int64, device=t.device)
        dist.all_reduce(n_chunks, op=dist.ReduceOp.MAX, group=self.group)
        n_chunks = int(n_chunks.item())
        if n_chunks == 0:
            return
        b = self._fused_buffers(min(self.chunk_keys, max(n, 1)) if n_chunks == 1 else self.chunk_keys)
        main = torch.cuda.current_stream(self.device)
        b["ovf_n"].zero_()
        self._s_part.wait_stream(main)  # keys and the zeroed overflow counter are ready
        self._s_comm.wait_stream(main)
        act = plan.active_windows(self.rank)
        for ci in range(n_chunks):
            h = ci & 1
            lo = min(ci * self.chunk_keys, n)
            hi = min(lo + self.chunk_keys, n)
            kb = pack_keys(t[lo:hi], sync=False) if hi > lo else pack_keys(t[:0], sync=False)
            with torch.cuda.stream(self._s_part):
                if ci >= 2:
                    self._s_part.wait_event(b["ev_comm"][h])  # the all-to-all that read this send half is done
                _native.call("pb_bloom_partition_keys", self._ctx_part.handle, kb.ref(), self._m, self._k, plan.window_log2,
                             plan.total_windows, b["cap"], C.c_void_p(b["send"][h].data_ptr()), C.c_void_p(b["scur"][h].data_ptr()),
                             C.c_void_p(b["ovf"].data_ptr()), b["ovf"].numel(), C.c_void_p(b["ovf_n"].data_ptr()))
                b["ev_part"][h].record(self._s_part)
            with torch.cuda.stream(self._s_comm):
                self._s_comm.wait_event(b["ev_part"][h])
                if ci >= 2:
                    self._s_comm.wait_event(b["ev_apply"][h])  # pass 2 that read this receive half is done
                dist.all_to_all_single(b["rcur"][h], b["scur"][h], group=self.group)
                dist.all_to_all_single(b["recv"][h], b["send"][h], group=self.group)
                b["ev_comm"][h].record(self._s_comm)
            main.wait_event(b["ev_comm"][h])
            if self._h is not None and act > 0:
                _native.call("pb_bloom_apply_window_lists", self._h, C.c_void_p(b["recv"][h].data_ptr()),
                             C.c_void_p(b["rcur"][h].data_ptr()), self.world, plan.windows_per_rank, act, b["cap"], plan.window_log2)
            b["ev_apply"][h].record(main)
        main.wait_stream(self._s_part)
        main.wait_stream(self._s_comm)
        # indices that did not fit their window list (heavily duplicated keys): exact slow path, all ranks together
        worst = b["ovf_n"].clone()
        dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=self.group)
        worst = int(worst.item())
        if worst > b["ovf"].numel():
            # more strays than the list holds (e.g. one key repeated millions of times): OR is idempotent, so
            # simply run the whole batch again through the exact u64 route with worst-case slots
            self._add_route(t, min(self.chunk_keys, 1 << 20), worst_case=True)
        elif worst > 0:
            self._route_indices(b["ovf"][: int(b["ovf_n"].item())])

    # -- fused compute + exchange over NVLink peer memory: see pb_p2p_* in include/pb200.h
    def _p2p_setup(self, chunk: int):
        torch, dist = self._torch, self._dist
        if self._p2p is not None and self._p2p["chunk"] >= chunk:
            return self._p2p
        if self._p2p is not None:
            raise ValueError("the P2P mailbox was sized for smaller chunks; construct the filter with the final chunk_keys")
        plan = self.plan
        W = plan.total_windows
        if self._s_part is None:
            self._s_part = torch.cuda.Stream(device=self.device)
            self._s_comm = torch.cuda.Stream(device=self.device)
            self._ctx_part = _native.Context(self.device, stream=self._s_part.cuda_stream)
        slack = C.c_uint64()
        _native.call("pb_bloom_partition_slack", self._ctx.handle, chunk, self._k, W, C.byref(slack))
        expect = chunk * self._k * (1 << plan.window_log2) / self._m
        cap = int(expect * 1.03 + 6.0 * math.sqrt(expect + 1.0)) + slack.value
        cap = (cap + 3) // 4 * 4
        h = C.c_void_p()
        _native.call("pb_p2p_create", self._ctx_part.handle, self.world, self.rank, plan.windows_per_rank, cap, C.byref(h))
        _native.call("pb_p2p_set_direct", h, 1 if self.mode == "p2p_direct" else 0)
        dev = f"cuda:{self.device}"
        mine = np.zeros(64, dtype=np.uint8)
        _native.call("pb_p2p_export", h, C.c_void_p(mine.ctypes.data))
        allh = torch.empty((self.world, 64), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, torch.from_numpy(mine).to(dev), group=self.group)
        handles = np.ascontiguousarray(allh.cpu().numpy())
        _native.call("pb_p2p_connect", h, C.c_void_p(handles.ctypes.data))
        dist.barrier(group=self.group)  # every mailbox is mapped everywhere before anyone writes
        self._p2p = {"chunk": chunk, "cap": cap, "h": h,
                     "ovf": torch.empty(1 << 22, dtype=torch.int64, device=dev), "ovf_n": torch.zeros(1, dtype=torch.int64, device=dev)}
        return self._p2p

    def _add_p2p(self, t) -> None:
        """pass 1 stores into the owners' mailboxes over NVLink (stream s_part); pass 2 on the filter's stream"""
        torch, dist = self._torch, self._dist
        plan = self.plan
        n = int(t.shape[0])
        n_chunks = torch.tensor([-(-n // self.chunk_keys)], dtype=torch.int64, device=t.device)
        dist.all_reduce(n_chunks, op=dist.ReduceOp.MAX, group=self.group)
        n_chunks = int(n_chunks.item())
        if n_chunks == 0:
            return
        b = self._p2p_setup(self.chunk_keys)
        main = torch.cuda.current_stream(self.device)
        b["ovf_n"].zero_()
        self._s_part.wait_stream(main)
        act = plan.active_windows(self.rank)
        for ci in range(n_chunks):
            lo = min(ci * self.chunk_keys, n)
            hi = min(lo + self.chunk_keys, n)
            kb = pack_keys(t[lo:hi], sync=False) if hi > lo else pack_keys(t[:0], sync=False)
            _native.call("pb_p2p_partition_send", b["h"], kb.ref(), self._m, self._k, plan.window_log2,
                         C.c_void_p(b["ovf"].data_ptr()), b["ovf"].numel(), C.c_void_p(b["ovf_n"].data_ptr()))
            _native.call("pb_p2p_apply", b["h"], self._h, act, plan.window_log2)
        main.wait_stream(self._s_part)  # (the last pass 2 on `main` already waited for every source's copies)
        worst = b["ovf_n"].clone()
        dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=self.group)
        worst = int(worst.item())
        if worst > b["ovf"].numel():
            self._add_route(t, min(self.chunk_keys, 1 << 20), worst_case=True)
        elif worst > 0:
            self._route_indices(b["ovf"][: int(b["ovf_n"].item())])

    def _route_indices(self, idx) -> None:
        """send global bit indices (int64 tensor) to their owners and OR them in (collective)"""
        torch = self._torch
        owner = torch.div(idx, self.plan.shard_bits, rounding_mode="floor")
        order = torch.argsort(owner)
        idx = idx[order].contiguous()
        counts = torch.bincount(owner, minlength=self.world)[: self.world]
        recv_counts = exchange_counts(counts, self.group)
        offs = [0] + torch.cumsum(counts, 0).tolist()
        segs = [idx[offs[d] : offs[d + 1]] for d in range(self.world)]
        recv, _ = exchange_indices(segs, recv_counts.tolist(), self.group)
        if recv.numel() and self._h is not None:
            _native.call("pb_bloom_add_bit_indices", self._h, C.c_void_p(recv.data_ptr()), recv.numel())

    def _add_gather(self, t) -> None:
        torch, dist = self._torch, self._dist
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        sizes = [torch.empty_like(n) for _ in range(self.world)]
        dist.all_gather(sizes, n, group=self.group)
        sizes = [int(s.item()) for s in sizes]
        mx = max(sizes)
        if mx == 0:
            return
        pad = torch.zeros((mx, t.shape[1]), dtype=torch.uint8, device=t.device)
        pad[: t.shape[0]] = t
        allk = torch.empty((self.world, mx, t.shape[1]), dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(allk, pad, group=self.group)
        if self._h is None:
            return
        for r in range(self.world):
            if sizes[r]:
                _native.call("pb_bloom_add_keys", self._h, pack_keys(allk[r, : sizes[r]], sync=False).ref())
        self._ctx.synchronize()

    def check_many(self, keys):
        """BloomFilter.check for this rank's keys -> bool tensor; collective"""
        torch, dist = self._torch, self._dist
        t = self._device_keys(keys)
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        sizes = [torch.empty_like(n) for _ in range(self.world)]
        dist.all_gather(sizes, n, group=self.group)
        sizes = [int(s.item()) for s in sizes]
        mx = max(sizes)
        if mx == 0:
            return torch.zeros(0, dtype=torch.bool, device=t.device)
        pad = torch.zeros((mx, t.shape[1]), dtype=torch.uint8, device=t.device)
        pad[: t.shape[0]] = t
        allk = torch.empty((self.world, mx, t.shape[1]), dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(allk, pad, group=self.group)
        partial = torch.ones((self.world, mx), dtype=torch.uint8, device=t.device)
        if self._h is not None:
            for r in range(self.world):
                if sizes[r]:
                    _native.call("pb_bloom_check_keys", self._h, pack_keys(allk[r, : sizes[r]], sync=False).ref(),
                                 C.c_void_p(partial[r].data_ptr()), 1)
            self._ctx.synchronize()
        dist.all_reduce(partial, op=dist.ReduceOp.MIN, group=self.group)
        return partial[self.rank, : t.shape[0]].bool()


class ShardedCountMinSketch:
    """data-parallel Count-Min: a private table per rank for its share of the stream; merge() gives every
    rank the sketch of the whole stream = rank 0's table joined with rank 1..G-1's in rank order, each join
    being the reference's CountMinSketch.join (countminsketch.py:356-399) run by the device kernel."""

    def __init__(self, width, depth, group=None, device=None, context=None):
        import torch
        import torch.distributed as dist

        from .countminsketch import CountMinSketch

        self._dist, self._torch = dist, torch
        self.group = group
        self.world = dist.get_world_size(group)
        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        ctx = context if context is not None else torch_stream_context(self.device)
        self.local = CountMinSketch(width=width, depth=depth, device=self.device, context=ctx)

    def add_many(self, keys, num_els=1) -> None:
        self.local.add_many(keys, num_els)

    def merge(self) -> None:
        torch, dist = self._torch, self._dist
        c = self.local
        n = c.width * c.depth
        mine = device_view(c.device_ptr(), n, "<i4", self.device)
        allt = torch.empty((self.world, n), dtype=torch.int32, device=mine.device)
        dist.all_gather_into_tensor(allt, mine, group=self.group)
        mine.copy_(allt[0])
        for r in range(1, self.world):
            _native.call("pb_cms_join_buffer", c._h, C.c_void_p(allt[r].data_ptr()), n)
        c._ctx.synchronize()
        ea = torch.tensor([c.elements_added], dtype=torch.int64, device=mine.device)
        dist.all_reduce(ea, op=dist.ReduceOp.SUM, group=self.group)
        c._elements_added = int(ea.item())

    def check_many(self, keys):
        return self.local.check_many(keys)
