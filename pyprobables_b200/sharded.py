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
check_many sends the k bit indices of every probe key to the owners of those bits, gets one byte per index
back and ANDs them at the source.

Count-Min: addition commutes, so every rank keeps a private full table for its share of the stream
and merge() sums them with one all-reduce into a separate merged table -- the reference's join()
(countminsketch.py:356-399) across ranks.
"""

from __future__ import annotations

import ctypes as C
import sys
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
    if dist.get_backend(group) == "gloo" and proto.is_cuda:
        # gloo moves host memory only (CPU tests; several ranks sharing one GPU): stage through the host
        recv, offs = exchange_indices([s.cpu() for s in send_segments], recv_counts, group)
        return recv.to(proto.device), offs
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


def _ctl_device(group, device: int):
    """where control-plane tensors live: on the GPU for NCCL, on the host for gloo (CPU tests, and several ranks
    sharing one GPU where NCCL refuses to run)"""
    import torch
    import torch.distributed as dist

    return torch.device(f"cuda:{device}") if dist.get_backend(group) == "nccl" else torch.device("cpu")


def _all_gather_rows(t, group):
    """all-gather of equally shaped device tensors -> list per rank (staged through the host under gloo)"""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if dist.get_backend(group) == "gloo" and t.is_cuda:
        host = t.cpu()
        outs = [torch.empty_like(host) for _ in range(world)]
        dist.all_gather(outs, host, group=group)
        return [o.to(t.device) for o in outs]
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    return outs


def _all_reduce_int(value: int, op, group, device: int) -> int:
    import torch
    import torch.distributed as dist

    t = torch.tensor([int(value)], dtype=torch.int64, device=_ctl_device(group, device))
    dist.all_reduce(t, op=op, group=group)
    return int(t.item())


class ShardedBloomFilter:
    """One logical BloomFilter(est_elements, false_positive_rate) range-sharded over the ranks of `group`.
    Every rank constructs it with the same arguments; `add_many` takes each rank's own keys (device
    resident uint8[n,16] tensors or anything pack_keys accepts).

    mode "p2p" (default): partition + exchange over NVLink peer memory + apply (csrc/pb_p2p.cu); the data path uses no
    NCCL.  mode "route": u64 bit indices binned by owner, NCCL all-to-all-v, RED apply -- exact for any key
    distribution, also the fallback for indices that overflow a p2p sublist.  mode "gather": all-gather the keys,
    every rank hashes everything and applies its own range (any key width)."""

    def __init__(self, est_elements, false_positive_rate, group=None, device=None, context=None, chunk_keys: int = 1 << 27,
                 mode: str = "p2p", window_log2: int = 27):
        import torch
        import torch.distributed as dist

        self._dist, self._torch = dist, torch
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._fpr, self._k, self._m = optimized_params(est_elements, false_positive_rate)
        self._est = est_elements
        self.plan = ShardPlan.make(self._m, self.world, window_log2)
        self.lo, self.hi = self.plan.bounds(self.rank)
        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        # run on torch's current stream so kernels, NCCL collectives and tensor ops are ordered without
        # host synchronization (handle 0 is the legacy default stream = cudaStreamLegacy, 0x1)
        self._ctx = context if context is not None else torch_stream_context(self.device)
        if mode not in ("p2p", "route", "gather"):
            raise ValueError("mode must be 'p2p', 'route' or 'gather'")
        if mode == "p2p" and self._k > 16:
            mode = "route"  # the partition kernels are instantiated for up to 16 hashes
        self.mode = mode
        self.chunk_keys = int(chunk_keys)
        # The shard lives on its own *state stream*: everything that touches the bit array (pass 2 of the insert, clear,
        # bit tests, popcount, downloads) is ordered there, and only operations that exchange tensors with the caller's
        # stream join the two (_state_after_main / _main_after_state).  add_many therefore returns once its pass 1 is
        # done: the exchange and pass 2 of its last chunks overlap with pass 1 of the caller's NEXT add_many instead of
        # draining the pipeline at every call (r2: 26 ms of a 109 ms step at N = 8).
        self._s_state = torch.cuda.Stream(device=self.device)
        self._ctx_state = _native.Context(self.device, stream=self._s_state.cuda_stream)
        h = C.c_void_p()
        if self.hi > self.lo:
            _native.call("pb_bloom_create_shard", self._ctx_state.handle, self._m, self._k, self.lo, self.hi, C.byref(h))
        self._h = h if self.hi > self.lo else None
        self._els_added = 0
        self._send = None
        self._counts = None
        self._s_part = self._ctx_part = None
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
            if not sys.is_finalizing():  # at interpreter exit the CUDA context may already be gone
                self.close()
        except Exception:
            pass

    def _state_after_main(self) -> None:
        """the state stream waits for what the caller's stream has enqueued (tensors it is about to consume)"""
        self._s_state.wait_stream(self._torch.cuda.current_stream(self.device))

    def _main_after_state(self) -> None:
        """the caller's stream waits for the state stream (tensors it produced; the bit array is up to date)"""
        self._torch.cuda.current_stream(self.device).wait_stream(self._s_state)

    def join(self) -> None:
        """order the caller's current stream after every insert issued so far (exchange and pass 2 included)"""
        main = self._torch.cuda.current_stream(self.device)
        if self._s_part is not None:
            main.wait_stream(self._s_part)
        main.wait_stream(self._s_state)

    def clear(self) -> None:
        self._els_added = 0
        if self._h is not None:
            _native.call("pb_bloom_clear", self._h)  # on the state stream: after every pass 2 issued so far

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

    def test_bit_indices(self, idx):
        """for global bit indices (int64 CUDA tensor) inside this shard's range: uint8 tensor of the bits"""
        torch = self._torch
        out = torch.zeros(idx.numel(), dtype=torch.uint8, device=idx.device)
        if idx.numel() and self._h is not None:
            self._state_after_main()
            _native.call("pb_bloom_test_bit_indices", self._h, C.c_void_p(idx.data_ptr()), idx.numel(), C.c_void_p(out.data_ptr()))
            self._main_after_state()
        return out

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

    def _n_chunks(self, n: int, chunk: int) -> int:
        """all ranks walk the same number of chunks"""
        return _all_reduce_int(-(-n // chunk), self._dist.ReduceOp.MAX, self.group, self.device)

    def add_many(self, keys) -> None:
        """BloomFilter.add (bloom.py:234-250) for this rank's keys; collective: every rank must call it
        (an empty batch is fine)"""
        t = self._device_keys(keys)
        n = int(t.shape[0])
        if self.mode == "gather":
            self._add_gather(t)
        elif t.shape[1] != 16:
            raise TypeError("p2p/route modes take 16-byte keys; use mode='gather' for other widths")
        elif self.mode == "p2p":
            self._add_p2p(t)
        else:
            self._add_route(t, min(self.chunk_keys, 1 << 26))
        self._els_added += n

    def _add_route(self, t, chunk_keys: int) -> None:
        """u64 global indices binned by owner + all-to-all-v with counts + RED apply.  Slots are sized for a uniform
        spread; a chunk whose keys are too skewed for them is redone with worst-case slots -- the decision is
        collective (all-reduce of the overflow flag), so no rank is ever left alone inside a collective."""
        torch, dist = self._torch, self._dist
        n = int(t.shape[0])
        for ci in range(self._n_chunks(n, chunk_keys)):
            lo = min(ci * chunk_keys, n)
            hi = min(lo + chunk_keys, n)
            for worst_case in (False, True):
                self._ensure_buffers(max(hi - lo, 1), worst_case)
                self._counts.zero_()
                if hi > lo:
                    kb = pack_keys(t[lo:hi], sync=False)
                    _native.call("pb_bloom_route_keys", self._ctx.handle, kb.ref(), self._m, self._k, self.plan.shard_bits,
                                 self.world, C.c_void_p(self._send.data_ptr()), self._slot, C.c_void_p(self._counts.data_ptr()))
                counts = self._counts[: self.world].clone()
                over = _all_reduce_int(int(counts.max().item()) > self._slot, dist.ReduceOp.MAX, self.group, self.device)
                if not over:
                    break
            recv_counts = exchange_counts(counts.to(_ctl_device(self.group, self.device)), self.group)
            c_host = counts.tolist()
            segs = [self._send[d * self._slot : d * self._slot + c_host[d]] for d in range(self.world)]
            recv, _ = exchange_indices(segs, recv_counts.tolist(), self.group)
            if recv.numel() and self._h is not None:
                self._state_after_main()
                _native.call("pb_bloom_add_bit_indices", self._h, C.c_void_p(recv.data_ptr()), recv.numel())  # synchronous

    # -- partition + exchange over NVLink peer memory: see pb_p2p_* in include/pb200.h
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
            self._ctx_part = _native.Context(self.device, stream=self._s_part.cuda_stream)
        n_sub, sub_cap = C.c_uint32(), C.c_uint32()
        _native.call("pb_bloom_partition_layout", self._ctx_part.handle, chunk, self._k, self._m, plan.window_log2, W,
                     C.byref(n_sub), C.byref(sub_cap))
        h = C.c_void_p()
        _native.call("pb_p2p_create", self._ctx_part.handle, self.world, self.rank, plan.windows_per_rank, n_sub.value, sub_cap.value,
                     C.byref(h))
        dev = f"cuda:{self.device}"
        mine = np.zeros(64, dtype=np.uint8)
        _native.call("pb_p2p_export", h, C.c_void_p(mine.ctypes.data))
        ctl = _ctl_device(self.group, self.device)
        allh = [torch.empty(64, dtype=torch.uint8, device=ctl) for _ in range(self.world)]
        dist.all_gather(allh, torch.from_numpy(mine).to(ctl), group=self.group)
        handles = np.ascontiguousarray(torch.stack(allh).cpu().numpy())
        _native.call("pb_p2p_connect", h, C.c_void_p(handles.ctypes.data))
        dist.barrier(group=self.group)  # every mailbox is mapped everywhere before anyone writes
        self._p2p = {"chunk": chunk, "h": h, "n_sub": n_sub.value, "sub_cap": sub_cap.value,
                     "ovf": torch.empty(1 << 22, dtype=torch.int64, device=dev), "ovf_n": torch.zeros(1, dtype=torch.int64, device=dev)}
        return self._p2p

    def _add_p2p(self, t) -> None:
        """pass 1 + copy-engine pushes on stream s_part; pass 2 on the shard's state stream.  Returns when pass 1 of
        every chunk is done (the keys may be reused): the tail of the exchange and of pass 2 keeps running."""
        torch, dist = self._torch, self._dist
        plan = self.plan
        n = int(t.shape[0])
        n_chunks = self._n_chunks(n, self.chunk_keys)
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
        main.wait_stream(self._s_part)  # pass 1 has read the keys and counted its overflows
        # one control-plane exchange per batch: did anybody's flag wait time out, did any sublist overflow?
        mine_ovf = int(b["ovf_n"].item())  # synchronizes `main`, i.e. pass 1 -- not the exchange or pass 2
        aborted = C.c_int(0)
        _native.call("pb_p2p_check", b["h"], C.byref(aborted))
        if _all_reduce_int(aborted.value, dist.ReduceOp.MAX, self.group, self.device):
            raise RuntimeError("multi-GPU insert aborted: a rank stopped answering (flag wait timed out, p2p_timeout_ms)")
        worst = _all_reduce_int(mine_ovf, dist.ReduceOp.MAX, self.group, self.device)
        if worst > b["ovf"].numel():
            # more strays than the list holds (e.g. one key repeated millions of times): OR is idempotent, so
            # simply run the whole batch again through the exact u64 route
            self._add_route(t, min(self.chunk_keys, 1 << 20))
        elif worst > 0:
            self._route_indices(b["ovf"][:mine_ovf])

    def _route_indices(self, idx) -> None:
        """send global bit indices (int64 tensor) to their owners and OR them in (collective)"""
        recv, _, _ = self._exchange_by_owner(idx)
        if recv.numel() and self._h is not None:
            self._state_after_main()
            _native.call("pb_bloom_add_bit_indices", self._h, C.c_void_p(recv.data_ptr()), recv.numel())  # synchronous

    def _exchange_by_owner(self, idx):
        """all-to-all-v of global bit indices to the ranks that own them.  Returns (what I received, the order my
        indices were sent in, how many went to each rank) -- the last two let answers travel back."""
        torch = self._torch
        owner = torch.div(idx, self.plan.shard_bits, rounding_mode="floor")
        order = torch.argsort(owner, stable=True)
        sent = idx[order].contiguous()
        counts = torch.bincount(owner, minlength=self.world)[: self.world]
        recv_counts = exchange_counts(counts.to(_ctl_device(self.group, self.device)), self.group)
        offs = [0] + torch.cumsum(counts, 0).tolist()
        segs = [sent[offs[d] : offs[d + 1]] for d in range(self.world)]
        recv, _ = exchange_indices(segs, recv_counts.tolist(), self.group)
        return recv, order, (counts.tolist(), recv_counts.tolist())

    def _add_gather(self, t) -> None:
        torch, dist = self._torch, self._dist
        sizes = self._all_sizes(int(t.shape[0]))
        mx = max(sizes)
        if mx == 0:
            return
        pad = torch.zeros((mx, t.shape[1]), dtype=torch.uint8, device=t.device)
        pad[: t.shape[0]] = t
        allk = _all_gather_rows(pad, self.group)
        if self._h is None:
            return
        self._state_after_main()
        for r in range(self.world):
            if sizes[r]:
                _native.call("pb_bloom_add_keys", self._h, pack_keys(allk[r][: sizes[r]], sync=False).ref())
        self._ctx_state.synchronize()

    def _all_sizes(self, n: int) -> list[int]:
        torch, dist = self._torch, self._dist
        mine = torch.tensor([n], dtype=torch.int64, device=_ctl_device(self.group, self.device))
        sizes = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(sizes, mine, group=self.group)
        return [int(s.item()) for s in sizes]

    def check_many(self, keys, chunk_keys: int = 1 << 24):
        """BloomFilter.check (bloom.py:252-272) for this rank's keys -> bool tensor; collective.
        The k bit indices of every key travel to the ranks that own the bits (all-to-all-v, 8 B per index), the
        owners answer with one byte per index, and the source ANDs the k answers of each key: the traffic per key does
        not grow with the number of ranks and every key is hashed once, on its own rank.  mode 'gather' keeps the
        simple scheme (all-gather the keys, AND of partial answers by all-reduce) for keys that are not 16 bytes."""
        torch, dist = self._torch, self._dist
        t = self._device_keys(keys)
        if self.mode == "gather" or t.shape[1] != 16:
            return self._check_gather(t)
        n = int(t.shape[0])
        out = torch.empty(n, dtype=torch.uint8, device=t.device)
        for ci in range(self._n_chunks(n, chunk_keys)):
            lo = min(ci * chunk_keys, n)
            hi = min(lo + chunk_keys, n)
            m = hi - lo
            idx = torch.empty(m * self._k, dtype=torch.int64, device=t.device)
            if m:
                _native.call("pb_bloom_index_keys", self._ctx.handle, pack_keys(t[lo:hi], sync=False).ref(), self._m, self._k,
                             C.c_void_p(idx.data_ptr()))
            recv, order, (sent_counts, recv_counts) = self._exchange_by_owner(idx)
            answers = self.test_bit_indices(recv)
            a_off = [0] + np.cumsum(recv_counts).tolist()
            back, _ = exchange_indices([answers[a_off[r] : a_off[r + 1]] for r in range(self.world)], sent_counts, self.group)
            if m:
                bits = torch.empty(m * self._k, dtype=torch.uint8, device=t.device)
                bits[order] = back  # answers arrive in the order the indices were sent
                _native.call("pb_bloom_and_rows", self._ctx.handle, C.c_void_p(bits.data_ptr()), m, self._k,
                             C.c_void_p(out[lo:hi].data_ptr()))
        self._ctx.synchronize()
        return out.bool()

    def _check_gather(self, t):
        torch, dist = self._torch, self._dist
        sizes = self._all_sizes(int(t.shape[0]))
        mx = max(sizes)
        if mx == 0:
            return torch.zeros(0, dtype=torch.bool, device=t.device)
        pad = torch.zeros((mx, t.shape[1]), dtype=torch.uint8, device=t.device)
        pad[: t.shape[0]] = t
        allk = _all_gather_rows(pad, self.group)
        partial = torch.ones((self.world, mx), dtype=torch.uint8, device=t.device)
        if self._h is not None:
            self._state_after_main()
            for r in range(self.world):
                if sizes[r]:
                    _native.call("pb_bloom_check_keys", self._h, pack_keys(allk[r][: sizes[r]], sync=False).ref(),
                                 C.c_void_p(partial[r].data_ptr()), 1)
            self._ctx_state.synchronize()
        if dist.get_backend(self.group) == "gloo":
            host = partial.cpu()
            dist.all_reduce(host, op=dist.ReduceOp.MIN, group=self.group)
            partial = host.to(t.device)
        else:
            dist.all_reduce(partial, op=dist.ReduceOp.MIN, group=self.group)
        return partial[self.rank, : t.shape[0]].bool()


class ShardedCountMinSketch:
    """data-parallel Count-Min: a private table per rank for its share of the stream (`local`).  merge() builds the
    sketch of the whole stream in a SECOND table (`merged`) that every rank holds: one all-reduce (sum) of the
    tables widened to int64, narrowed back with the reference's saturation -- for non-negative counts exactly what
    joining the ranks' sketches one after the other gives (CountMinSketch.join, countminsketch.py:356-399).
    The private tables are never overwritten, so merge() can be called again after more add_many calls (periodic
    merging) without counting anything twice."""

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
        self.merged = CountMinSketch(width=width, depth=depth, device=self.device, context=ctx)

    def add_many(self, keys, num_els=1) -> None:
        self.local.add_many(keys, num_els)

    def merge(self) -> None:
        torch, dist = self._torch, self._dist
        c = self.local
        n = c.width * c.depth
        sums = torch.empty(n, dtype=torch.int64, device=f"cuda:{self.device}")
        _native.call("pb_cms_widen", c._h, C.c_void_p(sums.data_ptr()), n)
        c._ctx.synchronize()
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        else:
            host = sums.cpu()
            dist.all_reduce(host, op=dist.ReduceOp.SUM, group=self.group)
            sums.copy_(host)
        _native.call("pb_cms_load_sums", self.merged._h, C.c_void_p(sums.data_ptr()), n)
        c._ctx.synchronize()
        total = _all_reduce_int(c.elements_added, dist.ReduceOp.SUM, self.group, self.device)
        self.merged._elements_added = min(total, (1 << 63) - 1)  # countminsketch.py:285-287

    @property
    def elements_added(self) -> int:
        """of the merged sketch (the whole stream as of the last merge())"""
        return self.merged.elements_added

    def check_many(self, keys):
        """estimates from the merged sketch (call merge() first)"""
        return self.merged.check_many(keys)
