"""CountingCuckooFilter on the device (probables/cuckoo/countingcuckoo.py).

The reference stores (fingerprint, count) bins: adding a key whose fingerprint is already stored bumps the count
(:156-171), a new fingerprint goes in with count 1 and keeps its count through evictions (:230-265).  A fingerprint is
stored at most once, so the counts are a map fingerprint -> count that placement never touches.  Here the fingerprints
live in the CuckooFilter table (same kernels, same eviction walk) and the counts in an open-addressing map on the same
handle (pb_cuckoo_counts_*, include/pb200.h): add = set insert + one atomicAdd per key, check = one lookup,
remove = decrement and, at 0, take the fingerprint out of the table.  The parity contract is the reference's: the
multiset of stored (fingerprint, count) bins, elements_added, unique_elements and every check / remove result.
"""

from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from . import _native
from .cuckoo import CuckooFilter
from .exceptions import CuckooFilterFullError
from .keys import pack_keys

_BIN = struct.Struct("II")  # countingcuckoo.py:64


def _first_occurrences_win(flags: np.ndarray, fps: np.ndarray) -> np.ndarray:
    """The device hands out as many successful removals per fingerprint as its count allows, to whichever occurrences drew
    the low tickets; one key at a time it is the FIRST occurrences that succeed (countingcuckoo.py:199-208).  Same number
    of successes per fingerprint, moved to its earliest occurrences."""
    order = np.argsort(fps, kind="stable")
    sorted_fps = fps[order]
    new_group = np.r_[True, sorted_fps[1:] != sorted_fps[:-1]]
    starts = np.flatnonzero(new_group)
    group = np.cumsum(new_group) - 1
    rank = np.arange(flags.size) - starts[group]  # occurrence number of each key within its fingerprint
    wins = np.add.reduceat(flags[order].astype(np.int64), starts)  # removals that succeeded per fingerprint
    ordered = np.zeros_like(flags)
    ordered[order] = rank < wins[group]
    return ordered


class CountingCuckooBin:
    """countingcuckoo.py:337-381 (a read-only view: the state lives on the device)"""

    __slots__ = ("finger", "count")

    def __init__(self, fingerprint: int, count: int) -> None:
        self.finger = int(fingerprint)
        self.count = int(count)

    def __contains__(self, val: int) -> bool:
        return self.finger == val

    def __repr__(self) -> str:
        return f"(fingerprint:{self.finger} count:{self.count})"

    __str__ = __repr__


class CountingCuckooFilter(CuckooFilter):
    """Args as CuckooFilter (countingcuckoo.py:39-61)."""

    def __init__(self, *args, **kw):
        self._total = 0  # the reference's _inserted_elements: adds minus removes; self._inserted counts distinct fingerprints
        super().__init__(*args, **kw)

    # ------------------------------------------------------------------ device state
    def _create(self) -> None:
        super()._create()
        _native.call("pb_cuckoo_counts_enable", self._h)

    def _set_fingerprint_bits(self, bits: int) -> None:
        """as CuckooFilter; the counts of a live filter move over to the re-created handle"""
        live = (getattr(self, "_h", None) is not None and getattr(self, "_inserted", 0) > 0
                and getattr(self, "_fingerprint_bits", None) != bits)
        if live:
            slots, has_zero = self.slots_numpy()
            stored = slots[slots != 0]
            if has_zero:
                stored = np.concatenate([stored, np.zeros(1, np.uint32)])
            stored = np.ascontiguousarray(stored, dtype=np.uint32)
            counts = self._counts_of(stored)
        super()._set_fingerprint_bits(bits)
        if live and stored.size:
            _native.call("pb_cuckoo_counts_set", self._h, C.c_void_p(stored.ctypes.data), C.c_void_p(counts.ctypes.data), stored.size)

    @property
    def elements_added(self) -> int:
        return self._total

    @property
    def unique_elements(self) -> int:
        """:139-142"""
        return self._inserted

    def load_factor(self) -> float:
        """:152-154"""
        return self.unique_elements / (self.capacity * self.bucket_size)

    def __contains__(self, key) -> bool:
        return self.check(key) > 0

    def _counts_of(self, fps: np.ndarray) -> np.ndarray:
        fps = np.ascontiguousarray(fps, dtype=np.uint32)
        out = np.zeros(fps.size, dtype=np.uint32)
        if fps.size:
            _native.call("pb_cuckoo_counts_get_fingerprints", self._h, C.c_void_p(fps.ctypes.data), fps.size, C.c_void_p(out.ctypes.data))
        return out

    @property
    def buckets(self) -> list:
        """the reference's list-of-lists of bins (:144-150), a host copy"""
        slots, has_zero = self.slots_numpy()
        counts = self._counts_of(slots.reshape(-1)).reshape(slots.shape)
        out = [[CountingCuckooBin(f, c) for f, c in zip(row, crow) if f] for row, crow in zip(slots, counts)]
        if has_zero:
            out[0].append(CountingCuckooBin(0, int(self._counts_of(np.zeros(1, np.uint32))[0])))
        return out

    def bins(self) -> list:
        """sorted (fingerprint, count) pairs of everything stored"""
        return sorted((b.finger, b.count) for bucket in self.buckets for b in bucket)

    # ------------------------------------------------------------------ add / check / remove
    def add_many(self, keys) -> None:
        """CountingCuckooFilter.add (:156-173) for every key.  Failure handling as CuckooFilter.add_many; a fingerprint
        that is given up as homeless (no auto_expand) loses its count with it, as the bin the reference hands back does."""
        if self._fused:
            kb = pack_keys(keys)
            if kb.n == 0:
                return
            n = kb.n
            _native.call("pb_cuckoo_counts_add_keys", self._h, kb.ref())
            keys = kb
        else:
            fps, _ = self._plugin_pairs(self._as_list(keys))
            n = fps.size
            if n == 0:
                return
            _native.call("pb_cuckoo_counts_add_fingerprints", self._h, C.c_void_p(fps.ctypes.data), None, n, 0)
        self._total += n
        try:
            super().add_many(keys)
        except CuckooFilterFullError:
            self._resync_after_failure()
            raise

    def _resync_after_failure(self) -> None:
        """homeless fingerprints were dropped from the table: their counts go too (the reference hands the bin back to the
        caller, :264-265), and the totals follow what is stored"""
        slots, has_zero = self.slots_numpy()
        flat = np.ascontiguousarray(slots.reshape(-1))
        stored = flat[flat != 0]
        if has_zero:
            stored = np.concatenate([stored, np.zeros(1, np.uint32)])
        stored = np.ascontiguousarray(stored, dtype=np.uint32)
        counts = self._counts_of(stored)
        _native.call("pb_cuckoo_clear", self._h)  # table and map
        _native.call("pb_cuckoo_upload", self._h, C.c_void_p(flat.ctypes.data), flat.size, 1 if has_zero else 0)
        self._upload_alt(flat)
        if stored.size:
            _native.call("pb_cuckoo_counts_set", self._h, C.c_void_p(stored.ctypes.data), C.c_void_p(counts.ctypes.data), stored.size)
        self._inserted = int(stored.size)
        self._total = int(counts.astype(np.uint64).sum())

    def check_many(self, keys) -> np.ndarray:
        """CountingCuckooFilter.check (:175-191) for every key -> uint32[n] counts (0: not stored)"""
        if self._fused:
            kb = pack_keys(keys)
            out = np.zeros(kb.n, dtype=np.uint32)
            if kb.n:
                _native.call("pb_cuckoo_counts_get_keys", self._h, kb.ref(), C.c_void_p(out.ctypes.data), 0)
            return out
        fps, _ = self._plugin_pairs(self._as_list(keys))
        return self._counts_of(fps)

    def check(self, key) -> int:
        return int(self.check_many([key])[0])

    def remove_many(self, keys) -> np.ndarray:
        """CountingCuckooFilter.remove (:193-210) for every key -> bool[n].  Counts, the table, elements_added and
        unique_elements end up as after the one-at-a-time loop.  A key that occurs more often in the batch than its
        count reports True `count` times -- its first occurrences, as in the reference, for batches of up to 2^20 keys
        (the flags are put in order on the host); beyond that, whichever threads drew the low tickets."""
        removed, bins_removed = C.c_uint64(0), C.c_uint64(0)
        if self._fused:
            kb = pack_keys(keys)
            out = np.zeros(kb.n, dtype=np.uint8)
            if kb.n:
                _native.call("pb_cuckoo_counts_remove_keys", self._h, kb.ref(), C.c_void_p(out.ctypes.data), 0, C.byref(removed),
                             C.byref(bins_removed))
            fps = None
        else:
            fps, alt = self._plugin_pairs(self._as_list(keys))
            out = np.zeros(fps.size, dtype=np.uint8)
            if fps.size:
                _native.call("pb_cuckoo_counts_remove_fingerprints", self._h, C.c_void_p(fps.ctypes.data), C.c_void_p(alt.ctypes.data),
                             fps.size, C.c_void_p(out.ctypes.data), C.byref(removed), C.byref(bins_removed))
        self._total -= removed.value
        self._inserted -= bins_removed.value
        res = out.astype(bool)
        if 1 < res.size <= (1 << 20) and res.any() and not res.all():
            if fps is None:
                fps = self.fingerprint_info_many(keys)[2]
            res = _first_occurrences_win(res, fps)
        return res

    def remove(self, key) -> bool:
        return bool(self.remove_many([key])[0])

    # ------------------------------------------------------------------ expansion (:212-214, :305-316)
    def _expand_device(self) -> None:
        super()._expand_device()
        _native.call("pb_cuckoo_counts_enable", self._h)

    # ------------------------------------------------------------------ wire format (:216-228, :275-303, :325-334)
    def export(self, file) -> None:
        from io import IOBase
        from mmap import mmap
        from pathlib import Path

        if not isinstance(file, (IOBase, mmap)):
            with open(Path(file).expanduser(), "wb") as fp:
                self.export(fp)
            return
        slots, _ = self.slots_numpy()
        counts = self._counts_of(slots.reshape(-1)).reshape(slots.shape)
        # a reference bucket is a list without holes: stored bins first, zero padding after (:325-334)
        order = np.argsort(slots == 0, axis=1, kind="stable")
        f = np.take_along_axis(slots, order, axis=1)
        c = np.where(f != 0, np.take_along_axis(counts, order, axis=1), 0).astype(np.uint32)
        file.write(np.stack([f, c], axis=2).astype(np.uint32).tobytes())
        file.write(_BIN.pack(self._bucket_size, self._max_swaps))

    def _load(self, data: bytes) -> None:
        body = len(data) - _BIN.size
        self._bucket_size, self._max_swaps = _BIN.unpack(data[body:])
        self._capacity = body // _BIN.size // self._bucket_size  # :289
        self._create()
        pairs = np.frombuffer(data[: self._capacity * self._bucket_size * _BIN.size], dtype=np.uint32).reshape(-1, 2)
        fps = np.ascontiguousarray(pairs[:, 0])
        _native.call("pb_cuckoo_upload", self._h, C.c_void_p(fps.ctypes.data), fps.size, 0)
        self._upload_alt(fps)
        keep = fps != 0  # :296: a zero fingerprint is an empty slot
        kf, kc = np.ascontiguousarray(fps[keep]), np.ascontiguousarray(pairs[keep, 1])
        if kf.size:
            _native.call("pb_cuckoo_counts_set", self._h, C.c_void_p(kf.ctypes.data), C.c_void_p(kc.ctypes.data), kf.size)
        self._inserted = int(kf.size)
        self._total = int(kc.astype(np.uint64).sum())
