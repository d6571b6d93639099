"""CuckooFilter with a device-resident bucket table: the reference's class surface
(probables/cuckoo/cuckoo.py:21-524) plus `add_many` / `check_many`.

Parity contract (see csrc/pb_cuckoo.cu): the *set of stored fingerprints*, `elements_added` and every
`check` result equal the reference's; slot placement inside the table is free, as it is in the reference
itself (its eviction walk draws from Python's unseeded global RNG, cuckoo.py:373/:377).

A custom `hash_function` decides BOTH the fingerprint and idx_2 = hash_function(str(fingerprint)) % capacity
(cuckoo.py:489, :499): such filters run *pre-indexed* -- the host evaluates the user's function (that is the plugin),
ships (fingerprint, idx_2) pairs, and the device keeps the idx_2 of every stored fingerprint next to it so that the
eviction walk can move victims without re-hashing them.  A filter exported by the reference with any hash function
therefore loads and answers here exactly as it does there.
"""

from __future__ import annotations

import ctypes as C
import sys
import math
import mmap as _mmap
import struct
from io import BytesIO, IOBase
from numbers import Number
from pathlib import Path

import numpy as np

from . import _native
from .exceptions import CuckooFilterFullError, InitializationError, NotSupportedError
from .hashes import fnv_1a, is_default_hash
from .keys import pack_keys

_FOOTER = struct.Struct("II")  # bucket_size, max_swaps (cuckoo.py:404)
_FAILED_CAP = 1 << 16


class CuckooFilter:
    """Args: capacity, bucket_size, max_swaps, expansion_rate, auto_expand, finger_size, filepath,
    hash_function -- cuckoo.py:51-61.  Extra keywords: device, context, rng_seed (eviction choices)."""

    def __init__(
        self,
        capacity: int = 10000,
        bucket_size: int = 4,
        max_swaps: int = 500,
        expansion_rate: int = 2,
        auto_expand: bool = True,
        finger_size: int = 4,
        filepath=None,
        hash_function=None,
        *,
        device: int = 0,
        context=None,
        rng_seed: int = 0,
    ):
        ok = (
            isinstance(capacity, Number)
            and capacity >= 1
            and isinstance(bucket_size, Number)
            and bucket_size >= 1
            and isinstance(max_swaps, Number)
            and max_swaps >= 1
        )
        if not ok:
            raise InitializationError("CuckooFilter: capacity, bucket_size, and max_swaps must be an integer greater than 0")
        self._ctx_arg = (context, device)  # resolved in _create(): argument errors surface before any device work
        self._ctx = context
        self._h = None
        self._rng_seed = int(rng_seed)
        self._bucket_size = int(bucket_size)
        self._capacity = int(capacity)
        self._max_swaps = int(max_swaps)
        self.expansion_rate = expansion_rate
        self.auto_expand = auto_expand
        self._fingerprint_bits = 32
        self.fingerprint_size = finger_size
        self._hash_func = hash_function if hash_function is not None else fnv_1a
        self._fused = is_default_hash(hash_function)
        self._inserted = 0
        if filepath is None:
            self._create()
        elif Path(filepath).expanduser().is_file():
            self._load(Path(filepath).expanduser().read_bytes())
        else:
            raise InitializationError("CuckooFilter: failed to load provided file")
        self._error_rate = float(self._calc_error_rate())

    # ------------------------------------------------------------------ alternative constructors
    @classmethod
    def init_error_rate(cls, error_rate, capacity=10000, bucket_size=4, max_swaps=500, expansion_rate=2, auto_expand=True,
                        hash_function=None, **kw):
        """cuckoo.py:102-135"""
        cku = cls(capacity=capacity, bucket_size=bucket_size, auto_expand=auto_expand, max_swaps=max_swaps,
                  expansion_rate=expansion_rate, hash_function=hash_function, **kw)
        cku._set_error_rate(error_rate)
        return cku

    @classmethod
    def load_error_rate(cls, error_rate, filepath, hash_function=None, **kw):
        cku = cls(filepath=filepath, hash_function=hash_function, **kw)
        cku._set_error_rate(error_rate)
        return cku

    @classmethod
    def frombytes(cls, b, error_rate=None, hash_function=None, **kw):
        cku = cls(hash_function=hash_function, **kw)
        cku._load(bytes(b))
        cku._set_error_rate(error_rate)
        return cku

    # ------------------------------------------------------------------ device state
    def _create(self) -> None:
        if self._h is not None:
            _native.lib().pb_cuckoo_destroy(self._h)
            self._h = None
        if self._fingerprint_bits > 32:
            raise NotSupportedError("fingerprints wider than 32 bits do not fit the u32 slot format (cuckoo.py:402)")
        if self._ctx is None:
            self._ctx = _native.default_context(self._ctx_arg[1])
        h = C.c_void_p()
        _native.call("pb_cuckoo_create", self._ctx.handle, self._capacity, self._bucket_size, self._max_swaps,
                     self._fingerprint_bits, self._rng_seed, C.byref(h))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and _native._lib is not None:
            _native._lib.pb_cuckoo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if not sys.is_finalizing():  # at interpreter exit the CUDA context may already be gone
                self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ properties (cuckoo.py:196-289)
    @property
    def elements_added(self) -> int:
        return self._inserted

    @property
    def capacity(self) -> int:
        return self._capacity

    @property
    def max_swaps(self) -> int:
        return self._max_swaps

    @property
    def bucket_size(self) -> int:
        return self._bucket_size

    @property
    def expansion_rate(self) -> int:
        return self._expansion_rate

    @expansion_rate.setter
    def expansion_rate(self, val: int):
        self._expansion_rate = val

    @property
    def error_rate(self) -> float:
        return self._error_rate

    @property
    def auto_expand(self) -> bool:
        return self._auto_expand

    @auto_expand.setter
    def auto_expand(self, val: bool):
        self._auto_expand = bool(val)

    @property
    def fingerprint_size_bits(self) -> int:
        return self._fingerprint_bits

    @property
    def fingerprint_size(self) -> int:
        return math.ceil(self._fingerprint_bits / 8)

    @fingerprint_size.setter
    def fingerprint_size(self, val: int):
        if not 1 <= val <= 4:
            raise ValueError(f"{self.__class__.__name__}: fingerprint size must be between 1 and 4")
        self._set_fingerprint_bits(val * 8)

    def _set_fingerprint_bits(self, bits: int) -> None:
        """the device handle carries the fingerprint mask: changing it on a live filter re-creates the handle around
        the same table (the reference only changes the mask applied to new keys, cuckoo.py:276-285)"""
        old = getattr(self, "_fingerprint_bits", None)
        self._fingerprint_bits = bits
        if getattr(self, "_h", None) is None or old == bits:
            return
        if self._inserted == 0:
            self._create()
            return
        slots, z = self.slots_numpy()
        self._create()
        flat = np.ascontiguousarray(slots.reshape(-1))
        _native.call("pb_cuckoo_upload", self._h, C.c_void_p(flat.ctypes.data), flat.size, int(z))
        self._upload_alt(flat)

    def load_factor(self) -> float:
        return self.elements_added / (self.capacity * self.bucket_size)

    def slots_numpy(self):
        """(uint32[capacity, bucket_size] host copy of the table, has_zero_fingerprint)"""
        out = np.empty(self._capacity * self._bucket_size, dtype=np.uint32)
        z = C.c_int(0)
        _native.call("pb_cuckoo_download", self._h, C.c_void_p(out.ctypes.data), out.size, C.byref(z))
        return out.reshape(self._capacity, self._bucket_size), bool(z.value)

    @property
    def buckets(self) -> list:
        """the reference's list-of-lists view (host copy; the zero fingerprint, kept as a device flag,
        is listed in its idx_1 bucket = bucket 0)"""
        slots, has_zero = self.slots_numpy()
        out = [[int(x) for x in row if x] for row in slots]
        if has_zero:
            out[0].append(0)
        return out

    def fingerprints(self) -> np.ndarray:
        """sorted uint32 array of every stored fingerprint"""
        slots, has_zero = self.slots_numpy()
        fps = slots[slots != 0]
        if has_zero:
            fps = np.concatenate([fps, np.zeros(1, dtype=np.uint32)])
        return np.sort(fps)

    def __contains__(self, key) -> bool:
        return self.check(key)

    def __str__(self):
        return (
            f"{self.__class__.__name__}:\n"
            f"\tCapacity: {self.capacity}\n"
            f"\tTotal Bins: {self.capacity * self.bucket_size}\n"
            f"\tLoad Factor: {self.load_factor() * 100}%\n"
            f"\tInserted Elements: {self.elements_added}\n"
            f"\tMax Swaps: {self.max_swaps}\n"
            f"\tExpansion Rate: {self.expansion_rate}\n"
            f"\tAuto Expand: {self.auto_expand}"
        )

    # ------------------------------------------------------------------ hot path
    def _plugin_pairs(self, keys):
        """custom hash_function: fp = low fp_bits of hash_function(key) (cuckoo.py:499-500) and
        idx_2 = hash_function(str(fp)) % capacity (:489), on the host -- the user's function IS the plugin"""
        mask = (1 << self._fingerprint_bits) - 1
        fps = [self._hash_func(k) & mask for k in keys]
        return np.asarray(fps, dtype=np.uint32), self._plugin_alt(fps)

    def _plugin_alt(self, fps) -> np.ndarray:
        cap = self._capacity
        return np.asarray([self._hash_func(str(int(fp))) % cap for fp in fps], dtype=np.uint64)

    @staticmethod
    def _as_list(keys):
        return [keys] if isinstance(keys, (str, bytes, bytearray, memoryview)) else list(keys)

    def _handle_failures(self, status_full: bool, n_failed: int, failed: np.ndarray) -> None:
        """_deal_with_insertion (cuckoo.py:508-516) for a batch"""
        if not status_full:
            return
        if not self._auto_expand:
            raise CuckooFilterFullError(f"The {self.__class__.__name__} is currently full")
        if n_failed > failed.size:
            raise CuckooFilterFullError(
                f"The {self.__class__.__name__} failed to expand: {n_failed} homeless fingerprints exceed the report buffer"
            )
        pending = failed[:n_failed].copy()
        while pending.size:
            self._expand_device()
            pending = self._add_fps_raw(pending)

    def _add_fps_raw(self, fps: np.ndarray) -> np.ndarray:
        """insert fingerprints; returns the homeless ones (empty when all went in)"""
        fps = np.ascontiguousarray(fps, dtype=np.uint32)
        if not self._fused:
            return self._add_pairs(fps, self._plugin_alt(fps))
        n_added, n_failed = C.c_uint64(0), C.c_uint64(0)
        failed = np.empty(max(fps.size, 1), dtype=np.uint32)
        st = _native.lib().pb_cuckoo_add_fingerprints(
            self._h, C.c_void_p(fps.ctypes.data), fps.size, 0, C.byref(n_added), C.byref(n_failed),
            C.c_void_p(failed.ctypes.data), failed.size,
        )
        if st not in (_native.PB_OK, _native.PB_ERR_CUCKOO_FULL):
            _native.check(st)
        self._inserted += n_added.value
        if n_failed.value > failed.size:
            raise CuckooFilterFullError(f"The {self.__class__.__name__} failed to expand")
        return failed[: n_failed.value].copy()

    def _add_pairs(self, fps: np.ndarray, alt: np.ndarray) -> np.ndarray:
        """pre-indexed insert (key order); returns the homeless fingerprints"""
        n_added, n_failed = C.c_uint64(0), C.c_uint64(0)
        failed = np.empty(max(fps.size, 1), dtype=np.uint32)
        failed_alt = np.empty(max(fps.size, 1), dtype=np.uint64)
        st = _native.lib().pb_cuckoo_add_indexed(
            self._h, C.c_void_p(fps.ctypes.data), C.c_void_p(alt.ctypes.data), fps.size, C.byref(n_added), C.byref(n_failed),
            C.c_void_p(failed.ctypes.data), C.c_void_p(failed_alt.ctypes.data), failed.size,
        )
        if st not in (_native.PB_OK, _native.PB_ERR_CUCKOO_FULL):
            _native.check(st)
        self._inserted += n_added.value
        return failed[: n_failed.value].copy()

    def add_many(self, keys) -> None:
        """CuckooFilter.add (cuckoo.py:291-304) for every key.  With auto_expand the table grows by
        expansion_rate until every fingerprint is stored; without it CuckooFilterFullError is raised after
        the batch (the fingerprints that did fit stay stored, as they would in the reference up to the
        failing key)."""
        if self._fused:
            n_added, n_failed = C.c_uint64(0), C.c_uint64(0)
            kb = pack_keys(keys)
            if kb.n == 0:
                return
            # room for every homeless fingerprint the batch could produce (untouched pages cost nothing)
            failed = np.empty(kb.n, dtype=np.uint32)
            st = _native.lib().pb_cuckoo_add_keys(
                self._h, kb.ref(), C.byref(n_added), C.byref(n_failed), C.c_void_p(failed.ctypes.data), failed.size
            )
            if st not in (_native.PB_OK, _native.PB_ERR_CUCKOO_FULL):
                _native.check(st)
            self._inserted += n_added.value
            self._handle_failures(st == _native.PB_ERR_CUCKOO_FULL, n_failed.value, failed)
            return
        fps, alt = self._plugin_pairs(self._as_list(keys))
        if fps.size == 0:
            return
        failed = self._add_pairs(fps, alt)
        self._handle_failures(failed.size > 0, failed.size, failed)

    def check_many(self, keys) -> np.ndarray:
        """CuckooFilter.check (cuckoo.py:306-315) for every key -> bool[n]"""
        if self._fused:
            kb = pack_keys(keys)
            out = np.empty(kb.n, dtype=np.uint8)
            if kb.n:
                _native.call("pb_cuckoo_check_keys", self._h, kb.ref(), C.c_void_p(out.ctypes.data), 0)
            return out.astype(bool)
        fps, alt = self._plugin_pairs(self._as_list(keys))
        out = np.empty(fps.size, dtype=np.uint8)
        if fps.size:
            _native.call("pb_cuckoo_check_indexed", self._h, C.c_void_p(fps.ctypes.data), C.c_void_p(alt.ctypes.data), fps.size,
                         C.c_void_p(out.ctypes.data))
        return out.astype(bool)

    def remove_many(self, keys) -> np.ndarray:
        """CuckooFilter.remove (cuckoo.py:317-330) for every key -> bool[n] (True: a stored fingerprint was removed).
        The table and elements_added end up as after the sequential loop.  When several keys of one batch share a
        fingerprint exactly one of them reports True -- the first one, as in the reference, for batches of up to 2^20
        keys (the flags are put in order on the host); beyond that it is whichever thread cleared the slot."""
        removed = C.c_uint64(0)
        if self._fused:
            kb = pack_keys(keys)
            out = np.zeros(kb.n, dtype=np.uint8)
            if kb.n:
                _native.call("pb_cuckoo_remove_keys", self._h, kb.ref(), C.c_void_p(out.ctypes.data), 0, C.byref(removed))
            fps = None
        else:
            fps, alt = self._plugin_pairs(self._as_list(keys))
            out = np.zeros(fps.size, dtype=np.uint8)
            if fps.size:
                _native.call("pb_cuckoo_remove_indexed", self._h, C.c_void_p(fps.ctypes.data), C.c_void_p(alt.ctypes.data), fps.size,
                             C.c_void_p(out.ctypes.data), C.byref(removed))
        self._inserted -= removed.value
        res = out.astype(bool)
        if 1 < res.size <= (1 << 20) and res.any():
            if fps is None:
                fps = self.fingerprint_info_many(keys)[2]
            _, first = np.unique(fps, return_index=True)
            winners = np.unique(fps[res])
            ordered = np.zeros_like(res)
            uniq = fps[first]
            ordered[first[np.isin(uniq, winners)]] = True
            res = ordered
        return res

    def add(self, key) -> None:
        self.add_many([key])

    def check(self, key) -> bool:
        return bool(self.check_many([key])[0])

    def remove(self, key) -> bool:
        """cuckoo.py:317-330"""
        return bool(self.remove_many([key])[0])

    def fingerprint_info_many(self, keys):
        """_generate_fingerprint_info (cuckoo.py:492-506) for a batch -> (idx_1, idx_2, fingerprint) arrays"""
        if not self._fused:
            fps, alt = self._plugin_pairs(self._as_list(keys))
            return (fps.astype(np.uint64) % np.uint64(self._capacity)), alt, fps
        kb = pack_keys(keys)
        fp = np.empty(kb.n, dtype=np.uint32)
        i1 = np.empty(kb.n, dtype=np.uint64)
        i2 = np.empty(kb.n, dtype=np.uint64)
        if kb.n:
            _native.call("pb_cuckoo_fingerprint_info", self._h, kb.ref(), C.c_void_p(fp.ctypes.data),
                         C.c_void_p(i1.ctypes.data), C.c_void_p(i2.ctypes.data), 0)
        return i1, i2, fp

    def _generate_fingerprint_info(self, key):
        i1, i2, fp = self.fingerprint_info_many([key])
        return int(i1[0]), int(i2[0]), int(fp[0])

    # ------------------------------------------------------------------ expansion (cuckoo.py:351-353, :455-481)
    def _expand_device(self) -> None:
        new_cap = self._capacity * self._expansion_rate
        if not self._fused:
            # cuckoo.py:455-481 with the user's hash: collect, re-create at the new capacity, re-insert in bucket order
            slots, has_zero = self.slots_numpy()
            fps = slots.reshape(-1)
            fps = fps[fps != 0]
            if has_zero:
                fps = np.concatenate([np.zeros(1, dtype=np.uint32), fps])
            _native.call("pb_cuckoo_resize", self._h, new_cap)
            self._capacity = new_cap
            self._inserted = 0
            if fps.size and self._add_pairs(np.ascontiguousarray(fps), self._plugin_alt(fps)).size:
                raise CuckooFilterFullError("The CuckooFilter failed to expand")
            return
        n_failed = C.c_uint64(0)
        failed = np.zeros(_FAILED_CAP, dtype=np.uint32)
        st = _native.lib().pb_cuckoo_expand(self._h, new_cap, C.byref(n_failed), C.c_void_p(failed.ctypes.data), failed.size)
        self._capacity = new_cap
        if st == _native.PB_ERR_CUCKOO_FULL:
            raise CuckooFilterFullError("The CuckooFilter failed to expand")
        _native.check(st)

    def expand(self) -> None:
        self._expand_device()

    # ------------------------------------------------------------------ wire format (cuckoo.py:332-349, :394-431)
    def export(self, file) -> None:
        if not isinstance(file, (IOBase, _mmap.mmap)):
            with open(Path(file).expanduser(), "wb") as fp:
                self.export(fp)
            return
        slots, _ = self.slots_numpy()
        file.write(slots.tobytes())
        file.write(_FOOTER.pack(self._bucket_size, self._max_swaps))

    def __bytes__(self) -> bytes:
        with BytesIO() as f:
            self.export(f)
            return f.getvalue()

    def _load(self, data: bytes) -> None:
        body = len(data) - _FOOTER.size
        self._bucket_size, self._max_swaps = _FOOTER.unpack(data[body:])
        self._capacity = body // 4 // self._bucket_size
        self._create()
        slots = np.frombuffer(data[: self._capacity * self._bucket_size * 4], dtype=np.uint32)
        _native.call("pb_cuckoo_upload", self._h, C.c_void_p(slots.ctypes.data), slots.size, 0)
        self._inserted = int(np.count_nonzero(slots))  # zero entries are dropped on load (cuckoo.py:429)
        self._upload_alt(slots)

    def _upload_alt(self, flat_slots: np.ndarray) -> None:
        """pre-indexed filters: the idx_2 of every stored fingerprint, from the user's hash function"""
        if self._fused:
            return
        alt = np.zeros(flat_slots.size, dtype=np.uint64)
        nz = np.nonzero(flat_slots)[0]
        if nz.size:
            alt[nz] = self._plugin_alt(flat_slots[nz])
        _native.call("pb_cuckoo_set_alt", self._h, C.c_void_p(alt.ctypes.data), alt.size)

    # ------------------------------------------------------------------ sizing helpers (cuckoo.py:433-438, :518-524)
    def _set_error_rate(self, error_rate) -> None:
        if error_rate is not None:
            self._error_rate = error_rate
            bits = int(math.ceil(math.log2(1.0 / self._error_rate) + math.log2(self._bucket_size) + 1))
            self._set_fingerprint_bits(bits)

    def _calc_error_rate(self) -> float:
        return float(1 / (2 ** (self._fingerprint_bits - (math.log2(self._bucket_size) + 1))))
