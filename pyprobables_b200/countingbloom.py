"""CountingBloomFilter with device-resident counters: the reference's class surface
(probables/blooms/countingbloom.py:25-330) plus the batch seam `add_many` / `check_many` / `remove_many`.

Every add / check / remove runs in the CUDA kernels of csrc/pb_cbloom.cu through the C ABI.  Batch semantics:
  * add_many   -- counter values are order-free (saturating sums of non-negative increments), so the batch equals the
                  sequential loop, including the reference's quirk that a key whose hashes collide on one counter
                  increments it once per colliding hash (countingbloom.py:141-153);
  * remove_many -- equals the sequential loop by construction: an optimistic parallel pass that is provably exact when
                  no counter runs out, otherwise an in-order replay on the device (see pb_cbloom.cu).
"""

from __future__ import annotations

import ctypes as C
import math
import struct
from binascii import hexlify, unhexlify
from io import BytesIO
from pathlib import Path

import numpy as np

from . import _native
from .bloom import _FOOTER, _FOOTER_BE, _U64_MASK, BloomFilter, _is_file, _is_hex, optimized_params
from .exceptions import InitializationError, SimilarityError
from .hashes import default_fnv_1a, is_default_hash
from .keys import pack_keys

UINT32_T_MAX = 2**32 - 1
UINT64_T_MAX = 2**64 - 1
_MISMATCH_MSG = "The parameter second must be of type CountingBloomFilter"


class CountingBloomFilter(BloomFilter):
    """Counting Bloom filter whose uint32 counters live in GPU memory.

    Args: est_elements, false_positive_rate, filepath, hex_string, hash_function -- as countingbloom.py:47-54.
    Extra keywords: device (CUDA ordinal, default 0), context (a pyprobables_b200 Context to share a stream).
    """

    def __init__(self, est_elements=None, false_positive_rate=None, filepath=None, hex_string=None, hash_function=None, *,
                 device: int = 0, context=None):
        self._ctx_arg = (context, device)
        self._ctx = context
        self._h = None
        self._on_disk = False
        self._els_added = 0
        if _is_file(filepath):
            self._load(Path(filepath).expanduser().read_bytes(), hash_function)
        elif _is_hex(hex_string):
            self._load_hex(hex_string, hash_function)
        else:
            if est_elements is None or false_positive_rate is None:
                raise InitializationError("Insufecient parameters to set up the Counting Bloom Filter")
            fpr, k, m = optimized_params(est_elements, false_positive_rate)
            self._set_values(est_elements, fpr, k, m, hash_function)

    # ------------------------------------------------------------------ setup / teardown
    def _set_values(self, est_els, fpr, n_hashes, n_bits, hash_func):
        self._est_elements = est_els
        self._fpr = fpr
        self._number_hashes = int(n_hashes)
        self._num_bits = int(n_bits)
        self._bloom_length = int(n_bits)  # one counter per "bit" (countingbloom.py:77-78)
        self._hash_func = hash_func if hash_func is not None else default_fnv_1a
        self._fused = is_default_hash(hash_func)
        self._els_added = 0
        if self._ctx is None:
            self._ctx = _native.default_context(self._ctx_arg[1])
        if self._h is not None:
            _native.lib().pb_cbloom_destroy(self._h)
        h = C.c_void_p()
        _native.call("pb_cbloom_create", self._ctx.handle, self._num_bits, self._number_hashes, C.byref(h))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and _native._lib is not None:
            _native._lib.pb_cbloom_destroy(self._h)
            self._h = None

    # ------------------------------------------------------------------ state
    def bloom_numpy(self) -> np.ndarray:
        """host copy of the counters as uint32[number_bits]"""
        out = np.empty(self._bloom_length, dtype=np.uint32)
        _native.call("pb_cbloom_download", self._h, C.c_void_p(out.ctypes.data), out.size)
        return out

    @property
    def bloom(self):
        from array import array

        return array("I", self.bloom_numpy().tobytes())

    def device_ptr(self) -> int:
        p, n = C.c_void_p(), C.c_uint64()
        _native.call("pb_cbloom_device_ptr", self._h, C.byref(p), C.byref(n))
        return p.value

    def clear(self) -> None:
        self._els_added = 0
        _native.call("pb_cbloom_clear", self._h)

    def _stats(self):
        out = (C.c_uint64 * 4)()
        _native.call("pb_cbloom_stats", self._h, out)
        return out[0], out[1], out[2], out[3]

    def _cnt_number_bits_set(self) -> int:
        """countingbloom.py:328-330: counters that are non-zero"""
        return self._stats()[0]

    # ------------------------------------------------------------------ hot path
    def _rows(self, keys):
        if isinstance(keys, (str, bytes, bytearray, memoryview)):
            keys = [keys]
        return self._plugin_hashes(keys)

    def add_many(self, keys, num_els: int = 1) -> None:
        """CountingBloomFilter.add (countingbloom.py:125-153) for every key of the batch"""
        num_els = int(num_els)
        if not 0 <= num_els <= UINT32_T_MAX:
            raise ValueError("num_els must fit an unsigned 32-bit counter")
        if self._fused:
            kb = pack_keys(keys)
            if kb.n:
                _native.call("pb_cbloom_add_keys", self._h, kb.ref(), num_els)
            n = kb.n
        else:
            h, n = self._rows(keys)
            if n:
                _native.call("pb_cbloom_add_hashes", self._h, C.c_void_p(h.ctypes.data), n, 0, num_els)
        self._els_added = min(self._els_added + n * num_els, UINT64_T_MAX)

    def check_many(self, keys) -> np.ndarray:
        """CountingBloomFilter.check (countingbloom.py:155-174) for every key -> uint32[n] (smallest counter)"""
        if self._fused:
            kb = pack_keys(keys)
            out = np.empty(kb.n, dtype=np.uint32)
            if kb.n:
                _native.call("pb_cbloom_check_keys", self._h, kb.ref(), C.c_void_p(out.ctypes.data), 0)
            return out
        h, n = self._rows(keys)
        out = np.empty(n, dtype=np.uint32)
        if n:
            _native.call("pb_cbloom_check_hashes", self._h, C.c_void_p(h.ctypes.data), n, 0, C.c_void_p(out.ctypes.data), 0)
        return out

    def remove_many(self, keys, num_els: int = 1) -> None:
        """CountingBloomFilter.remove (countingbloom.py:176-208) for every key, in order"""
        num_els = int(num_els)
        if not 0 <= num_els <= UINT32_T_MAX:
            raise ValueError("num_els must fit an unsigned 32-bit counter")
        removed = C.c_uint64(0)
        if self._fused:
            kb = pack_keys(keys)
            if kb.n:
                _native.call("pb_cbloom_remove_keys", self._h, kb.ref(), num_els, C.byref(removed))
        else:
            h, n = self._rows(keys)
            if n:
                _native.call("pb_cbloom_remove_hashes", self._h, C.c_void_p(h.ctypes.data), n, 0, num_els, C.byref(removed))
        self._els_added -= removed.value

    def add(self, key, num_els: int = 1) -> int:
        """countingbloom.py:125-133: returns the smallest of the key's counters as the reference computes it --
        from the values BEFORE the update plus num_els (:144), saturated"""
        before = int(self.check_many([key])[0])
        self.add_many([key], num_els)
        return min(before + num_els, UINT32_T_MAX)

    def add_alt(self, hashes, num_els: int = 1) -> int:
        """countingbloom.py:135-153"""
        row = self._alt_row(hashes)
        before = self._check_row(row)
        _native.call("pb_cbloom_add_hashes", self._h, C.c_void_p(row.ctypes.data), 1, 0, int(num_els))
        self._els_added = min(self._els_added + num_els, UINT64_T_MAX)
        return min(before + num_els, UINT32_T_MAX)

    def _check_row(self, row: np.ndarray) -> int:
        out = np.empty(1, dtype=np.uint32)
        _native.call("pb_cbloom_check_hashes", self._h, C.c_void_p(row.ctypes.data), 1, 0, C.c_void_p(out.ctypes.data), 0)
        return int(out[0])

    def check(self, key) -> int:
        """countingbloom.py:155-162"""
        return int(self.check_many([key])[0])

    def check_alt(self, hashes) -> int:
        """countingbloom.py:164-174"""
        return self._check_row(self._alt_row(hashes))

    def __contains__(self, key):
        return self.check(key)

    def remove(self, key, num_els: int = 1) -> int:
        """countingbloom.py:176-184: returns the key's smallest counter after the removal"""
        before = int(self.check_many([key])[0])
        if before == UINT32_T_MAX or before == 0:
            return before  # :199-202
        self.remove_many([key], num_els)
        return before - min(num_els, before)

    def remove_alt(self, hashes, num_els: int = 1) -> int:
        """countingbloom.py:186-208"""
        row = self._alt_row(hashes)
        before = self._check_row(row)
        if before == UINT32_T_MAX or before == 0:
            return before
        removed = C.c_uint64(0)
        _native.call("pb_cbloom_remove_hashes", self._h, C.c_void_p(row.ctypes.data), 1, 0, int(num_els), C.byref(removed))
        self._els_added -= removed.value
        return before - min(num_els, before)

    # ------------------------------------------------------------------ statistics / set algebra
    def estimate_elements(self) -> int:
        setbits = self._cnt_number_bits_set()
        if setbits >= self._num_bits:
            return -1
        log_n = math.log(1 - (float(setbits) / float(self._num_bits)))
        return int(-1 * (float(self._num_bits) / float(self._number_hashes)) * log_n)

    def export_size(self) -> int:
        return self._bloom_length * 4 + _FOOTER.size

    def __str__(self) -> str:
        """countingbloom.py:100-123"""
        cnt, total, largest, largest_idx = self._stats()
        return (
            "CountingBloom:\n"
            f"\tbits: {self.number_bits}\n"
            f"\testimated elements: {self.estimated_elements}\n"
            f"\tnumber hashes: {self.number_hashes}\n"
            f"\tmax false positive rate: {self.false_positive_rate:.6f}\n"
            f"\telements added: {self.elements_added}\n"
            f"\tcurrent false positive rate: {self.current_false_positive_rate():.6f}\n"
            f"\tis on disk: {'no' if self.is_on_disk is False else 'yes'}\n"
            f"\tindex fullness: {total / self.number_bits:.6}\n"  # (the reference divides the SUM of the counters, :103-107)
            f"\tmax index usage: {largest}\n"
            f"\tmax index id: {largest_idx}\n"
            f"\tcalculated elements: {total // self.number_hashes}\n"
        )

    def _check_second(self, second) -> None:
        if not isinstance(second, CountingBloomFilter):
            raise TypeError(_MISMATCH_MSG)
        if self._verify_bloom_similarity(second) is False:
            raise SimilarityError("Counting Bloom Filters are not similar enough to calculate similarity")

    def _combined(self, second, op: int) -> "CountingBloomFilter":
        self._check_second(second)
        res = CountingBloomFilter(self.estimated_elements, self.false_positive_rate,
                                  hash_function=self._hash_func if not self._fused else None, context=self._ctx)
        _native.call("pb_cbloom_combine", res._h, self._h, second._h, op)
        res.elements_added = res.estimate_elements()
        return res

    def intersection(self, second) -> "CountingBloomFilter":
        """countingbloom.py:210-243"""
        return self._combined(second, 1)

    def union(self, second) -> "CountingBloomFilter":
        """countingbloom.py:300-326"""
        return self._combined(second, 0)

    def jaccard_index(self, second) -> float:
        """countingbloom.py:245-272"""
        self._check_second(second)
        counts = (C.c_uint64 * 2)()
        _native.call("pb_cbloom_pair_counts", self._h, second._h, counts)
        if counts[0] == 0:
            return 1.0
        return counts[1] / counts[0]

    # ------------------------------------------------------------------ wire formats (counters as native 'I', then the Bloom footer)
    def _state_bytes(self) -> bytes:
        return self.bloom_numpy().tobytes()

    def export_hex(self) -> str:
        footer = _FOOTER_BE.pack(self._est_elements, self._els_added, self._fpr)
        # bloom.py:284: bytearray(array('I')) is the raw counter bytes (four per counter, native order)
        return str(hexlify(self._state_bytes()) + hexlify(footer), "utf-8")

    def export(self, file) -> None:
        from io import IOBase
        import mmap as _mmap

        if not isinstance(file, (IOBase, _mmap.mmap)):
            with open(Path(file).expanduser(), "wb") as fp:
                self.export(fp)
            return
        file.write(self._state_bytes())
        file.write(_FOOTER.pack(self._est_elements, self._els_added, self._fpr))

    def __bytes__(self) -> bytes:
        with BytesIO() as f:
            self.export(f)
            return f.getvalue()

    @classmethod
    def frombytes(cls, b, hash_function=None, **kw) -> "CountingBloomFilter":
        est, added, fpr = _FOOTER.unpack_from(bytes(b[-_FOOTER.size :]))
        blm = cls(est_elements=est, false_positive_rate=float(fpr), hash_function=hash_function, **kw)
        blm._load(bytes(b), blm.hash_function)
        return blm

    def _upload_counts(self, counts: np.ndarray) -> None:
        if counts.size != self._bloom_length:
            raise InitializationError("CountingBloom: stored counter array does not match its footer")
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        _native.call("pb_cbloom_upload", self._h, C.c_void_p(counts.ctypes.data), counts.size)

    def _load(self, data: bytes, hash_function=None) -> None:
        est, added, fpr = _FOOTER.unpack_from(data[-_FOOTER.size :])
        fpr, k, m = optimized_params(est, float(fpr))
        self._set_values(int(est), fpr, k, m, hash_function)
        self._upload_counts(np.frombuffer(data[: 4 * self._bloom_length], dtype=np.uint32))
        self._els_added = int(added)

    def _load_hex(self, hex_string: str, hash_function=None) -> None:
        off = _FOOTER_BE.size * 2
        est, added, fpr = _FOOTER_BE.unpack_from(unhexlify(hex_string[-off:]))
        fpr, k, m = optimized_params(est, float(fpr))
        self._set_values(int(est), fpr, k, m, hash_function)
        # bloom.py:511: array('I', raw bytes) -- four hex-decoded bytes per counter, native order
        raw = unhexlify(hex_string[:-off])
        self._upload_counts(np.frombuffer(raw, dtype=np.uint32))
        self._els_added = int(added)
