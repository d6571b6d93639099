"""ctypes/numpy front-end of the CPU oracle (oracle/pb_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of pb_oracle.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product package
(pyprobables_b200/) never does.

Parity status: pinned (tests/test_oracle.py: reference KATs + tests/golden/golden.json generated
from the pure-Python reference by tests/golden/make_golden.py).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"


def build(force: bool = False) -> Path:
    """compile pb_oracle.c with the system gcc (oracle/Makefile)"""
    src = _HERE / "pb_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


class _Keys(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("offsets", C.c_void_p),
        ("n", C.c_uint64),
        ("stride", C.c_uint32),
        ("sym_width", C.c_uint32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        u64, u32, i64, vp, i32 = C.c_uint64, C.c_uint32, C.c_int64, C.c_void_p, C.c_int
        KP = C.POINTER(_Keys)
        L.orc_fnv1a.restype = u64
        L.orc_fnv1a.argtypes = [vp, u64, u64]
        L.orc_fnv1a_u32.restype = u64
        L.orc_fnv1a_u32.argtypes = [vp, u64, u64]
        L.orc_default_fnv1a_many.restype = None
        L.orc_default_fnv1a_many.argtypes = [KP, u32, vp]
        L.orc_sm64.restype = u64
        L.orc_sm64.argtypes = [u64]
        L.orc_gen_uniform_keys.restype = None
        L.orc_gen_uniform_keys.argtypes = [u64, u64, u64, vp]
        L.orc_gen_rank_keys.restype = None
        L.orc_gen_rank_keys.argtypes = [vp, u64, vp]
        L.orc_bloom_add.restype = None
        L.orc_bloom_add.argtypes = [vp, u64, u32, KP]
        L.orc_bloom_check.restype = None
        L.orc_bloom_check.argtypes = [vp, u64, u32, KP, vp]
        L.orc_bloom_add_hashes.restype = None
        L.orc_bloom_add_hashes.argtypes = [vp, u64, u32, vp, u64]
        L.orc_bloom_check_hashes.restype = None
        L.orc_bloom_check_hashes.argtypes = [vp, u64, u32, vp, u64, vp]
        L.orc_popcount.restype = u64
        L.orc_popcount.argtypes = [vp, u64]
        L.orc_cbloom_add.restype = u64
        L.orc_cbloom_add.argtypes = [vp, u64, u32, C.POINTER(u64), KP, u64, vp]
        L.orc_cbloom_check.restype = None
        L.orc_cbloom_check.argtypes = [vp, u64, u32, KP, vp]
        L.orc_cbloom_remove.restype = None
        L.orc_cbloom_remove.argtypes = [vp, u64, u32, C.POINTER(i64), KP, u64, vp]
        L.orc_cms_add.restype = u64
        L.orc_cms_add.argtypes = [vp, u32, u32, C.POINTER(i64), KP, vp, i64, i32, vp]
        L.orc_cms_check.restype = None
        L.orc_cms_check.argtypes = [vp, u32, u32, i64, KP, i32, vp]
        L.orc_cms_add_parallel.restype = None
        L.orc_cms_add_parallel.argtypes = [vp, u32, u32, KP, C.c_int32]
        L.orc_cuckoo_new.restype = vp
        L.orc_cuckoo_new.argtypes = [u64, u32, u32, u32, u64]
        L.orc_cuckoo_free.restype = None
        L.orc_cuckoo_free.argtypes = [vp]
        L.orc_cuckoo_inserted.restype = u64
        L.orc_cuckoo_inserted.argtypes = [vp]
        L.orc_cuckoo_fingerprint_info.restype = None
        L.orc_cuckoo_fingerprint_info.argtypes = [KP, u64, u32, vp, vp, vp]
        L.orc_cuckoo_add.restype = u64
        L.orc_cuckoo_add.argtypes = [vp, KP, vp, u64]
        L.orc_cuckoo_check.restype = None
        L.orc_cuckoo_check.argtypes = [vp, KP, vp]
        L.orc_cuckoo_export_slots.restype = None
        L.orc_cuckoo_export_slots.argtypes = [vp, vp]
        L.orc_cuckoo_fingerprints.restype = u64
        L.orc_cuckoo_fingerprints.argtypes = [vp, vp]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


QUERY = {"min": 0, "mean": 1, "mean-min": 2}


class Keys:
    """a packed key batch; keeps the numpy buffers alive"""

    def __init__(self, data: np.ndarray, offsets: np.ndarray | None, n: int, stride: int, sym_width: int):
        self.data, self.offsets, self.n, self.stride, self.sym_width = data, offsets, n, stride, sym_width
        self.c = _Keys(
            data.ctypes.data if data.size else None,
            offsets.ctypes.data if offsets is not None else None,
            n,
            stride,
            sym_width,
        )

    def ref(self):
        return C.byref(self.c)


def pack(keys) -> Keys:
    """list of bytes/str, or a 2-D uint8 array of fixed-width keys -> Keys.
    str keys are hashed per code point (probables/hashes.py:98): ASCII strs pack as bytes,
    a batch holding any non-latin-1 str packs as u32 symbols."""
    if isinstance(keys, np.ndarray):
        a = np.ascontiguousarray(keys, dtype=np.uint8)
        assert a.ndim == 2
        return Keys(a, None, a.shape[0], a.shape[1], 1)
    if isinstance(keys, tuple) and len(keys) == 2 and isinstance(keys[0], np.ndarray) and isinstance(keys[1], np.ndarray):
        # (packed uint8 buffer, uint64 offsets[n+1]): variable-length byte keys already packed
        data = np.ascontiguousarray(keys[0], dtype=np.uint8)
        offs = np.ascontiguousarray(keys[1], dtype=np.uint64)
        return Keys(data, offs, offs.size - 1, 0, 1)
    syms = []
    wide = False
    for k in keys:
        if isinstance(k, str):
            s = [ord(ch) for ch in k]
            if s and max(s) > 255:
                wide = True
        else:
            s = list(bytes(k))
        syms.append(s)
    offs = np.zeros(len(syms) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(s) for s in syms], dtype=np.uint64) if syms else []
    flat = [x for s in syms for x in s]
    data = np.asarray(flat, dtype=np.uint32 if wide else np.uint8)
    return Keys(data, offs, len(syms), 0, 4 if wide else 1)


def fnv_1a(key, seed: int = 0) -> int:
    """probables/hashes.py:86-103"""
    k = pack([key])
    if k.sym_width == 4:
        return lib().orc_fnv1a_u32(k.c.data, k.data.size, seed)
    return lib().orc_fnv1a(k.c.data, k.data.size, seed)


def default_fnv_1a(key, depth: int = 1) -> list[int]:
    """probables/hashes.py:71-83"""
    return [fnv_1a(key, s) for s in range(depth)]


def default_fnv_1a_many(keys: Keys, depth: int) -> np.ndarray:
    out = np.empty((keys.n, depth), dtype=np.uint64)
    lib().orc_default_fnv1a_many(keys.ref(), depth, out.ctypes.data)
    return out


def uniform_keys(first: int, n: int, seed: int = 0xB200) -> np.ndarray:
    """key i = LE64(sm64(seed+2i)) || LE64(sm64(seed+2i+1)) -> uint8[n,16]"""
    out = np.empty((n, 16), dtype=np.uint8)
    lib().orc_gen_uniform_keys(seed, first, n, out.ctypes.data)
    return out


def rank_keys(ranks: np.ndarray) -> np.ndarray:
    """zipf key for rank r = LE64(r) || LE64(sm64(r)) -> uint8[n,16]"""
    r = np.ascontiguousarray(ranks, dtype=np.uint64)
    out = np.empty((r.size, 16), dtype=np.uint8)
    lib().orc_gen_rank_keys(r.ctypes.data, r.size, out.ctypes.data)
    return out


def bloom_params(est_elements: int, fpr: float) -> tuple[float, int, int, int]:
    """probables/blooms/bloom.py:463-483 + :495 -> (f32 fpr, k, num_bits, bloom_length).
    Host float math: restated in Python (it is Python in the reference too)."""
    import math
    import struct

    t_fpr = struct.unpack("f", struct.pack("f", float(fpr)))[0]
    m = math.ceil((-est_elements * math.log(t_fpr)) / 0.4804530139182)
    k = int(round(0.6931471805599453 * m / est_elements))
    return t_fpr, k, m, math.ceil(m / 8.0)


class Bloom:
    """probables/blooms/bloom.py:234-272 on a numpy byte array"""

    def __init__(self, num_bits: int, k: int):
        self.num_bits, self.k = num_bits, k
        self.bloom = np.zeros((num_bits + 7) // 8, dtype=np.uint8)
        self.elements_added = 0

    def add(self, keys: Keys):
        lib().orc_bloom_add(self.bloom.ctypes.data, self.num_bits, self.k, keys.ref())
        self.elements_added += keys.n

    def check(self, keys: Keys) -> np.ndarray:
        out = np.empty(keys.n, dtype=np.uint8)
        lib().orc_bloom_check(self.bloom.ctypes.data, self.num_bits, self.k, keys.ref(), out.ctypes.data)
        return out.astype(bool)

    def add_hashes(self, h: np.ndarray):
        h = np.ascontiguousarray(h, dtype=np.uint64).reshape(-1, self.k)
        lib().orc_bloom_add_hashes(self.bloom.ctypes.data, self.num_bits, self.k, h.ctypes.data, h.shape[0])
        self.elements_added += h.shape[0]

    def check_hashes(self, h: np.ndarray) -> np.ndarray:
        h = np.ascontiguousarray(h, dtype=np.uint64).reshape(-1, self.k)
        out = np.empty(h.shape[0], dtype=np.uint8)
        lib().orc_bloom_check_hashes(
            self.bloom.ctypes.data, self.num_bits, self.k, h.ctypes.data, h.shape[0], out.ctypes.data
        )
        return out.astype(bool)

    def popcount(self) -> int:
        return lib().orc_popcount(self.bloom.ctypes.data, self.bloom.size)


class ExpandingBloom:
    """probables/blooms/expandingbloom.py:149-183 one key at a time (pure-Python loop over precomputed hash rows:
    small cases only).  `max_queue_size` set = RotatingBloomFilter (:320-361)."""

    def __init__(self, est_elements: int, fpr: float, max_queue_size: int | None = None):
        self.est_elements, self.fpr = est_elements, fpr
        _, self.k, self.num_bits, _ = bloom_params(est_elements, fpr)
        self.max_queue_size = max_queue_size
        self.blooms: list[Bloom] = [Bloom(self.num_bits, self.k)]
        self.elements_added = 0

    def _has(self, blm: Bloom, idx) -> bool:  # bloom.py:261-272
        return all((blm.bloom[i >> 3] >> (i & 7)) & 1 for i in idx)

    def _grow_if_full(self):
        last = self.blooms[-1]
        if self.max_queue_size is None:
            if last.elements_added >= self.est_elements:  # :180-183
                self.blooms.append(Bloom(self.num_bits, self.k))
        elif last.elements_added == self.est_elements:  # :350
            if len(self.blooms) >= self.max_queue_size:  # :351, :359-361
                self.blooms.pop(0)
            self.blooms.append(Bloom(self.num_bits, self.k))

    def push(self):  # :126-128, :343-345
        if self.max_queue_size is not None and len(self.blooms) >= self.max_queue_size:
            self.blooms.pop(0)
        self.blooms.append(Bloom(self.num_bits, self.k))

    def add(self, keys: Keys, force: bool = False):
        rows = default_fnv_1a_many(keys, self.k) % np.uint64(self.num_bits)
        for row in rows.tolist():
            self.elements_added += 1  # :166
            if force or not any(self._has(b, row) for b in self.blooms):  # :167
                self._grow_if_full()  # :168
                last = self.blooms[-1]
                for i in row:  # bloom.py:241-250
                    last.bloom[i >> 3] |= 1 << (i & 7)
                last.elements_added += 1

    def check(self, keys: Keys) -> np.ndarray:  # :130-147
        rows = default_fnv_1a_many(keys, self.k) % np.uint64(self.num_bits)
        return np.array([any(self._has(b, row) for b in self.blooms) for row in rows.tolist()], dtype=bool)

    def export(self) -> bytes:  # :185-207
        import struct

        out = b"".join(struct.pack("Q", b.elements_added) + b.bloom.tobytes() for b in self.blooms)
        return out + struct.pack("QQQf", len(self.blooms), self.est_elements, self.elements_added, self.fpr)


class CountingBloom:
    """probables/blooms/countingbloom.py:125-208 on a numpy uint32 array (one counter per 'bit', so the array
    length equals number_bits, countingbloom.py:37)"""

    def __init__(self, num_bits: int, k: int):
        self.num_bits, self.k = num_bits, k
        self.bloom = np.zeros(num_bits, dtype=np.uint32)
        self.elements_added = 0

    def add(self, keys: Keys, num_els: int = 1) -> np.ndarray:
        ea = C.c_uint64(self.elements_added)
        post = np.empty(keys.n, dtype=np.uint64)
        clamped = lib().orc_cbloom_add(self.bloom.ctypes.data, self.num_bits, self.k, C.byref(ea), keys.ref(), num_els, post.ctypes.data)
        assert clamped == 0, "reference would raise OverflowError (array('I') past UINT32_MAX)"
        self.elements_added = ea.value
        return post

    def check(self, keys: Keys) -> np.ndarray:
        out = np.empty(keys.n, dtype=np.uint64)
        lib().orc_cbloom_check(self.bloom.ctypes.data, self.num_bits, self.k, keys.ref(), out.ctypes.data)
        return out

    def remove(self, keys: Keys, num_els: int = 1) -> np.ndarray:
        ea = C.c_int64(self.elements_added)
        post = np.empty(keys.n, dtype=np.uint64)
        lib().orc_cbloom_remove(self.bloom.ctypes.data, self.num_bits, self.k, C.byref(ea), keys.ref(), num_els, post.ctypes.data)
        self.elements_added = ea.value
        return post


class CMS:
    """probables/countminsketch/countminsketch.py:257-340, :429-453"""

    def __init__(self, width: int, depth: int, query_type: str = "min"):
        self.width, self.depth = width, depth
        self.bins = np.zeros(width * depth, dtype=np.int32)
        self.elements_added = 0
        self.query_type = query_type

    def add(self, keys: Keys, num_els=1, want_returns: bool = False):
        ea = C.c_int64(self.elements_added)
        arr = None
        scalar = 1
        if isinstance(num_els, np.ndarray):
            arr = np.ascontiguousarray(num_els, dtype=np.int64)
        else:
            scalar = int(num_els)
        post = np.empty(keys.n, dtype=np.int64) if want_returns else None
        under = lib().orc_cms_add(
            self.bins.ctypes.data,
            self.width,
            self.depth,
            C.byref(ea),
            keys.ref(),
            arr.ctypes.data if arr is not None else None,
            scalar,
            QUERY[self.query_type],
            post.ctypes.data if post is not None else None,
        )
        assert under == 0, "reference would raise OverflowError (bin < INT32_MIN)"
        self.elements_added = ea.value
        return post

    def add_parallel(self, keys: Keys, n_each: int = 1):
        lib().orc_cms_add_parallel(self.bins.ctypes.data, self.width, self.depth, keys.ref(), n_each)
        self.elements_added += keys.n * n_each

    def check(self, keys: Keys) -> np.ndarray:
        out = np.empty(keys.n, dtype=np.int64)
        lib().orc_cms_check(
            self.bins.ctypes.data,
            self.width,
            self.depth,
            self.elements_added,
            keys.ref(),
            QUERY[self.query_type],
            out.ctypes.data,
        )
        return out


class HeavyHitters(CMS):
    """countminsketch.py:617-661: the dictionary bookkeeping, one key at a time, over the sketch's per-key return values"""

    def __init__(self, num_hitters: int, width: int, depth: int):
        super().__init__(width, depth)
        self.num_hitters, self.top_x, self.top_x_size, self.smallest = num_hitters, {}, 0, 0

    def add_tracked(self, names: list, keys: Keys, num_els=1) -> np.ndarray:
        rets = self.add(keys, num_els, want_returns=True)
        for key, res in zip(names, rets.tolist()):
            if self.top_x_size < self.num_hitters:  # :645-649
                tmp = self.top_x.get(key)
                self.top_x[key] = res
                if tmp is None:
                    self.top_x_size = len(self.top_x)
            elif key in self.top_x:  # :650-651
                self.top_x[key] = res
            elif res > self.smallest:  # :652-659
                self.top_x[key] = res
                self.top_x.pop(min(self.top_x, key=self.top_x.get), None)
                self.smallest = self.top_x[min(self.top_x, key=self.top_x.get)]
        return rets


class StreamThreshold(CMS):
    """countminsketch.py:787-803"""

    def __init__(self, threshold: int, width: int, depth: int):
        super().__init__(width, depth)
        self.threshold, self.meets = threshold, {}

    def add_tracked(self, names: list, keys: Keys, num_els=1) -> np.ndarray:
        rets = self.add(keys, num_els, want_returns=True)
        for key, res in zip(names, rets.tolist()):
            if res >= self.threshold:
                self.meets[key] = res
        return rets


class Cuckoo:
    """probables/cuckoo/cuckoo.py:291-315, :361-392, :440-453, :483-506"""

    def __init__(self, capacity: int, bucket_size: int = 4, max_swaps: int = 500, fp_bits: int = 32, rng_seed=1):
        assert 1 <= bucket_size <= 255
        self.capacity, self.bucket_size, self.max_swaps, self.fp_bits = capacity, bucket_size, max_swaps, fp_bits
        self.h = lib().orc_cuckoo_new(capacity, bucket_size, max_swaps, fp_bits, rng_seed)
        assert self.h

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.orc_cuckoo_free(self.h)
            self.h = None

    @property
    def elements_added(self) -> int:
        return lib().orc_cuckoo_inserted(self.h)

    def fingerprint_info(self, keys: Keys):
        fp = np.empty(keys.n, dtype=np.uint32)
        i1 = np.empty(keys.n, dtype=np.uint64)
        i2 = np.empty(keys.n, dtype=np.uint64)
        lib().orc_cuckoo_fingerprint_info(
            keys.ref(), self.capacity, self.fp_bits, fp.ctypes.data, i1.ctypes.data, i2.ctypes.data
        )
        return i1, i2, fp

    def add(self, keys: Keys, failed_cap: int = 1024) -> np.ndarray:
        failed = np.zeros(failed_cap, dtype=np.uint32)
        n = lib().orc_cuckoo_add(self.h, keys.ref(), failed.ctypes.data, failed_cap)
        self.n_failed = n
        return failed[: min(n, failed_cap)]

    def check(self, keys: Keys) -> np.ndarray:
        out = np.empty(keys.n, dtype=np.uint8)
        lib().orc_cuckoo_check(self.h, keys.ref(), out.ctypes.data)
        return out.astype(bool)

    def export_slots(self) -> np.ndarray:
        out = np.empty(self.capacity * self.bucket_size, dtype=np.uint32)
        lib().orc_cuckoo_export_slots(self.h, out.ctypes.data)
        return out

    def fingerprints(self) -> np.ndarray:
        out = np.empty(self.capacity * self.bucket_size, dtype=np.uint32)
        n = lib().orc_cuckoo_fingerprints(self.h, out.ctypes.data)
        return np.sort(out[:n])


class CountingCuckoo:
    """probables/cuckoo/countingcuckoo.py:156-210, :230-265 one key at a time (pure-Python loop over precomputed
    fingerprints and bucket indices: small cases only).  Buckets are lists of [fingerprint, count]."""

    def __init__(self, capacity: int, bucket_size: int = 4, max_swaps: int = 500, fp_bits: int = 32, rng_seed: int = 1):
        import random

        self.capacity, self.bucket_size, self.max_swaps, self.fp_bits = capacity, bucket_size, max_swaps, fp_bits
        self.buckets: list[list[list[int]]] = [[] for _ in range(capacity)]
        self.elements_added = 0  # _inserted_elements
        self.unique_elements = 0
        self._rng = random.Random(rng_seed)
        self._info = Cuckoo(capacity, bucket_size, max_swaps, fp_bits)

    def _indices(self, fp: int):  # cuckoo.py:483-490
        return fp % self.capacity, fnv_1a(str(fp)) % self.capacity

    def _find(self, i1: int, i2: int, fp: int):  # :267-273
        for idx in (i1, i2):
            for b in self.buckets[idx]:
                if b[0] == fp:
                    return b
        return None

    def _place(self, idx: int, fp: int, count: int) -> bool:  # :318-323
        if len(self.buckets[idx]) < self.bucket_size:
            self.buckets[idx].append([fp, count])
            return True
        return False

    def _insert(self, fp: int, i1: int, i2: int):  # :230-265
        if self._place(i1, fp, 1) or self._place(i2, fp, 1):
            self.elements_added += 1
            self.unique_elements += 1
            return None
        idx = self._rng.choice([i1, i2])
        prv = [fp, 1]
        for _ in range(self.max_swaps):
            j = self._rng.randint(0, self.bucket_size - 1)
            prv, self.buckets[idx][j] = self.buckets[idx][j], prv
            a, b = self._indices(prv[0])
            idx = b if idx == a else a
            if self._place(idx, prv[0], prv[1]):
                self.elements_added += 1
                self.unique_elements += 1
                return None
        return prv

    def add(self, keys: Keys) -> list:
        """-> the homeless bins (empty when everything found a slot)"""
        i1, i2, fp = self._info.fingerprint_info(keys)
        homeless = []
        for a, b, f in zip(i1.tolist(), i2.tolist(), fp.tolist()):
            hit = self._find(a, b, f)  # :161-171
            if hit is not None:
                hit[1] += 1
                self.elements_added += 1
                continue
            left = self._insert(f, a, b)
            if left is not None:
                homeless.append(left)
        return homeless

    def check(self, keys: Keys) -> np.ndarray:  # :175-191
        i1, i2, fp = self._info.fingerprint_info(keys)
        out = np.zeros(keys.n, dtype=np.uint32)
        for n, (a, b, f) in enumerate(zip(i1.tolist(), i2.tolist(), fp.tolist())):
            hit = self._find(a, b, f)
            out[n] = hit[1] if hit is not None else 0
        return out

    def remove(self, keys: Keys) -> np.ndarray:  # :193-210
        i1, i2, fp = self._info.fingerprint_info(keys)
        out = np.zeros(keys.n, dtype=bool)
        for n, (a, b, f) in enumerate(zip(i1.tolist(), i2.tolist(), fp.tolist())):
            for idx in (a, b):
                hit = next((x for x in self.buckets[idx] if x[0] == f), None)
                if hit is not None:
                    hit[1] -= 1
                    self.elements_added -= 1
                    if hit[1] == 0:
                        self.buckets[idx].remove(hit)
                        self.unique_elements -= 1
                    out[n] = True
                    break
        return out

    def bins(self) -> list:
        return sorted((b[0], b[1]) for bucket in self.buckets for b in bucket)

    def export(self) -> bytes:  # :216-228, :325-334
        import struct

        out = bytearray()
        for bucket in self.buckets:
            for b in bucket:
                out += struct.pack("II", b[0], b[1])
            out += b"\0" * (8 * (self.bucket_size - len(bucket)))
        return bytes(out) + struct.pack("II", self.bucket_size, self.max_swaps)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_threads(n: int) -> None:
    lib().orc_set_threads(n)


if __name__ == "__main__":
    build(force=True)
    print("built", _LIB_PATH, "threads", num_threads(), "cpu_count", os.cpu_count())
