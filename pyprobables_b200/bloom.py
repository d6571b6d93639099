"""BloomFilter with device-resident state: the reference's class surface (probables/blooms/bloom.py:35-568)
plus the batch seam `add_many` / `check_many`.

Every add/check -- single key or batch, default hash or plugin hash -- runs in the CUDA kernels of
csrc/pb_bloom.cu through the C ABI; nothing is computed on the host except the float sizing formulas
(bloom.py:463-483 are float64 host math in the reference as well) and, for a custom `hash_function`,
the user's own callable.
"""

from __future__ import annotations

import ctypes as C
import sys
import math
import mmap as _mmap
import struct
from array import array
from binascii import hexlify, unhexlify
from io import BytesIO, IOBase
from numbers import Number
from pathlib import Path

import numpy as np

from . import _native
from .exceptions import InitializationError, SimilarityError
from .hashes import default_fnv_1a, is_default_hash
from .keys import pack_keys

_FOOTER = struct.Struct("QQf")  # est_elements, elements_added, fpr -- native order (bloom.py:108)
_FOOTER_BE = struct.Struct(">QQf")  # hex export is big-endian (bloom.py:109)
_U64_MASK = (1 << 64) - 1


def optimized_params(estimated_elements, false_positive_rate):
    """(f32-rounded fpr, number_hashes, number_bits) -- bloom.py:463-483: the FPR takes a float32 round
    trip "to mimic the c version", m = ceil(-n ln p / ln(2)^2), k = round(ln2 * m / n)."""
    if not (isinstance(estimated_elements, Number) and estimated_elements > 0):
        raise InitializationError("Bloom: estimated elements must be greater than 0")
    if not (isinstance(false_positive_rate, Number) and 0.0 <= false_positive_rate < 1.0):
        raise InitializationError("Bloom: false positive rate must be between 0.0 and 1.0")
    fpr32 = struct.unpack("f", struct.pack("f", float(false_positive_rate)))[0]
    n_bits = math.ceil((-estimated_elements * math.log(fpr32)) / 0.4804530139182)
    n_hashes = int(round(0.6931471805599453 * n_bits / estimated_elements))
    if n_hashes == 0:
        raise InitializationError("Bloom: Number hashes is zero; unusable parameters provided")
    return fpr32, n_hashes, n_bits


def _is_file(path) -> bool:
    return path is not None and Path(path).expanduser().is_file()


def _is_hex(s) -> bool:
    if not isinstance(s, str) or not s:
        return False
    try:
        int(s, 16)
        return True
    except ValueError:
        return False


class BloomFilter:
    """Bloom filter whose bit array lives in GPU memory.

    Args: est_elements, false_positive_rate, filepath, hex_string, hash_function -- as bloom.py:69-76.
    Extra keyword: device (CUDA ordinal, default 0), context (a pyprobables_b200 Context to share a stream).
    """

    def __init__(
        self,
        est_elements=None,
        false_positive_rate=None,
        filepath=None,
        hex_string=None,
        hash_function=None,
        *,
        device: int = 0,
        context=None,
    ):
        self._ctx_arg = (context, device)  # resolved in _set_values(): argument errors surface before any device work
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
                raise InitializationError("Insufecient parameters to set up the Bloom Filter")
            fpr, k, m = optimized_params(est_elements, false_positive_rate)
            self._set_values(est_elements, fpr, k, m, hash_function)

    # ------------------------------------------------------------------ setup / teardown
    def _set_values(self, est_els, fpr, n_hashes, n_bits, hash_func):
        self._est_elements = est_els
        self._fpr = fpr
        self._number_hashes = int(n_hashes)
        self._num_bits = int(n_bits)
        self._bloom_length = math.ceil(n_bits / 8.0)
        self._hash_func = hash_func if hash_func is not None else default_fnv_1a
        self._fused = is_default_hash(hash_func)
        self._els_added = 0
        if self._ctx is None:
            self._ctx = _native.default_context(self._ctx_arg[1])
        if self._h is not None:
            _native.lib().pb_bloom_destroy(self._h)
        h = C.c_void_p()
        _native.call("pb_bloom_create", self._ctx.handle, self._num_bits, self._number_hashes, C.byref(h))
        self._h = h

    def close(self) -> None:
        """free the device bitmap"""
        if getattr(self, "_h", None) is not None and _native._lib is not None:
            _native._lib.pb_bloom_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if not sys.is_finalizing():  # at interpreter exit the CUDA context may already be gone
                self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ properties (bloom.py:142-214)
    @property
    def false_positive_rate(self) -> float:
        return self._fpr

    @property
    def estimated_elements(self) -> int:
        return self._est_elements

    @property
    def number_hashes(self) -> int:
        return self._number_hashes

    @property
    def number_bits(self) -> int:
        return self._num_bits

    @property
    def elements_added(self) -> int:
        return self._els_added

    @elements_added.setter
    def elements_added(self, val: int):
        self._els_added = val

    @property
    def is_on_disk(self) -> bool:
        return self._on_disk

    @property
    def bloom_length(self) -> int:
        return self._bloom_length

    @property
    def bloom(self) -> array:
        """a host copy of the bit array as array('B') (what the reference exposes)"""
        return array("B", self.bloom_numpy().tobytes())

    @property
    def hash_function(self):
        return self._hash_func

    def bloom_numpy(self) -> np.ndarray:
        """host copy of the bit array as uint8[bloom_length]"""
        out = np.empty(self._bloom_length, dtype=np.uint8)
        _native.call("pb_bloom_download", self._h, C.c_void_p(out.ctypes.data), out.size)
        return out

    def device_ptr(self) -> int:
        p, n = C.c_void_p(), C.c_uint64()
        _native.call("pb_bloom_device_ptr", self._h, C.byref(p), C.byref(n))
        return p.value

    # ------------------------------------------------------------------ hot path
    def hashes(self, key, depth=None):
        """bloom.py:223-232"""
        return self._hash_func(key, depth if depth is not None else self._number_hashes)

    def _plugin_hashes(self, keys) -> tuple[np.ndarray, int]:
        k, m = self._number_hashes, self._num_bits
        rows = []
        for key in keys:
            hs = list(self._hash_func(key, k))[:k]
            if len(hs) < k:
                raise ValueError(f"hash_function returned {len(hs)} hashes, {k} needed")
            # values beyond u64 are reduced here; (h % m) % m == h % m keeps bloom.py:247 exact
            rows.append([h if 0 <= h <= _U64_MASK else h % m for h in hs])
        return np.asarray(rows, dtype=np.uint64).reshape(len(rows), k), len(rows)

    def add_many(self, keys) -> None:
        """BloomFilter.add (bloom.py:234-250) for every key of the batch"""
        if self._fused:
            kb = pack_keys(keys)
            if kb.n:
                _native.call("pb_bloom_add_keys", self._h, kb.ref())
            self._els_added += kb.n
        else:
            if isinstance(keys, (str, bytes, bytearray, memoryview)):
                keys = [keys]
            h, n = self._plugin_hashes(keys)
            self._add_hash_rows(h, n)

    def check_many(self, keys) -> np.ndarray:
        """BloomFilter.check (bloom.py:252-272) for every key -> bool[n] (a CUDA tensor of keys gets a CUDA bool
        tensor back: nothing crosses PCIe, and large batches of 16-byte keys take the partitioned query)"""
        if self._fused:
            kb = pack_keys(keys)
            if kb.on_device:
                import torch

                res = torch.empty(kb.n, dtype=torch.uint8, device=f"cuda:{self._ctx.device}")
                if kb.n:
                    _native.call("pb_bloom_check_keys", self._h, kb.ref(), C.c_void_p(res.data_ptr()), 1)
                    self._ctx.synchronize()
                return res.bool()
            out = np.empty(kb.n, dtype=np.uint8)
            if kb.n:
                _native.call("pb_bloom_check_keys", self._h, kb.ref(), C.c_void_p(out.ctypes.data), 0)
            return out.astype(bool)
        if isinstance(keys, (str, bytes, bytearray, memoryview)):
            keys = [keys]
        h, n = self._plugin_hashes(keys)
        return self._check_hash_rows(h, n)

    def _add_hash_rows(self, h: np.ndarray, n: int) -> None:
        if n:
            _native.call("pb_bloom_add_hashes", self._h, C.c_void_p(h.ctypes.data), n, 0)
        self._els_added += n

    def _check_hash_rows(self, h: np.ndarray, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint8)
        if n:
            _native.call("pb_bloom_check_hashes", self._h, C.c_void_p(h.ctypes.data), n, 0, C.c_void_p(out.ctypes.data), 0)
        return out.astype(bool)

    def add(self, key) -> None:
        """bloom.py:234-239"""
        self.add_many([key])

    def check(self, key) -> bool:
        """bloom.py:252-259"""
        return bool(self.check_many([key])[0])

    def __contains__(self, key) -> bool:
        return self.check(key)

    def _alt_row(self, hashes) -> np.ndarray:
        k, m = self._number_hashes, self._num_bits
        hs = list(hashes)[:k]
        if len(hs) < k:
            raise IndexError("list index out of range")  # what bloom.py:246 raises on a short list
        return np.asarray([h if 0 <= h <= _U64_MASK else h % m for h in hs], dtype=np.uint64).reshape(1, k)

    def add_alt(self, hashes) -> None:
        """bloom.py:241-250"""
        self._add_hash_rows(self._alt_row(hashes), 1)

    def check_alt(self, hashes) -> bool:
        """bloom.py:261-272"""
        return bool(self._check_hash_rows(self._alt_row(hashes), 1)[0])

    def clear(self) -> None:
        """bloom.py:217-221"""
        self._els_added = 0
        _native.call("pb_bloom_clear", self._h)

    # ------------------------------------------------------------------ statistics
    def _cnt_number_bits_set(self) -> int:
        """bloom.py:552-557 as a device popcount"""
        n = C.c_uint64()
        _native.call("pb_bloom_popcount", self._h, C.byref(n))
        return n.value

    def estimate_elements(self) -> int:
        """bloom.py:340-352"""
        setbits = self._cnt_number_bits_set()
        if setbits >= self._num_bits:
            return -1
        log_n = math.log(1 - (float(setbits) / float(self._num_bits)))
        return int(-1 * (float(self._num_bits) / float(self._number_hashes)) * log_n)

    def export_size(self) -> int:
        return self._bloom_length + _FOOTER.size

    def current_false_positive_rate(self) -> float:
        """bloom.py:361-369"""
        dbl = (self._number_hashes * -1 * self._els_added) / self._num_bits
        return math.pow((1 - math.exp(dbl)), self._number_hashes)

    def __str__(self) -> str:
        return (
            "BloomFilter:\n"
            f"\tbits: {self.number_bits}\n"
            f"\testimated elements: {self.estimated_elements}\n"
            f"\tnumber hashes: {self.number_hashes}\n"
            f"\tmax false positive rate: {self.false_positive_rate:.6f}\n"
            f"\tbloom length (8 bits): {self.bloom_length}\n"
            f"\telements added: {self.elements_added}\n"
            f"\testimated elements added: {self.estimate_elements()}\n"
            f"\tcurrent false positive rate: {self.current_false_positive_rate():.6f}\n"
            f"\texport size (bytes): {self.export_size()}\n"
            f"\tnumber bits set: {self._cnt_number_bits_set()}\n"
            f"\tis on disk: {'yes' if self.is_on_disk else 'no'}\n"
        )

    # ------------------------------------------------------------------ set algebra (bloom.py:371-460, :563-568)
    def _verify_bloom_similarity(self, second) -> bool:
        return not (
            self.number_hashes != second.number_hashes
            or self.number_bits != second.number_bits
            or self.hashes("test") != second.hashes("test")
        )

    def _check_second(self, second) -> None:
        if not isinstance(second, BloomFilter):
            raise TypeError("The parameter second must be of type BloomFilter or a BloomFilterOnDisk")
        if self._verify_bloom_similarity(second) is False:
            raise SimilarityError("Bloom Filters are not similar")

    def _combined(self, second, op: int) -> "BloomFilter":
        self._check_second(second)
        res = BloomFilter(self.estimated_elements, self.false_positive_rate, hash_function=self._hash_func
                          if not self._fused else None, context=self._ctx)
        res._hash_func, res._fused = self._hash_func, self._fused
        _native.call("pb_bloom_combine", res._h, self._h, second._h, op)
        res.elements_added = res.estimate_elements()
        return res

    def intersection(self, second) -> "BloomFilter":
        """bloom.py:371-399 (bitwise AND on the device)"""
        return self._combined(second, 1)

    def union(self, second) -> "BloomFilter":
        """bloom.py:401-428 (bitwise OR on the device)"""
        return self._combined(second, 0)

    def jaccard_index(self, second) -> float:
        """bloom.py:430-460"""
        self._check_second(second)
        counts = (C.c_uint64 * 2)()
        _native.call("pb_bloom_pair_popcounts", self._h, second._h, counts)
        if counts[0] == 0:
            return 1.0
        return counts[1] / counts[0]

    # ------------------------------------------------------------------ wire formats (bloom.py:274-338, :504-550)
    def export_hex(self) -> str:
        footer = _FOOTER_BE.pack(self._est_elements, self._els_added, self._fpr)
        return str(hexlify(self.bloom_numpy().tobytes()) + hexlify(footer), "utf-8")

    def export(self, file) -> None:
        if not isinstance(file, (IOBase, _mmap.mmap)):
            with open(Path(file).expanduser(), "wb") as fp:
                self.export(fp)
            return
        file.write(self.bloom_numpy().tobytes())
        file.write(_FOOTER.pack(self._est_elements, self._els_added, self._fpr))

    def export_c_header(self, filename) -> None:
        """bloom.py:306-322: the hex export as a C array plus the sizing constants"""
        from textwrap import wrap

        data = ("  " + line for line in wrap(", ".join(f"0x{e:02x}" for e in bytearray.fromhex(self.export_hex())), 80))
        bloom_type = "CountingBloomFilter" if type(self).__name__ == "CountingBloomFilter" else "standard BloomFilter"
        with open(filename, "w", encoding="utf-8") as file:
            print(f"/* BloomFilter Export of a {bloom_type} */", file=file)
            print("#include <inttypes.h>", file=file)
            print("const uint64_t estimated_elements = ", self.estimated_elements, ";", sep="", file=file)
            print("const uint64_t elements_added = ", self.elements_added, ";", sep="", file=file)
            print("const float false_positive_rate = ", self.false_positive_rate, ";", sep="", file=file)
            print("const uint64_t number_bits = ", self.number_bits, ";", sep="", file=file)
            print("const unsigned int number_hashes = ", self.number_hashes, ";", sep="", file=file)
            print("const unsigned char bloom[] = {", *data, "};", sep="\n", file=file)

    def __bytes__(self) -> bytes:
        with BytesIO() as f:
            self.export(f)
            return f.getvalue()

    @classmethod
    def frombytes(cls, b, hash_function=None, **kw) -> "BloomFilter":
        est, added, fpr = _FOOTER.unpack_from(bytes(b[-_FOOTER.size :]))
        blm = cls(est_elements=est, false_positive_rate=float(fpr), hash_function=hash_function, **kw)
        blm._load(bytes(b), blm.hash_function)
        return blm

    def _upload(self, raw: bytes) -> None:
        if len(raw) != self._bloom_length:
            raise InitializationError("Bloom: stored bit array does not match its footer")
        buf = np.frombuffer(raw, dtype=np.uint8)
        _native.call("pb_bloom_upload", self._h, C.c_void_p(buf.ctypes.data), buf.size)

    def _load(self, data: bytes, hash_function=None) -> None:
        est, added, fpr = _FOOTER.unpack_from(data[-_FOOTER.size :])
        fpr, k, m = optimized_params(est, float(fpr))
        self._set_values(int(est), fpr, k, m, hash_function)
        self._upload(data[: self._bloom_length])
        self._els_added = int(added)

    def _load_hex(self, hex_string: str, hash_function=None) -> None:
        off = _FOOTER_BE.size * 2
        est, added, fpr = _FOOTER_BE.unpack_from(unhexlify(hex_string[-off:]))
        fpr, k, m = optimized_params(est, float(fpr))
        self._set_values(int(est), fpr, k, m, hash_function)
        self._upload(unhexlify(hex_string[:-off]))
        self._els_added = int(added)
