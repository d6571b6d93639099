"""Count-Min sketch family with a device-resident counter table: the reference's class surface
(probables/countminsketch/countminsketch.py:22-529) plus `add_many` / `check_many`.

Counters are int32[depth][width] row-major in HBM (20 MiB at the BASELINE config: L2 resident).  All adds
and queries run in csrc/pb_cms.cu through the C ABI.

Batch semantics: `add_many(keys, num_els)` leaves the table exactly as the reference's sequential
`for k in keys: add(k, n)` would whenever every n >= 0 (saturating addition at INT32_MAX is then
order-free) or no counter saturates; `add()` of a single key keeps the reference's return value.
"""

from __future__ import annotations

import ctypes as C
import sys
import math
import mmap as _mmap
import struct
from array import array
from io import BytesIO, IOBase
from numbers import Number
from pathlib import Path

import numpy as np

from . import _native
from ._limits import INT64_T_MAX, INT64_T_MIN
from .exceptions import CountMinSketchError, InitializationError, NotSupportedError
from .hashes import default_fnv_1a, is_default_hash
from .keys import pack_keys

_FOOTER = struct.Struct("IIq")  # width, depth, elements_added (countminsketch.py:122)
_U64_MASK = (1 << 64) - 1
_QUERY_CODE = {"min": 0, "mean": 1, "mean-min": 2}


class CountMinSketch:
    """Args: width, depth, confidence, error_rate, filepath, hash_function -- countminsketch.py:59-67.
    Extra keywords: device (CUDA ordinal), context."""

    def __init__(
        self,
        width=None,
        depth=None,
        confidence=None,
        error_rate=None,
        filepath=None,
        hash_function=None,
        *,
        device: int = 0,
        context=None,
    ):
        self._ctx_arg = (context, device)  # resolved in _create(): argument errors surface before any device work
        self._ctx = context
        self._h = None
        self._elements_added = 0
        self._query_type = "min"
        self._hash_function = hash_function if hash_function is not None else default_fnv_1a
        self._fused = is_default_hash(hash_function)
        if filepath is not None and Path(filepath).expanduser().is_file():
            self._parse_bytes(Path(filepath).expanduser().read_bytes())
            return
        if width is not None and depth is not None:
            if not (isinstance(width, Number) and width > 0 and isinstance(depth, Number) and depth > 0):
                raise InitializationError("CountMinSketch: width and depth must be greater than 0")
            self._width, self._depth = int(width), int(depth)
            self._confidence = 1 - (1 / math.pow(2, self._depth))
            self._error_rate = 2 / self._width
        elif confidence is not None and error_rate is not None:
            ok = isinstance(confidence, Number) and confidence > 0 and isinstance(error_rate, Number) and error_rate > 0
            if not ok:
                raise InitializationError("CountMinSketch: width and depth must be greater than 0")
            self._confidence, self._error_rate = confidence, error_rate
            self._width = math.ceil(2 / error_rate)
            self._depth = math.ceil((-1 * math.log(1 - confidence)) / 0.6931471805599453)
        else:
            raise InitializationError(
                "Must provide one of the following to initialize the Count-Min Sketch:\n"
                "    A file to load,\n"
                "    The width and depth,\n"
                "    OR confidence and error rate"
            )
        self._create()

    def _create(self) -> None:
        if self._ctx is None:
            self._ctx = _native.default_context(self._ctx_arg[1])
        if self._h is not None:
            _native.lib().pb_cms_destroy(self._h)
        h = C.c_void_p()
        _native.call("pb_cms_create", self._ctx.handle, self._width, self._depth, C.byref(h))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and _native._lib is not None:
            _native._lib.pb_cms_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if not sys.is_finalizing():  # at interpreter exit the CUDA context may already be gone
                self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ properties (:168-238)
    @property
    def width(self) -> int:
        return self._width

    @property
    def depth(self) -> int:
        return self._depth

    @property
    def confidence(self) -> float:
        return self._confidence

    @property
    def error_rate(self) -> float:
        return self._error_rate

    @property
    def elements_added(self) -> int:
        return self._elements_added

    @property
    def query_type(self) -> str:
        return self._query_type

    @query_type.setter
    def query_type(self, val):
        val = val.lower() if val is not None else "min"
        self._query_type = val if val in ("mean", "mean-min") else "min"

    @property
    def hash_function(self):
        return self._hash_function

    @property
    def _bins(self) -> array:
        """host copy of the counters as array('i') (the reference's attribute)"""
        return array("i", self.bins_numpy().tobytes())

    def bins_numpy(self) -> np.ndarray:
        out = np.empty(self._width * self._depth, dtype=np.int32)
        _native.call("pb_cms_download", self._h, C.c_void_p(out.ctypes.data), out.size)
        return out

    def device_ptr(self) -> int:
        p, n = C.c_void_p(), C.c_uint64()
        _native.call("pb_cms_device_ptr", self._h, C.byref(p), C.byref(n))
        return p.value

    def __str__(self) -> str:
        return (
            "Count-Min Sketch:\n"
            f"\tWidth: {self.width}\n"
            f"\tDepth: {self.depth}\n"
            f"\tConfidence: {self.confidence}\n"
            f"\tError Rate: {self.error_rate}\n"
            f"\tElements Added: {self.elements_added}"
        )

    def __contains__(self, key) -> bool:
        return self.check(key) != 0

    def clear(self) -> None:
        """:240-244"""
        self._elements_added = 0
        _native.call("pb_cms_clear", self._h)

    # ------------------------------------------------------------------ hot path
    def hashes(self, key, depth=None):
        """:246-255"""
        return self._hash_function(key, self._depth if depth is None else depth)

    def _hash_rows(self, keys) -> np.ndarray:
        d, w = self._depth, self._width
        rows = []
        for key in keys:
            hs = list(self._hash_function(key, d))[:d]
            if len(hs) < d:
                raise ValueError(f"hash_function returned {len(hs)} hashes, {d} needed")
            rows.append([h if 0 <= h <= _U64_MASK else h % w for h in hs])
        return np.asarray(rows, dtype=np.uint64).reshape(len(rows), d)

    @staticmethod
    def _num_els_args(num_els, n):
        if isinstance(num_els, (int, np.integer)):
            # beyond int64 nothing changes any more: counters saturate at INT32_MAX and elements_added at INT64_MAX
            return None, max(INT64_T_MIN, min(INT64_T_MAX, int(num_els)))
        arr = np.ascontiguousarray(num_els, dtype=np.int64)
        if arr.shape != (n,):
            raise ValueError("num_els must be an int or one int per key")
        return arr, 0

    def add_many(self, keys, num_els=1) -> None:
        """CountMinSketch.add (:257-288) for every key; num_els is an int or an int64 array (one per key)"""
        ea = C.c_int64(max(INT64_T_MIN, min(INT64_T_MAX, self._elements_added)))
        if self._fused:
            kb = pack_keys(keys)
            if kb.n == 0:
                return
            if kb.on_device and not isinstance(num_els, (int, np.integer)):
                raise NotSupportedError("device-resident keys take a scalar num_els")
            arr, scalar = self._num_els_args(num_els, kb.n)
            _native.call(
                "pb_cms_add_keys", self._h, kb.ref(), C.c_void_p(arr.ctypes.data) if arr is not None else None, scalar, C.byref(ea)
            )
        else:
            if isinstance(keys, (str, bytes, bytearray, memoryview)):
                keys = [keys]
            h = self._hash_rows(keys)
            if h.shape[0] == 0:
                return
            arr, scalar = self._num_els_args(num_els, h.shape[0])
            self._add_rows(h, arr, scalar, ea)
        self._elements_added = ea.value

    def _add_rows(self, h, arr, scalar, ea) -> None:
        _native.call(
            "pb_cms_add_hashes",
            self._h,
            C.c_void_p(h.ctypes.data),
            h.shape[0],
            0,
            C.c_void_p(arr.ctypes.data) if arr is not None else None,
            scalar,
            C.byref(ea),
        )

    def check_many(self, keys) -> np.ndarray:
        """CountMinSketch.check (:323-340) for every key -> int64[n] with the current query_type"""
        qt = _QUERY_CODE[self._query_type]
        if self._fused:
            kb = pack_keys(keys)
            out = np.empty(kb.n, dtype=np.int64)
            if kb.n:
                _native.call("pb_cms_check_keys", self._h, kb.ref(), qt, self._elements_added, C.c_void_p(out.ctypes.data), 0)
            return out
        if isinstance(keys, (str, bytes, bytearray, memoryview)):
            keys = [keys]
        return self._check_rows(self._hash_rows(keys))

    def _check_rows(self, h: np.ndarray) -> np.ndarray:
        out = np.empty(h.shape[0], dtype=np.int64)
        if h.shape[0]:
            _native.call(
                "pb_cms_check_hashes",
                self._h,
                C.c_void_p(h.ctypes.data),
                h.shape[0],
                0,
                _QUERY_CODE[self._query_type],
                self._elements_added,
                C.c_void_p(out.ctypes.data),
                0,
            )
        return out

    def add(self, key, num_els: int = 1) -> int:
        """:257-265 -- returns the estimate after the insertion"""
        self.add_many([key], num_els)
        return int(self.check_many([key])[0])

    def check(self, key) -> int:
        """:323-330"""
        return int(self.check_many([key])[0])

    def _alt_row(self, hashes) -> np.ndarray:
        d, w = self._depth, self._width
        hs = list(hashes)[:d]
        if len(hs) < d:
            raise IndexError("list index out of range")
        return np.asarray([h if 0 <= h <= _U64_MASK else h % w for h in hs], dtype=np.uint64).reshape(1, d)

    def add_alt(self, hashes, num_els: int = 1) -> int:
        """:267-288"""
        row = self._alt_row(hashes)
        ea = C.c_int64(self._elements_added)
        arr, scalar = self._num_els_args(num_els, 1)
        self._add_rows(row, arr, scalar, ea)
        self._elements_added = ea.value
        return int(self._check_rows(row)[0])

    def check_alt(self, hashes) -> int:
        """:332-340"""
        return int(self._check_rows(self._alt_row(hashes))[0])

    def remove(self, key, num_els: int = 1) -> int:
        """:290-299 (the lower clamp at INT32_MIN, :310-316, is what the device CAS path applies)"""
        return self.add(key, -int(num_els))

    def remove_alt(self, hashes, num_els: int = 1) -> int:
        return self.add_alt(hashes, -int(num_els))

    # ------------------------------------------------------------------ merge (:356-399)
    def join(self, second: "CountMinSketch") -> None:
        if not isinstance(second, CountMinSketch):
            raise TypeError(f"Unable to merge a count-min sketch with {type(second)}")
        if self.width != second.width or self.depth != second.depth or self.hashes("test") != second.hashes("test"):
            raise CountMinSketchError("Unable to merge as the count-min sketches are mismatched")
        if second._ctx.device != self._ctx.device:
            raise NotSupportedError("join needs both sketches on the same device")
        second._ctx.synchronize()
        _native.call("pb_cms_join_buffer", self._h, C.c_void_p(second.device_ptr()), self._width * self._depth)
        self._ctx.synchronize()
        self._elements_added = max(INT64_T_MIN, min(INT64_T_MAX, self._elements_added + second.elements_added))

    # ------------------------------------------------------------------ wire format (:342-354, :401-427)
    def export(self, file) -> None:
        if not isinstance(file, (IOBase, _mmap.mmap)):
            with open(Path(file).expanduser(), "wb") as fp:
                self.export(fp)
            return
        file.write(self.bins_numpy().tobytes())
        file.write(_FOOTER.pack(self._width, self._depth, self._elements_added))

    def __bytes__(self) -> bytes:
        with BytesIO() as f:
            self.export(f)
            return f.getvalue()

    @classmethod
    def frombytes(cls, b, hash_function=None, **kw):
        width, depth, _ = _FOOTER.unpack_from(bytes(b[-_FOOTER.size :]))
        cms = cls(width=width, depth=depth, hash_function=hash_function, **kw)
        cms._parse_bytes(bytes(b))
        return cms

    def _parse_bytes(self, data: bytes) -> None:
        width, depth, added = _FOOTER.unpack_from(data[-_FOOTER.size :])
        self._width, self._depth, self._elements_added = int(width), int(depth), int(added)
        self._confidence = 1 - (1 / math.pow(2, self._depth))
        self._error_rate = 2 / self._width
        self._create()
        bins = np.frombuffer(data[: 4 * self._width * self._depth], dtype=np.int32)
        if bins.size != self._width * self._depth:
            raise InitializationError("CountMinSketch: stored counters do not match the footer")
        _native.call("pb_cms_upload", self._h, C.c_void_p(bins.ctypes.data), bins.size)


class CountMeanSketch(CountMinSketch):
    """query_type 'mean' (:456-491)"""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.query_type = "mean"


class CountMeanMinSketch(CountMinSketch):
    """query_type 'mean-min' (:494-529)"""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.query_type = "mean-min"
