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
from .keys import device_batch, pack_keys, slice_batch

_FOOTER = struct.Struct("IIq")  # width, depth, elements_added (countminsketch.py:122)
_U64_MASK = (1 << 64) - 1
_QUERY_CODE = {"min": 0, "mean": 1, "mean-min": 2}
INT32_T_MAX = (1 << 31) - 1
_RETURNS_CHUNK = 1 << 23  # keys per sort pass of add_many_returns


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

    # ------------------------------------------------------------------ ordered batches: what add() returns for every key
    def add_many_returns(self, keys, num_els=1):
        """`[self.add(k, n) for k in keys]` as one batch -> int64 CUDA tensor [n]: the estimate right after each key's own
        insertion (:267-288), earlier keys of the batch included.  The table ends up as after add_many.

        A counter's value after key i is its old value plus the num_els of all keys up to i that fall on it (saturating
        at INT32_MAX, which commutes with a running sum of non-negative amounts): per row a stable sort of the batch by
        counter and a segmented running sum give that value for every key at once (torch.sort / cumsum on the
        context's stream; the counter updates themselves are the add kernels of pb_cms.cu)."""
        import torch

        stream = torch.cuda.ExternalStream(self._ctx.stream, device=f"cuda:{self._ctx.device}")
        dev = f"cuda:{self._ctx.device}"
        d, w = self._depth, self._width
        if self._fused:
            kb = pack_keys(keys)
            n = kb.n
            dkb = device_batch(kb, self._ctx.device)
        else:
            if isinstance(keys, (str, bytes, bytearray, memoryview)):
                keys = [keys]
            h = self._hash_rows(keys)
            n = h.shape[0]
        scalar = isinstance(num_els, (int, np.integer))
        if scalar:
            if num_els < 0:
                raise NotSupportedError("add_many_returns takes non-negative num_els (removals are not order-free)")
        else:
            num_els = np.ascontiguousarray(num_els, dtype=np.int64)
            if num_els.shape != (n,):
                raise ValueError("num_els must be an int or one int per key")
            if n and num_els.min() < 0:
                raise NotSupportedError("add_many_returns takes non-negative num_els (removals are not order-free)")
        from .sharded import device_view

        with torch.cuda.stream(stream):
            out = torch.empty(n, dtype=torch.int64, device=dev)
            if n == 0:
                return out
            table = device_view(self.device_ptr(), d * w, "<i4", self._ctx.device)
            amounts = None if scalar else torch.from_numpy(num_els).to(dev)
            ea = max(INT64_T_MIN, min(INT64_T_MAX, self._elements_added))
            for lo in range(0, n, _RETURNS_CHUNK):
                hi = min(lo + _RETURNS_CHUNK, n)
                cn = hi - lo
                # counter columns of this chunk's keys (rows of d indices; the scratch stays at chunk size)
                if self._fused:
                    rows = torch.empty((cn, d), dtype=torch.int64, device=dev)
                    _native.call("pb_bloom_index_keys", self._ctx.handle, slice_batch(dkb, lo, hi).ref(), w, d, C.c_void_p(rows.data_ptr()))
                else:
                    rows = torch.from_numpy((h[lo:hi] % np.uint64(w)).view(np.int64)).to(dev)
                wv = torch.full((cn,), int(num_els), dtype=torch.int64, device=dev) if scalar else amounts[lo:hi]
                ar = torch.arange(cn, device=dev)
                vals = torch.empty((cn, d), dtype=torch.int64, device=dev)
                for r in range(d):
                    cs, order = torch.sort(rows[:, r].contiguous(), stable=True)
                    ws = wv[order]
                    csum = torch.cumsum(ws, 0)
                    start = torch.ones(cn, dtype=torch.bool, device=dev)
                    start[1:] = cs[1:] != cs[:-1]
                    first = torch.cummax(torch.where(start, ar, torch.zeros_like(ar)), 0).values  # start of each key's run
                    run = csum - (csum - ws)[first]  # amounts on this counter up to and including each key
                    vals[order, r] = torch.clamp(table[r * w + cs].to(torch.int64) + run, max=INT32_T_MAX)
                # the table itself moves through the add kernel (same end state as add_many, saturation included)
                ea_c = C.c_int64(0)
                _native.call("pb_cms_add_hashes", self._h, C.c_void_p(rows.data_ptr()), cn, 1,
                             None if scalar else C.c_void_p(wv.data_ptr()), int(num_els) if scalar else 0, C.byref(ea_c))
                ea_run = torch.clamp(torch.cumsum(wv, 0) + ea, max=INT64_T_MAX)  # (:283-286; int64 wrap is out of reach here)
                if self._query_type == "min":
                    res = vals.min(1).values
                elif self._query_type == "mean":
                    res = torch.div(vals.sum(1), d, rounding_mode="floor")
                else:  # mean-min, :438-453, with elements_added as it stands after this key
                    srt = torch.sort(vals, 1).values
                    calc = srt - torch.div(ea_run[:, None] - srt, w - 1, rounding_mode="floor")
                    calc = torch.sort(calc, 1).values
                    mid = torch.div(calc[:, d // 2] + calc[:, d // 2 - 1], 2, rounding_mode="floor") if d % 2 == 0 else calc[:, d // 2]
                    res = torch.where((srt[:, 0] == 0) & (srt[:, -1] == 0), torch.zeros_like(mid), mid)
                out[lo:hi] = res
                ea = int(ea_run[-1])
            self._ctx.synchronize()
        self._elements_added = ea
        return out

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

    def _add_alt_row(self, hashes, num_els: int) -> int:
        row = self._alt_row(hashes)
        ea = C.c_int64(self._elements_added)
        arr, scalar = self._num_els_args(num_els, 1)
        self._add_rows(row, arr, scalar, ea)
        self._elements_added = ea.value
        return int(self._check_rows(row)[0])

    def add_alt(self, hashes, num_els: int = 1) -> int:
        """:267-288"""
        return self._add_alt_row(hashes, num_els)

    def check_alt(self, hashes) -> int:
        """:332-340"""
        return int(self._check_rows(self._alt_row(hashes))[0])

    def remove(self, key, num_els: int = 1) -> int:
        """:290-299 (the lower clamp at INT32_MIN, :310-316, is what the device CAS path applies)"""
        return self.add(key, -int(num_els))

    def remove_alt(self, hashes, num_els: int = 1) -> int:
        return self._add_alt_row(hashes, -int(num_els))

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


def _batch_key_getter(keys):
    """index -> the object the reference would use as dictionary key for that element of the batch"""
    if isinstance(keys, (str, bytes, bytearray, memoryview)):
        return lambda i: keys
    if isinstance(keys, np.ndarray):
        return lambda i: keys[i].tobytes()
    if type(keys).__module__.startswith("torch"):
        host = {}

        def get(i):
            if "a" not in host:
                host["a"] = keys.cpu().numpy()
            return host["a"][i].tobytes()

        return get
    seq = keys if isinstance(keys, (list, tuple)) else list(keys)
    return lambda i: seq[i]


class _TrackingSketch(CountMinSketch):
    """shared batch plumbing of HeavyHitters and StreamThreshold: the per-key return values of an ordered batch
    (add_many_returns) and the few of them that can touch the dictionary, brought to the host in order"""

    def _returns_and_getter(self, keys, num_els):
        if not isinstance(keys, (str, bytes, bytearray, memoryview, np.ndarray, list, tuple)) and not type(keys).__module__.startswith("torch"):
            keys = list(keys)
        return self.add_many_returns(keys, num_els), _batch_key_getter(keys), keys

    @staticmethod
    def _result(res, keys):
        return res if type(keys).__module__.startswith("torch") and keys.is_cuda else res.cpu().numpy()


class HeavyHitters(_TrackingSketch):
    """countminsketch.py:532-697: a Count-Min sketch that tracks the `num_hitters` most frequent keys in a dictionary.
    `add_many` applies the reference's one-key-at-a-time bookkeeping (:617-661) to an ordered batch: the value add()
    would return is computed for every key on the device, and only the keys whose value reaches the smallest tracked
    count (a superset of those that can change the dictionary) are walked on the host, in order."""

    def __init__(self, num_hitters: int = 100, width=None, depth=None, confidence=None, error_rate=None, filepath=None,
                 hash_function=None, **kw):
        super().__init__(width, depth, confidence, error_rate, filepath, hash_function, **kw)
        self._top_x: dict = {}
        self._top_x_size = 0
        self._num_hitters = num_hitters
        self._smallest = 0

    @classmethod
    def frombytes(cls, b, num_hitters: int = 100, hash_function=None, **kw) -> "HeavyHitters":
        """:576-590"""
        width, depth, _ = _FOOTER.unpack(bytes(b[-_FOOTER.size :]))
        hh = cls(num_hitters=num_hitters, width=width, depth=depth, hash_function=hash_function, **kw)
        hh._parse_bytes(bytes(b))
        return hh

    def __str__(self) -> str:
        return (f"Heavy Hitters {super().__str__()}\n\tNumber Hitters: {self.number_heavy_hitters}\n"
                f"\tNumber Recorded: {self._top_x_size}")

    @property
    def heavy_hitters(self) -> dict:
        return self._top_x

    @property
    def number_heavy_hitters(self) -> int:
        return self._num_hitters

    def _track(self, key, res: int) -> None:
        """:644-660"""
        if self._top_x_size < self._num_hitters:
            tmp = self._top_x.get(key, None)
            self._top_x[key] = res
            if tmp is None:
                self._top_x_size = len(self._top_x)
        elif key in self._top_x:
            self._top_x[key] = res
        elif res > self._smallest:
            self._top_x[key] = res
            tmp_key = min(self._top_x, key=self._top_x.get)
            self._top_x.pop(tmp_key, None)
            new_min = min(self._top_x, key=self._top_x.get)
            self._smallest = self._top_x[new_min]

    def add(self, key, num_els: int = 1) -> int:
        """:617-627"""
        return self.add_alt(key, self.hashes(key), num_els)

    def add_alt(self, key, hashes, num_els: int = 1) -> int:
        """:629-661 (note the extra `key` argument, as in the reference)"""
        res = self._add_alt_row(hashes, num_els)
        self._track(key, res)
        return res

    def add_many(self, keys, num_els=1):
        """HeavyHitters.add for every key of the batch, in order -> the per-key return values"""
        import torch

        res, key_at, keys = self._returns_and_getter(keys, num_els)
        n = res.shape[0]
        is_tensor = type(keys).__module__.startswith("torch")
        as_rows = is_tensor or isinstance(keys, np.ndarray)
        pos, step = 0, 4096  # short first slices: the smallest tracked count rises fast and prunes what follows
        while pos < n:
            hi = min(pos + step, n)
            part = res[pos:hi]
            # a key can change the dictionary only if its value reaches the smallest tracked count as of now: tracked
            # keys never fall below it, and an untracked key must exceed it (:652); the bound only rises
            cand = torch.nonzero(part >= self._smallest).flatten()
            if cand.numel() and as_rows:
                at = cand + pos
                rows = keys[at.to(keys.device)].to(res.device) if is_tensor else torch.from_numpy(keys[at.cpu().numpy()]).to(res.device)
                self._replay_rows(rows, part[cand])
            elif cand.numel():
                for i, v in zip(cand.cpu().tolist(), part[cand].cpu().tolist()):
                    self._track(key_at(pos + i), v)
            pos, step = hi, min(step * 4, 1 << 22)
        return self._result(res, keys)

    def _replay_rows(self, rows, vals) -> None:
        """the bookkeeping of :644-660 for an ordered run of (key row, value) pairs, without a Python step per pair:
        pairs whose key is tracked only overwrite that key's value (the last one wins), so they are applied in bulk up to
        the next pair that changes WHICH keys are tracked; only those pairs go through `_track` one by one."""
        import torch

        length = rows.shape[1]
        uniq, inv = torch.unique(rows, dim=0, return_inverse=True)
        uniq_h, gid, val = uniq.cpu().numpy(), inv.cpu().numpy(), vals.cpu().numpy()
        tracked = np.zeros(uniq_h.shape[0], dtype=bool)  # per distinct key of this run: in the dictionary right now?
        cur = np.zeros(uniq_h.shape[0], dtype=np.int64)  # its latest value, written back before every structural step
        dirty = np.zeros(uniq_h.shape[0], dtype=bool)
        gid_of = {}
        for k in self._top_x:
            if isinstance(k, bytes) and len(k) == length:
                hit = torch.nonzero((uniq == torch.frombuffer(bytearray(k), dtype=torch.uint8).to(uniq.device)).all(1)).flatten()
                if hit.numel():
                    gid_of[k] = int(hit[0])
                    tracked[gid_of[k]] = True

        def flush():
            for k, g in gid_of.items():
                if dirty[g]:
                    self._top_x[k] = int(cur[g])
                    dirty[g] = False

        block = 8192
        for lo in range(0, gid.size, block):
            g, v = gid[lo : lo + block], val[lo : lo + block]
            while g.size:
                t = tracked[g]
                structural = ~t if self._top_x_size < self._num_hitters else ~t & (v > self._smallest)
                j = int(np.argmax(structural)) if structural.any() else g.size
                upd = t[:j]
                cur[g[:j][upd]] = v[:j][upd]  # repeated keys: the last value wins, as one-at-a-time would leave it
                dirty[g[:j][upd]] = True
                if j == g.size:
                    break
                flush()
                key = uniq_h[g[j]].tobytes()
                before = set(self._top_x)
                self._track(key, int(v[j]))
                for gone in before - set(self._top_x):
                    if gone in gid_of:
                        tracked[gid_of.pop(gone)] = False
                if key in self._top_x:
                    gid_of[key] = int(g[j])
                    tracked[g[j]] = True
                g, v = g[j + 1 :], v[j + 1 :]
        flush()

    def remove_alt(self, hashes, num_els: int = 1):
        """:663-676"""
        raise NotSupportedError("Unable to remove elements in the HeavyHitters class as it is an un supported action (and does not"
                                "make sense)!")

    def remove(self, key, num_els: int = 1):
        return self.remove_alt(self.hashes(key), num_els)

    def clear(self) -> None:
        """:678-683"""
        super().clear()
        self._top_x = {}
        self._top_x_size = 0
        self._smallest = 0

    def join(self, second) -> None:
        """:685-691"""
        raise NotSupportedError("Joining is not supported for heavy hitters")


class StreamThreshold(_TrackingSketch):
    """countminsketch.py:694-831: a Count-Min sketch with a dictionary of the keys whose estimate reached `threshold`"""

    def __init__(self, threshold: int = 100, width=None, depth=None, confidence=None, error_rate=None, filepath=None,
                 hash_function=None, **kw):
        super().__init__(width, depth, confidence, error_rate, filepath, hash_function, **kw)
        self._threshold = threshold
        self._meets_threshold: dict = {}

    @classmethod
    def frombytes(cls, b, threshold: int = 100, hash_function=None, **kw) -> "StreamThreshold":
        """:735-749"""
        width, depth, _ = _FOOTER.unpack(bytes(b[-_FOOTER.size :]))
        st = cls(threshold=threshold, width=width, depth=depth, hash_function=hash_function, **kw)
        st._parse_bytes(bytes(b))
        return st

    def __str__(self) -> str:
        return (f"Stream Threshold {super().__str__()}\n\tThreshold: {self.threshold}\n"
                f"\tNumber Meeting Threshold: {len(self._meets_threshold)}")

    @property
    def meets_threshold(self) -> dict:
        return self._meets_threshold

    @property
    def threshold(self) -> int:
        return self._threshold

    def clear(self) -> None:
        super().clear()
        self._meets_threshold = {}

    def add(self, key, num_els: int = 1) -> int:
        return self.add_alt(key, self.hashes(key), num_els)

    def add_alt(self, key, hashes, num_els: int = 1) -> int:
        """:787-803"""
        res = self._add_alt_row(hashes, num_els)
        if res >= self._threshold:
            self._meets_threshold[key] = res
        return res

    def add_many(self, keys, num_els=1):
        """StreamThreshold.add for every key of the batch, in order -> the per-key return values"""
        import torch

        res, key_at, keys = self._returns_and_getter(keys, num_els)
        cand = torch.nonzero(res >= self._threshold).flatten()
        if cand.numel():
            is_tensor = type(keys).__module__.startswith("torch")
            if (is_tensor or isinstance(keys, np.ndarray)) and cand.numel() > 256:
                # array batches: a key's values only rise, so its last write is its largest, and the dictionary keeps the
                # position of its first write -- one grouped pass on the device instead of a Python step per occurrence
                rows = keys[cand.to(keys.device)].to(res.device) if is_tensor else torch.from_numpy(keys[cand.cpu().numpy()]).to(res.device)
                uniq, inv = torch.unique(rows, dim=0, return_inverse=True)
                g = uniq.shape[0]
                best = torch.zeros(g, dtype=torch.int64, device=res.device).scatter_reduce_(0, inv, res[cand], "amax", include_self=False)
                first = torch.full((g,), cand.numel(), dtype=torch.int64, device=res.device).scatter_reduce_(
                    0, inv, torch.arange(cand.numel(), device=res.device), "amin", include_self=False)
                order = torch.argsort(first)
                uniq_h, best_h = uniq[order].cpu().numpy(), best[order].cpu().tolist()
                for row, v in zip(uniq_h, best_h):
                    self._meets_threshold[row.tobytes()] = v  # :801-802
            else:
                for i, v in zip(cand.cpu().tolist(), res[cand].cpu().tolist()):
                    self._meets_threshold[key_at(i)] = v  # :801-802
        return self._result(res, keys)

    def join(self, second) -> None:
        """:837-843"""
        raise NotSupportedError("Joining is not supported for stream threshold")

    def remove(self, key, num_els: int = 1) -> int:
        return self.remove_alt(key, self.hashes(key), num_els)

    def remove_alt(self, key, hashes, num_els: int = 1) -> int:
        """:818-831"""
        res = self._add_alt_row(hashes, -int(num_els))
        if res < self._threshold:
            self._meets_threshold.pop(key, None)
        else:
            self._meets_threshold[key] = res
        return res
