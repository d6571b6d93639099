"""Packing of key batches into the pb_keys layout (include/pb200.h).

A key is what the reference calls KeyT = str | bytes (probables/hashes.py:10).  The reference hashes a
`str` per code point -- `list(map(ord, key))`, hashes.py:98 -- and `bytes` per byte, so:
  * bytes / bytearray / memoryview      -> one u8 symbol per byte
  * str whose code points are all < 256  -> one u8 symbol per character (latin-1 bytes, NOT utf-8)
  * a batch holding a str with a code point >= 256 -> the whole batch is packed as u32 symbols
"""

from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import numpy as np

from ._native import pb_keys


class KeyBatch:
    """a packed batch; owns (references) the buffers the pb_keys struct points into"""

    __slots__ = ("c", "n", "_keep", "on_device")

    def __init__(self, data_ptr: int, offsets_ptr: int | None, n: int, stride: int, sym_width: int, on_device: bool, keep):
        self.c = pb_keys(data_ptr or None, offsets_ptr or None, n, stride, sym_width, 1 if on_device else 0, 0)
        self.n = n
        self.on_device = on_device
        self._keep = keep

    def ref(self):
        return C.byref(self.c)


def _is_torch_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _finish(data: np.ndarray, lens: np.ndarray, sym_width: int) -> KeyBatch:
    n = lens.size
    if sym_width == 1 and n and lens[0] > 0 and (lens == lens[0]).all():
        # equal-length byte keys: the fixed-stride layout (16-byte keys take the register fast path)
        return KeyBatch(data.ctypes.data, None, n, int(lens[0]), sym_width, False, (data,))
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    return KeyBatch(data.ctypes.data if data.size else None, offsets.ctypes.data, n, 0, sym_width, False, (data, offsets))


def _pack_sequence(keys: Sequence) -> KeyBatch:
    """Packs a list of keys.  The two homogeneous cases -- all bytes-like, all str -- are joined in C (one
    `join`, one `encode`), which is what keeps Python out of the way of the GPU; mixed lists take the
    per-element path."""
    n = len(keys)
    if n == 0:
        return _finish(np.zeros(0, dtype=np.uint8), np.zeros(0, dtype=np.uint64), 1)
    first = keys[0]
    try:
        if isinstance(first, (bytes, bytearray, memoryview)):
            joined = b"".join(keys)  # TypeError if a str hides in the list
            lens = np.fromiter(map(len, keys), dtype=np.uint64, count=n)
            if int(lens.sum()) == len(joined):  # (a memoryview of wider items would report items, not bytes)
                return _finish(np.frombuffer(joined, dtype=np.uint8), lens, 1)
        elif isinstance(first, str):
            text = "".join(keys)  # TypeError if a bytes object hides in the list
            lens = np.fromiter(map(len, keys), dtype=np.uint64, count=n)  # code points = symbols (hashes.py:98)
            if text.isascii():
                return _finish(np.frombuffer(text.encode("ascii"), dtype=np.uint8), lens, 1)
            try:
                return _finish(np.frombuffer(text.encode("latin-1"), dtype=np.uint8), lens, 1)
            except UnicodeEncodeError:
                data = np.frombuffer(text.encode("utf-32-le", "surrogatepass"), dtype="<u4")
                return _finish(data, lens, 4)
    except TypeError:
        pass
    return _pack_sequence_mixed(keys)


def _pack_sequence_mixed(keys: Sequence) -> KeyBatch:
    n = len(keys)
    lens = np.empty(n, dtype=np.uint64)
    parts: list[bytes] = []
    for i, k in enumerate(keys):
        if isinstance(k, str):
            try:
                b = k.encode("latin-1")
            except UnicodeEncodeError:
                return _pack_sequence_wide(keys)
        elif isinstance(k, (bytes, bytearray, memoryview)):
            b = bytes(k)
        else:
            raise TypeError(f"keys must be str or bytes-like, not {type(k).__name__}")
        parts.append(b)
        lens[i] = len(b)
    return _finish(np.frombuffer(b"".join(parts), dtype=np.uint8), lens, 1)


def _pack_sequence_wide(keys: Sequence) -> KeyBatch:
    n = len(keys)
    lens = np.empty(n, dtype=np.uint64)
    parts: list[np.ndarray] = []
    for i, k in enumerate(keys):
        if isinstance(k, str):
            a = np.frombuffer(k.encode("utf-32-le", "surrogatepass"), dtype="<u4")
        elif isinstance(k, (bytes, bytearray, memoryview)):
            a = np.frombuffer(bytes(k), dtype=np.uint8).astype("<u4")
        else:
            raise TypeError(f"keys must be str or bytes-like, not {type(k).__name__}")
        parts.append(a)
        lens[i] = a.size
    data = np.ascontiguousarray(np.concatenate(parts) if parts else np.zeros(0, dtype="<u4"), dtype="<u4")
    return _finish(data, lens, 4)


def pack_keys(keys, sync: bool = True) -> KeyBatch:
    """list/tuple of str|bytes, a single str|bytes (one key), a 2-D uint8 numpy array [n, L],
    a 2-D uint8 CUDA torch tensor [n, L] (zero-copy, device resident), an existing KeyBatch, or a
    (packed uint8 buffer, uint64 offsets[n+1]) pair -> KeyBatch.

    sync: a CUDA tensor may still be being written on torch's current stream while the engine reads it on its
    own stream; by default that stream is synchronized first.  Callers whose context runs ON torch's stream
    (pyprobables_b200.sharded) pass sync=False and stay asynchronous."""
    if isinstance(keys, KeyBatch):
        return keys
    if isinstance(keys, (str, bytes, bytearray, memoryview)):
        return _pack_sequence([keys])
    if isinstance(keys, np.ndarray):
        if keys.dtype != np.uint8 or keys.ndim != 2:
            raise TypeError("a numpy key batch must be a 2-D uint8 array [n_keys, key_len]")
        a = np.ascontiguousarray(keys)
        return KeyBatch(a.ctypes.data if a.size else None, None, a.shape[0], a.shape[1], 1, False, (a,))
    if _is_torch_tensor(keys):
        import torch

        if keys.dtype != torch.uint8 or keys.dim() != 2:
            raise TypeError("a torch key batch must be a 2-D uint8 tensor [n_keys, key_len]")
        t = keys.contiguous()
        if t.is_cuda:
            if sync:
                torch.cuda.current_stream(t.device).synchronize()
            return KeyBatch(t.data_ptr(), None, t.shape[0], t.shape[1], 1, True, (t,))
        return KeyBatch(t.data_ptr(), None, t.shape[0], t.shape[1], 1, False, (t,))
    if isinstance(keys, tuple) and len(keys) == 2 and isinstance(keys[0], np.ndarray) and isinstance(keys[1], np.ndarray):
        data = np.ascontiguousarray(keys[0], dtype=np.uint8)
        offsets = np.ascontiguousarray(keys[1], dtype=np.uint64)
        if offsets.ndim != 1 or offsets.size < 1:
            raise TypeError("offsets must be a 1-D uint64 array of n+1 entries")
        if offsets.size > 1 and (np.diff(offsets.astype(np.int64)) < 0).any():
            raise ValueError("offsets must be non-decreasing")
        if int(offsets[-1]) > data.size:
            raise ValueError("offsets run past the end of the packed buffer")
        return KeyBatch(data.ctypes.data if data.size else None, offsets.ctypes.data, offsets.size - 1, 0, 1, False, (data, offsets))
    if isinstance(keys, Iterable):
        return _pack_sequence(list(keys))
    raise TypeError(f"cannot interpret {type(keys).__name__} as a batch of keys")


def device_batch(kb: KeyBatch, device: int) -> KeyBatch:
    """a host-resident batch copied to device memory (same layout); device batches pass through.  For the entry points
    that take device keys only (pb_bloom_index_keys)."""
    if kb.on_device:
        return kb
    import torch

    dev = f"cuda:{device}"
    offs = None
    sw = int(kb.c.sym_width)
    first_sym, n_sym = 0, int(kb.c.stride) * kb.n
    if kb.c.offsets:
        o = np.frombuffer((C.c_uint64 * (kb.n + 1)).from_address(kb.c.offsets), dtype=np.uint64)
        first_sym, n_sym = int(o[0]), int(o[-1]) - int(o[0])  # (a slice of a larger batch starts somewhere inside `data`)
        offs = torch.from_numpy((o - o[0]).astype(np.int64)).to(dev)
    nbytes = n_sym * sw
    raw = (np.frombuffer((C.c_uint8 * nbytes).from_address(kb.c.data + first_sym * sw), dtype=np.uint8) if nbytes
           else np.zeros(0, np.uint8))
    data = torch.from_numpy(np.concatenate([raw, np.zeros(16, np.uint8)])).to(dev)  # (copy + tail padding)
    torch.cuda.current_stream(data.device).synchronize()
    return KeyBatch(data.data_ptr(), offs.data_ptr() if offs is not None else None, kb.n, int(kb.c.stride), int(kb.c.sym_width),
                    True, (data, offs))


def slice_batch(kb: KeyBatch, lo: int, hi: int) -> KeyBatch:
    """keys [lo, hi) of a packed batch as a batch of their own (a view: same buffers, same residency)"""
    lo = min(max(0, lo), kb.n)
    hi = max(lo, min(kb.n, hi))
    sw = int(kb.c.sym_width)
    if kb.c.offsets:  # offsets are absolute symbol positions in `data`: the window just starts later in the offsets array
        return KeyBatch(kb.c.data, kb.c.offsets + 8 * lo, hi - lo, 0, sw, kb.on_device, kb._keep)
    stride = int(kb.c.stride)
    return KeyBatch((kb.c.data or 0) + lo * stride * sw, None, hi - lo, stride, sw, kb.on_device, kb._keep)
