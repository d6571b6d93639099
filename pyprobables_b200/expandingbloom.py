"""ExpandingBloomFilter / RotatingBloomFilter on the device: a stack of pyprobables_b200.BloomFilter bitmaps with the
reference's add / check semantics (probables/blooms/expandingbloom.py).

The reference adds one key at a time: count it (:166, :327), look it up in every filter of the stack, and if it is
not found (or `force`) make room -- a new filter when the newest one holds `est_elements` keys (:180-183), dropping the
oldest one in a full rotating queue (:347-361) -- and add it to the newest filter (:167-169).  Whether key i is added
depends on the keys before it, so a batch is not a plain scatter.  `add_many` gives exactly the one-at-a-time result:

  * rows = the k bit indices of every key (pb_bloom_index_keys, one pass of the hash kernels);
  * rows found in an older filter of the stack are out (those filters do not change while the newest has room);
  * in the newest filter a key that is found adds no bit, so the bits set before row i's turn are the filter's bits plus
    the bits of all earlier rows still in play: row i is added iff it holds a clear bit that no earlier row touches
    (pb_bloom_novel_rows: atomicMin of the row number per bit, then one compare);
  * a prefix sum over the "added" flags finds the row that fills the newest filter; rows up to it are applied
    (pb_bloom_add_rows), and the rest of the batch is taken up again -- the next row that no filter finds grows the
    stack first, exactly where the reference would.

torch is used for the glue on the context's stream (prefix sums, slices); the hashing, lookups and bit updates are
the library's kernels.
"""

from __future__ import annotations

import ctypes as C
import struct
from io import BytesIO, IOBase
from mmap import mmap
from pathlib import Path

import numpy as np

from . import _native
from .bloom import BloomFilter, _is_file
from .exceptions import RotatingBloomFilterError
from .hashes import default_fnv_1a, is_default_hash
from .keys import device_batch, pack_keys, slice_batch

_FOOTER = struct.Struct("QQQf")  # expandingbloom.py:72
_U64 = struct.Struct("Q")  # :73
_CHUNK_ROWS = 1 << 24  # rows of one pb_bloom_novel_rows call (24-bit row numbers in its table)
_BLOCK_KEYS = 1 << 27  # keys hashed to rows of bit indices at once: bounds the scratch (keys x k x 8 bytes) whatever the batch size


class ExpandingBloomFilter:
    """expandingbloom.py:20-247.  Extra keywords: device, context (as BloomFilter)."""

    def __init__(self, est_elements=None, false_positive_rate=None, filepath=None, hash_function=None, *, device: int = 0,
                 context=None):
        self._blooms: list[BloomFilter] = []
        self._fpr = false_positive_rate if false_positive_rate is not None else 0.0  # :54
        self._est_elements = est_elements if est_elements is not None else 100  # :55
        self._hash_func = hash_function if hash_function is not None else default_fnv_1a
        self._fused = is_default_hash(hash_function)
        self._added_elements = 0
        self._ctx = context if context is not None else _native.default_context(device)
        if _is_file(filepath):
            self._load(Path(filepath).expanduser().read_bytes())
        else:
            self._add_bloom_filter()  # :69

    # ------------------------------------------------------------------ properties (:101-124)
    @classmethod
    def frombytes(cls, b, hash_function=None, **kw) -> "ExpandingBloomFilter":
        """:75-88"""
        size, est_els, added_els, fpr = cls._parse_footer(b)
        blm = cls(est_elements=est_els, false_positive_rate=fpr, hash_function=hash_function, **kw)
        blm._parse_blooms(bytes(b), size)
        blm._added_elements = added_els
        return blm

    def __contains__(self, key) -> bool:
        return self.check(key)

    def __bytes__(self) -> bytes:
        with BytesIO() as f:
            self.export(f)
            return f.getvalue()

    @property
    def expansions(self) -> int:
        return len(self._blooms) - 1

    @property
    def false_positive_rate(self) -> float:
        return self._fpr

    @property
    def estimated_elements(self) -> int:
        return self._est_elements

    @property
    def elements_added(self) -> int:
        return self._added_elements

    @property
    def hash_function(self):
        return self._hash_func

    def push(self) -> None:
        """:126-128"""
        self._add_bloom_filter()

    # ------------------------------------------------------------------ the stack
    def _new_bloom(self) -> BloomFilter:
        return BloomFilter(est_elements=self._est_elements, false_positive_rate=self._fpr,
                           hash_function=None if self._fused else self._hash_func, context=self._ctx)

    def _add_bloom_filter(self) -> None:
        """:171-178"""
        blm = self._new_bloom()
        if self._blooms:  # the first-setter table of the batch kernels goes with the newest filter
            _native.call("pb_bloom_move_scratch", self._blooms[-1]._h, blm._h)
        self._blooms.append(blm)

    def _newest_is_full(self) -> bool:
        """:180-183"""
        return self._blooms[-1].elements_added >= self._est_elements

    def _room(self):
        """keys the newest filter still takes before the stack grows (None: no limit)"""
        return max(self._est_elements - self._blooms[-1].elements_added, 0)

    def _grow(self) -> None:
        self._add_bloom_filter()

    # ------------------------------------------------------------------ rows of bit indices
    def _torch(self):
        import torch

        return torch, torch.cuda.ExternalStream(self._ctx.stream, device=f"cuda:{self._ctx.device}")

    def _row_blocks(self, keys, block_keys: int = _BLOCK_KEYS):
        """the batch in key order as blocks of at most `block_keys` rows of bit indices (bounds the scratch at
        block_keys x k x 8 bytes whatever the batch size); yields (idx, n_rows, keys_were_on_device)"""
        torch, stream = self._torch()
        first = self._blooms[0]
        k, m = first.number_hashes, first.number_bits
        dev = f"cuda:{self._ctx.device}"
        if self._fused:
            kb = pack_keys(keys)
            was_dev = kb.on_device
            for lo in range(0, max(kb.n, 1), block_keys):
                part = slice_batch(kb, lo, lo + block_keys)
                dpart = device_batch(part, self._ctx.device)
                with torch.cuda.stream(stream):
                    idx = torch.empty((part.n, k), dtype=torch.int64, device=dev)
                    if part.n:
                        _native.call("pb_bloom_index_keys", self._ctx.handle, dpart.ref(), m, k, C.c_void_p(idx.data_ptr()))
                        self._ctx.synchronize()  # dpart's buffers may go away with this frame
                yield idx, part.n, was_dev
            return
        if isinstance(keys, (str, bytes, bytearray, memoryview)):
            keys = [keys]
        keys = list(keys)
        for lo in range(0, max(len(keys), 1), block_keys):
            h, n = first._plugin_hashes(keys[lo : lo + block_keys])
            yield self._rows_from_hashes(h % np.uint64(m)), n, False

    def _rows_from_hashes(self, h: np.ndarray):
        torch, stream = self._torch()
        with torch.cuda.stream(stream):
            return torch.from_numpy(np.ascontiguousarray(h).view(np.int64)).to(f"cuda:{self._ctx.device}")

    def _found(self, blooms, idx):
        """uint8 CUDA tensor [n]: 1 where a filter of `blooms` holds every bit of the row (:140-147)"""
        torch, _ = self._torch()
        n = idx.shape[0]
        found = torch.zeros(n, dtype=torch.uint8, device=idx.device)
        if n and blooms:
            handles = (C.c_void_p * len(blooms))(*[b._h.value for b in blooms])
            _native.call("pb_bloom_rows_in_any", handles, len(blooms), C.c_void_p(idx.data_ptr()), n, C.c_void_p(found.data_ptr()))
        return found

    # ------------------------------------------------------------------ check (:130-147)
    def check_many(self, keys):
        """ExpandingBloomFilter.check for every key -> bool[n] (a CUDA tensor of keys gets a CUDA tensor back)"""
        torch, stream = self._torch()
        parts, was_dev = [], False
        for idx, _, was_dev in self._row_blocks(keys):
            with torch.cuda.stream(stream):
                parts.append(self._found(self._blooms, idx).bool())
                self._ctx.synchronize()
        found = torch.cat(parts) if len(parts) != 1 else parts[0]
        return found if was_dev else found.cpu().numpy()

    def check(self, key) -> bool:
        return bool(self.check_many([key])[0])

    def check_alt(self, hashes) -> bool:
        first = self._blooms[0]
        torch, stream = self._torch()
        idx = self._rows_from_hashes(first._alt_row(hashes) % np.uint64(first.number_bits))
        with torch.cuda.stream(stream):
            res = bool(self._found(self._blooms, idx)[0])
        return res

    # ------------------------------------------------------------------ add (:149-169)
    def add_many(self, keys, force: bool = False) -> None:
        """ExpandingBloomFilter.add for every key of the batch, in order; the result (bitmaps, per-filter counts, number
        of filters) is the one the reference reaches adding the keys one at a time"""
        for idx, n, _ in self._row_blocks(keys):
            self._add_rows(idx, n, force)

    def add(self, key, force: bool = False) -> None:
        self.add_many([key], force)

    def add_alt(self, hashes, force: bool = False) -> None:
        first = self._blooms[0]
        self._add_rows(self._rows_from_hashes(first._alt_row(hashes) % np.uint64(first.number_bits)), 1, force)

    def _add_rows(self, idx, n: int, force: bool) -> None:
        torch, stream = self._torch()
        self._added_elements += n  # :166 -- counted whether or not the key ends up in a filter
        pos = 0
        with torch.cuda.stream(stream):
            while pos < n:
                newest = self._blooms[-1]
                room = self._room()
                # no more than `room` rows of a slice can be applied before the stack has to grow, and the rows after
                # that point are taken up again anyway: size the slice to the room (plus the few rows that get skipped)
                take_rows = _CHUNK_ROWS if not room else min(_CHUNK_ROWS, room + room // 4 + 4096)
                rows = idx[pos : pos + take_rows]
                cnt = rows.shape[0]
                if force:
                    if room == 0:  # :168 -- the next key grows the stack whatever it is
                        self._grow()
                        continue
                    take = cnt if room is None else min(cnt, room)
                    _native.call("pb_bloom_add_rows", newest._h, C.c_void_p(rows.data_ptr()), take, None)
                    newest._els_added += take
                    pos += take
                    continue
                if room == 0:
                    # nothing changes until a key turns up that no filter holds; that key grows the stack (:167-168)
                    found = self._found(self._blooms, rows)
                    missing = torch.nonzero(found == 0)
                    if missing.numel() == 0:
                        pos += cnt
                        continue
                    pos += int(missing[0])
                    self._grow()
                    continue
                skip = self._found(self._blooms[:-1], rows) if len(self._blooms) > 1 else None
                novel = torch.empty(cnt, dtype=torch.uint8, device=rows.device)
                _native.call("pb_bloom_novel_rows", newest._h, C.c_void_p(rows.data_ptr()), cnt,
                             C.c_void_p(skip.data_ptr()) if skip is not None else None, C.c_void_p(novel.data_ptr()))
                total = int(novel.sum(dtype=torch.int64))
                if room is None or total <= room:
                    _native.call("pb_bloom_add_rows", newest._h, C.c_void_p(rows.data_ptr()), cnt, C.c_void_p(novel.data_ptr()))
                    newest._els_added += total
                    pos += cnt
                    continue
                # the room-th added row fills the filter: apply up to it, take the rest up again
                csum = torch.cumsum(novel, 0, dtype=torch.int64)
                cut = int(torch.searchsorted(csum, torch.tensor([room], device=csum.device, dtype=torch.int64))[0]) + 1
                _native.call("pb_bloom_add_rows", newest._h, C.c_void_p(rows.data_ptr()), cut, C.c_void_p(novel.data_ptr()))
                newest._els_added += room
                pos += cut
            self._ctx.synchronize()

    # ------------------------------------------------------------------ wire format (:185-247)
    def export(self, file) -> None:
        """:185-207: per filter Q(elements_added) + bit array, then the footer QQQf"""
        if not isinstance(file, (IOBase, mmap)):
            with open(Path(file).expanduser(), "wb") as fp:
                self.export(fp)
            return
        for blm in self._blooms:
            file.write(_U64.pack(blm.elements_added))
            file.write(blm.bloom_numpy().tobytes())
        file.write(_FOOTER.pack(len(self._blooms), self.estimated_elements, self.elements_added, self.false_positive_rate))

    def _load(self, data: bytes) -> None:
        size, est_els, els_added, fpr = self._parse_footer(data)
        self._blooms = []
        self._added_elements = els_added
        self._fpr = fpr
        self._est_elements = est_els
        self._parse_blooms(data, size)

    @classmethod
    def _parse_footer(cls, b):
        size, est_els, els_added, fpr = _FOOTER.unpack(bytes(b[-_FOOTER.size :]))
        return int(size), int(est_els), int(els_added), float(fpr)

    def _parse_blooms(self, b: bytes, size: int) -> None:
        """:229-247"""
        for blm in self._blooms:
            blm.close()
        self._blooms = []
        start = 0
        for _ in range(size):
            blm = self._new_bloom()
            end = start + _U64.size + blm.bloom_length
            blm._upload(bytes(b[start + _U64.size : end]))
            blm._els_added = int(_U64.unpack(bytes(b[start : start + _U64.size]))[0])
            self._blooms.append(blm)
            start = end

    def close(self) -> None:
        for blm in self._blooms:
            blm.close()
        self._blooms = []


class RotatingBloomFilter(ExpandingBloomFilter):
    """expandingbloom.py:250-371: the stack is a queue of at most `max_queue_size` filters; growing a full queue drops
    the oldest filter, and with it the keys only that filter held."""

    def __init__(self, est_elements=None, false_positive_rate=None, max_queue_size: int = 10, filepath=None,
                 hash_function=None, *, device: int = 0, context=None):
        super().__init__(est_elements=est_elements, false_positive_rate=false_positive_rate, filepath=filepath,
                         hash_function=hash_function, device=device, context=context)
        self._queue_size = max_queue_size

    @classmethod
    def frombytes(cls, b, max_queue_size: int, hash_function=None, **kw) -> "RotatingBloomFilter":
        """:289-308"""
        size, est_els, added_els, fpr = cls._parse_footer(b)
        blm = cls(est_elements=est_els, false_positive_rate=fpr, max_queue_size=max_queue_size, hash_function=hash_function, **kw)
        blm._parse_blooms(bytes(b), size)
        blm._added_elements = added_els
        return blm

    @property
    def max_queue_size(self) -> int:
        return self._queue_size

    @property
    def current_queue_size(self) -> int:
        return len(self._blooms)

    def _room(self):
        """:350 rotates when elements_added == estimated_elements, not >=: a filter loaded with more never fills"""
        newest = self._blooms[-1]
        if newest.elements_added > self._est_elements:
            return None
        return self._est_elements - newest.elements_added

    def _grow(self) -> None:
        """:347-361 (the automatic branches).  A full queue drops its oldest filter and appends an empty one of the same
        geometry: the dropped filter's device memory is cleared and becomes the new one (no cudaFree / cudaMalloc per
        rotation -- allocation calls, not GPU time, were what a long batch of rotations cost)."""
        if self.current_queue_size >= self._queue_size:
            blm = self._blooms.pop(0)
            blm.clear()
            if self._blooms:
                _native.call("pb_bloom_move_scratch", self._blooms[-1]._h, blm._h)
            self._blooms.append(blm)
        else:
            self._add_bloom_filter()

    def pop(self) -> None:
        """:332-341"""
        if self.current_queue_size == 1:
            raise RotatingBloomFilterError("Popping a Bloom Filter will result in an unusable system!")
        self._blooms.pop(0).close()

    def push(self) -> None:
        """:343-345"""
        self._grow()
