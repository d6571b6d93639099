"""Hash plugin surface (mirror of probables/hashes.py).

`default_fnv_1a` / `fnv_1a` are the *identities* the filters dispatch on: a structure built with one of
them (or with hash_function=None) runs the fused hash+scatter CUDA kernels.  Called directly they hash
on the GPU as well (pb_hash_keys), so there is exactly one implementation of the arithmetic in the
product -- the device function in csrc/pb_hash.cuh.  Any other callable is a host plugin: the filter
calls it per key and ships the integers to the pre-hashed kernels (add_alt/check_alt semantics).
"""

from __future__ import annotations

import ctypes as C
import hashlib
import struct
from functools import wraps
from typing import Callable, Union

import numpy as np

from . import _native
from .keys import pack_keys

KeyT = Union[str, bytes]
HashResultsT = list
HashFuncT = Callable[..., list]
SimpleHashT = Callable[..., int]


def hash_many(keys, depth: int = 1, device: int = 0) -> np.ndarray:
    """default_fnv_1a for a whole batch on the GPU -> uint64[n, depth] (hashes.py:71-103)"""
    kb = pack_keys(keys)
    out = np.empty((kb.n, depth), dtype=np.uint64)
    if kb.n:
        ctx = _native.default_context(device)
        _native.call("pb_hash_keys", ctx.handle, kb.ref(), int(depth), C.c_void_p(out.ctypes.data), 0)
    return out


def default_fnv_1a(key: KeyT, depth: int = 1) -> list:
    """[fnv_1a(key, seed) for seed in range(depth)] (hashes.py:71-83)"""
    return [int(x) for x in hash_many([key], depth)[0]]


def _hash_from(key: KeyT, start: int, bits: int, device: int = 0) -> int:
    kb = pack_keys([key])
    out = np.empty(1, dtype=np.uint64)
    ctx = _native.default_context(device)
    _native.call("pb_hash_keys_from", ctx.handle, kb.ref(), start & 0xFFFFFFFFFFFFFFFF, bits, C.c_void_p(out.ctypes.data), 0)
    return int(out[0])


def fnv_1a(key: KeyT, seed: int = 0) -> int:
    """64-bit FNV-1a started from basis + 31*seed, any integer seed (hashes.py:86-103)"""
    return _hash_from(key, (14695981039346656037 + 31 * int(seed)) & 0xFFFFFFFFFFFFFFFF, 64)


def fnv_1a_32(key: KeyT, seed: int = 0) -> int:
    """32-bit FNV-1a started from 0x811C9DC5 + 31*seed (hashes.py:106-122)"""
    return _hash_from(key, (0x811C9DC5 + 31 * int(seed)) & 0xFFFFFFFF, 32)


def hash_with_depth_bytes(func):
    """decorator: `func(key_bytes, idx) -> digest bytes` becomes a (key, depth) -> [u64] hash function.
    The digest of round i is the input of round i+1 and the first 8 digest bytes (native order) are the
    hash, as hashes.py:18-41 defines it; str keys are utf-8 encoded first (hashes.py:35)."""

    @wraps(func)
    def hashing_func(key, depth=1):
        cur = key.encode("utf-8") if isinstance(key, str) else key
        out = []
        for i in range(depth):
            cur = func(cur, i)
            out.append(struct.unpack("Q", cur[:8])[0])
        return out

    return hashing_func


def hash_with_depth_int(func):
    """decorator: `func(key, idx) -> int` becomes a (key, depth) -> [int] hash function; round i+1 hashes
    the lower-case hex text of round i's result (hashes.py:44-68)"""

    @wraps(func)
    def hashing_func(key, depth=1):
        cur = func(key, 0)
        out = [cur]
        for i in range(1, depth):
            cur = func(format(cur, "x"), i)
            out.append(cur)
        return out

    return hashing_func


@hash_with_depth_bytes
def default_md5(key, *args, **kwargs) -> bytes:
    """md5 digest chain (hashes.py:125-137); runs in hashlib on the host: a host plugin"""
    return hashlib.md5(key).digest()


@hash_with_depth_bytes
def default_sha256(key, *args, **kwargs) -> bytes:
    """sha256 digest chain (hashes.py:139-150); host plugin"""
    return hashlib.sha256(key).digest()


def is_default_hash(func) -> bool:
    """True when `func` selects the fused on-device FNV-1a path"""
    return func is None or func is default_fnv_1a or func is fnv_1a
