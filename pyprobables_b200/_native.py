"""ctypes binding of libpb200.so (include/pb200.h) -- the only door from Python into the CUDA engine.

There is deliberately no CPU fallback: if the library or a CUDA device is missing every entry point
raises, so a test or benchmark can never silently run somewhere else.
"""

from __future__ import annotations

import ctypes as C
import sys
import os
import subprocess
import threading
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libpb200.so"
CSRC = _PKG / "csrc"

PB_OK = 0
PB_ERR_BAD_ARG = -1
PB_ERR_CUDA = -2
PB_ERR_OOM = -3
PB_ERR_NO_DEVICE = -4
PB_ERR_UNSUPPORTED = -5
PB_ERR_CUCKOO_FULL = -6


class NativeError(RuntimeError):
    """a libpb200 call failed; .status is the pb_status code"""

    def __init__(self, status: int, message: str):
        super().__init__(f"libpb200: {message} (status {status})")
        self.status = status
        self.message = message


class NoDeviceError(NativeError):
    pass


class pb_keys(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("offsets", C.c_void_p),
        ("n", C.c_uint64),
        ("stride", C.c_uint32),
        ("sym_width", C.c_uint32),
        ("on_device", C.c_int32),
        ("reserved", C.c_int32),
    ]


def build(force: bool = False, verbose: bool = False) -> Path:
    """compile csrc/*.cu for sm_100a into pyprobables_b200/libpb200.so (nvcc cross-compiles without a GPU)"""
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [_PKG.parent / "include" / "pb200.h"]
    newest = max(p.stat().st_mtime for p in srcs)
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        cmd = ["make", "-C", str(CSRC), "-j", str(min(8, os.cpu_count() or 1))]
        if force:
            cmd.append("-B")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            print(r.stdout[-4000:])
            print(r.stderr[-4000:])
        if r.returncode != 0:
            raise RuntimeError("building libpb200.so failed")
    return LIB_PATH


_u64, _u32, _i64, _i32, _vp, _cp = C.c_uint64, C.c_uint32, C.c_int64, C.c_int, C.c_void_p, C.c_char_p
_KP = C.POINTER(pb_keys)
_P = C.POINTER

# name -> argtypes; every function returns int (pb_status) unless listed in _SPECIAL
SIGNATURES: dict[str, list] = {
    "pb_device_count": [_P(C.c_int)],
    "pb_ctx_create": [_i32, _vp, _P(_vp)],
    "pb_ctx_destroy": [_vp],
    "pb_ctx_synchronize": [_vp],
    "pb_ctx_stream": [_vp, _P(_vp)],
    "pb_ctx_launch_count": [_vp, _P(_u64)],
    "pb_ctx_kernel_times": [_vp, _cp, C.c_size_t],
    "pb_ctx_set_option": [_vp, _cp, _i64],
    "pb_ctx_get_option": [_vp, _cp, _P(_i64)],
    "pb_host_alloc": [C.c_size_t, _P(_vp)],
    "pb_host_free": [_vp],
    "pb_dev_alloc": [_vp, C.c_size_t, _P(_vp)],
    "pb_dev_free": [_vp, _vp],
    "pb_memcpy_h2d": [_vp, _vp, _vp, C.c_size_t],
    "pb_memcpy_d2h": [_vp, _vp, _vp, C.c_size_t],
    "pb_memset_dev": [_vp, _vp, _i32, C.c_size_t],
    "pb_flush_l2": [_vp],
    "pb_hash_keys": [_vp, _KP, _u32, _vp, _i32],
    "pb_hash_keys_from": [_vp, _KP, _u64, _i32, _vp, _i32],
    "pb_gen_uniform_keys": [_vp, _u64, _u64, _u64, _vp],
    "pb_gen_rank_keys": [_vp, _vp, _u64, _vp],
    "pb_gen_zipf_ranks": [_vp, _u64, _u64, _u64, C.c_double, _vp],
    "pb_bloom_create": [_vp, _u64, _u32, _P(_vp)],
    "pb_bloom_destroy": [_vp],
    "pb_bloom_clear": [_vp],
    "pb_bloom_upload": [_vp, _vp, _u64],
    "pb_bloom_download": [_vp, _vp, _u64],
    "pb_bloom_device_ptr": [_vp, _P(_vp), _P(_u64)],
    "pb_bloom_add_keys": [_vp, _KP],
    "pb_bloom_check_keys": [_vp, _KP, _vp, _i32],
    "pb_bloom_add_hashes": [_vp, _vp, _u64, _i32],
    "pb_bloom_check_hashes": [_vp, _vp, _u64, _i32, _vp, _i32],
    "pb_bloom_popcount": [_vp, _P(_u64)],
    "pb_bloom_combine": [_vp, _vp, _vp, _i32],
    "pb_bloom_pair_popcounts": [_vp, _vp, _P(_u64)],
    "pb_bloom_create_shard": [_vp, _u64, _u32, _u64, _u64, _P(_vp)],
    "pb_bloom_route_keys": [_vp, _KP, _u64, _u32, _u64, _u32, _vp, _u64, _vp],
    "pb_bloom_partition_layout": [_vp, _u64, _u32, _u64, _u32, _u32, _P(_u32), _P(_u32)],
    "pb_p2p_create": [_vp, _u32, _u32, _u32, _u32, _u32, _P(_vp)],
    "pb_p2p_export": [_vp, _vp],
    "pb_p2p_connect": [_vp, _vp],
    "pb_p2p_connect_local": [_vp, _P(_vp)],
    "pb_p2p_partition_send": [_vp, _KP, _u64, _u32, _u32, _vp, _u64, _vp],
    "pb_p2p_apply": [_vp, _vp, _u32, _u32],
    "pb_p2p_check": [_vp, _P(C.c_int)],
    "pb_p2p_destroy": [_vp],
    "pb_bloom_index_keys": [_vp, _KP, _u64, _u32, _vp],
    "pb_bloom_and_rows": [_vp, _vp, _u64, _u32, _vp],
    "pb_bloom_add_bit_indices": [_vp, _vp, _u64],
    "pb_bloom_test_bit_indices": [_vp, _vp, _u64, _vp],
    "pb_bloom_novel_rows": [_vp, _vp, _u64, _vp, _vp],
    "pb_bloom_add_rows": [_vp, _vp, _u64, _vp],
    "pb_bloom_release_scratch": [_vp],
    "pb_bloom_move_scratch": [_vp, _vp],
    "pb_bloom_rows_in_any": [_vp, _u32, _vp, _u64, _vp],
    "pb_cbloom_create": [_vp, _u64, _u32, _P(_vp)],
    "pb_cbloom_destroy": [_vp],
    "pb_cbloom_clear": [_vp],
    "pb_cbloom_upload": [_vp, _vp, _u64],
    "pb_cbloom_download": [_vp, _vp, _u64],
    "pb_cbloom_device_ptr": [_vp, _P(_vp), _P(_u64)],
    "pb_cbloom_add_keys": [_vp, _KP, _u64],
    "pb_cbloom_add_hashes": [_vp, _vp, _u64, _i32, _u64],
    "pb_cbloom_check_keys": [_vp, _KP, _vp, _i32],
    "pb_cbloom_check_hashes": [_vp, _vp, _u64, _i32, _vp, _i32],
    "pb_cbloom_remove_keys": [_vp, _KP, _u64, _P(_u64)],
    "pb_cbloom_remove_hashes": [_vp, _vp, _u64, _i32, _u64, _P(_u64)],
    "pb_cbloom_stats": [_vp, _P(_u64)],
    "pb_cbloom_combine": [_vp, _vp, _vp, _i32],
    "pb_cbloom_pair_counts": [_vp, _vp, _P(_u64)],
    "pb_cms_create": [_vp, _u32, _u32, _P(_vp)],
    "pb_cms_destroy": [_vp],
    "pb_cms_clear": [_vp],
    "pb_cms_upload": [_vp, _vp, _u64],
    "pb_cms_download": [_vp, _vp, _u64],
    "pb_cms_device_ptr": [_vp, _P(_vp), _P(_u64)],
    "pb_cms_add_keys": [_vp, _KP, _vp, _i64, _P(_i64)],
    "pb_cms_check_keys": [_vp, _KP, _i32, _i64, _vp, _i32],
    "pb_cms_add_hashes": [_vp, _vp, _u64, _i32, _vp, _i64, _P(_i64)],
    "pb_cms_check_hashes": [_vp, _vp, _u64, _i32, _i32, _i64, _vp, _i32],
    "pb_cms_join_buffer": [_vp, _vp, _u64],
    "pb_cms_widen": [_vp, _vp, _u64],
    "pb_cms_load_sums": [_vp, _vp, _u64],
    "pb_cuckoo_create": [_vp, _u64, _u32, _u32, _u32, _u64, _P(_vp)],
    "pb_cuckoo_destroy": [_vp],
    "pb_cuckoo_clear": [_vp],
    "pb_cuckoo_add_keys": [_vp, _KP, _P(_u64), _P(_u64), _vp, _u64],
    "pb_cuckoo_add_fingerprints": [_vp, _vp, _u64, _i32, _P(_u64), _P(_u64), _vp, _u64],
    "pb_cuckoo_check_keys": [_vp, _KP, _vp, _i32],
    "pb_cuckoo_check_fingerprints": [_vp, _vp, _u64, _i32, _vp, _i32],
    "pb_cuckoo_remove_keys": [_vp, _KP, _vp, _i32, _P(_u64)],
    "pb_cuckoo_add_indexed": [_vp, _vp, _vp, _u64, _P(_u64), _P(_u64), _vp, _vp, _u64],
    "pb_cuckoo_check_indexed": [_vp, _vp, _vp, _u64, _vp],
    "pb_cuckoo_remove_indexed": [_vp, _vp, _vp, _u64, _vp, _P(_u64)],
    "pb_cuckoo_set_alt": [_vp, _vp, _u64],
    "pb_cuckoo_resize": [_vp, _u64],
    "pb_cuckoo_fingerprint_info": [_vp, _KP, _vp, _vp, _vp, _i32],
    "pb_cuckoo_count": [_vp, _P(_u64)],
    "pb_cuckoo_download": [_vp, _vp, _u64, _P(C.c_int)],
    "pb_cuckoo_upload": [_vp, _vp, _u64, _i32],
    "pb_cuckoo_device_ptr": [_vp, _P(_vp), _P(_u64)],
    "pb_cuckoo_capacity": [_vp, _P(_u64)],
    "pb_cuckoo_expand": [_vp, _u64, _P(_u64), _vp, _u64],
    "pb_cuckoo_counts_enable": [_vp],
    "pb_cuckoo_counts_add_keys": [_vp, _KP],
    "pb_cuckoo_counts_add_fingerprints": [_vp, _vp, _vp, _u64, _i32],
    "pb_cuckoo_counts_get_keys": [_vp, _KP, _vp, _i32],
    "pb_cuckoo_counts_get_fingerprints": [_vp, _vp, _u64, _vp],
    "pb_cuckoo_counts_set": [_vp, _vp, _vp, _u64],
    "pb_cuckoo_counts_remove_keys": [_vp, _KP, _vp, _i32, _P(_u64), _P(_u64)],
    "pb_cuckoo_counts_remove_fingerprints": [_vp, _vp, _vp, _u64, _vp, _P(_u64), _P(_u64)],
    "pb_microbench_random_atomic": [_vp, _u64, _u64, _i32, _i32, _P(C.c_float)],
}
_SPECIAL = {"pb_version": (C.c_int, []), "pb_last_error": (_cp, [])}
# host-callable arithmetic checks (same code the kernels compile; no GPU needed)
_TEST_HOOKS = {
    "pbt_fnv1a": (_u64, [_vp, _u64, _u64]),
    "pbt_fastmod": (_u64, [_u64, _u64]),
    "pbt_mod_fast33": (_u64, [_u64, _u64]),
    "pbt_cuckoo_info": (None, [_u64, _u32, _u64, _P(_u32), _P(_u64), _P(_u64)]),
    "pbt_sm64": (_u64, [_u64]),
    "pbt_pick_group": (C.c_int, [_u32]),
}

_lib = None
_lock = threading.Lock()


def lib():
    """the loaded library (loads it on first use; raises if it has not been built)"""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not LIB_PATH.exists():
                    raise NativeError(
                        PB_ERR_UNSUPPORTED,
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "or `make -C pyprobables_b200/csrc` (there is no CPU fallback)",
                    )
                L = C.CDLL(str(LIB_PATH))
                for name, args in SIGNATURES.items():
                    f = getattr(L, name)
                    f.restype = C.c_int
                    f.argtypes = args
                for name, (res, args) in {**_SPECIAL, **_TEST_HOOKS}.items():
                    f = getattr(L, name)
                    f.restype = res
                    f.argtypes = args
                _lib = L
    return _lib


def last_error() -> str:
    return (lib().pb_last_error() or b"").decode("utf-8", "replace")


def check(status: int) -> None:
    if status == PB_OK:
        return
    msg = last_error()
    if status == PB_ERR_NO_DEVICE:
        raise NoDeviceError(status, msg)
    if status == PB_ERR_OOM:
        raise MemoryError(f"libpb200: {msg}")
    raise NativeError(status, msg)


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args))


def device_count() -> int:
    n = C.c_int(0)
    call("pb_device_count", C.byref(n))
    return n.value


class Context:
    """one CUDA device + stream + reusable staging buffers (pb_ctx)"""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = _vp()
        call("pb_ctx_create", int(device), _vp(stream) if stream else None, C.byref(h))
        self.handle = h
        self.device = int(device)

    def close(self) -> None:
        if getattr(self, "handle", None):
            lib().pb_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            if not sys.is_finalizing():  # at interpreter exit the CUDA context may already be gone
                self.close()
        except Exception:
            pass

    def synchronize(self) -> None:
        call("pb_ctx_synchronize", self.handle)

    @property
    def stream(self) -> int:
        s = _vp()
        call("pb_ctx_stream", self.handle, C.byref(s))
        return s.value or 0

    @property
    def launch_count(self) -> int:
        n = _u64(0)
        call("pb_ctx_launch_count", self.handle, C.byref(n))
        return n.value

    def set_option(self, name: str, value: int) -> None:
        call("pb_ctx_set_option", self.handle, name.encode(), int(value))

    def get_option(self, name: str) -> int:
        v = _i64(0)
        call("pb_ctx_get_option", self.handle, name.encode(), C.byref(v))
        return v.value

    def kernel_times(self) -> dict:
        """{kernel name: (launches, total device ms)} since the last call (needs set_option('kernel_timing', 1))"""
        buf = C.create_string_buffer(1 << 14)
        call("pb_ctx_kernel_times", self.handle, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.rsplit(" ", 2)
            out[name] = (int(n), float(ms))
        return out

    def flush_l2(self) -> None:
        call("pb_flush_l2", self.handle)

    # raw device buffers for ctypes-only callers (bench, tests)
    def dev_alloc(self, nbytes: int) -> int:
        p = _vp()
        call("pb_dev_alloc", self.handle, int(nbytes), C.byref(p))
        return p.value

    def dev_free(self, ptr: int) -> None:
        call("pb_dev_free", self.handle, _vp(ptr))

    def h2d(self, dst: int, src: int, nbytes: int) -> None:
        call("pb_memcpy_h2d", self.handle, _vp(dst), _vp(src), int(nbytes))

    def d2h(self, dst: int, src: int, nbytes: int) -> None:
        call("pb_memcpy_d2h", self.handle, _vp(dst), _vp(src), int(nbytes))

    def gen_uniform_keys(self, first: int, n: int, out_dev: int, seed: int = 0xB200) -> None:
        call("pb_gen_uniform_keys", self.handle, seed, int(first), int(n), _vp(out_dev))

    def gen_rank_keys(self, ranks_dev: int, n: int, out_dev: int) -> None:
        call("pb_gen_rank_keys", self.handle, _vp(ranks_dev), int(n), _vp(out_dev))

    def gen_zipf_ranks(self, first: int, n: int, out_dev: int, a: float = 1.1, seed: int = 0xB200) -> None:
        call("pb_gen_zipf_ranks", self.handle, seed, int(first), int(n), float(a), _vp(out_dev))

    def microbench(self, words: int, n: int, op: int, reps: int = 3) -> float:
        """device milliseconds for n random atomics (op 0 RED.OR, 1 RED.ADD, 2 gather) or a copy (op 3)"""
        ms = C.c_float(0)
        call("pb_microbench_random_atomic", self.handle, int(words), int(n), int(op), int(reps), C.byref(ms))
        return ms.value


_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx
