"""CPU-side checks of the product: the C-ABI library loads and exports every symbol include/pb200.h
declares, the integer arithmetic the kernels compile (host-callable pbt_* hooks of the same header code)
matches Python / the golden vectors, key packing follows the reference's symbol rules, and the engine
fails loudly without a GPU.  No compute call touches a device here."""

import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def native():
    from pyprobables_b200 import _native

    _native.build()
    return _native


def test_library_exports_every_declared_symbol(native):
    header = (ROOT / "include" / "pb200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 55
    lib = C.CDLL(str(native.LIB_PATH))
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, f"declared in pb200.h but not exported: {missing}"
    # and the Python binding table covers the header exactly
    bound = set(native.SIGNATURES) | set(native._SPECIAL)
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    assert native.lib().pb_version() == 100


def test_no_gpu_means_loud_failure(native):
    import pyprobables_b200 as pb

    if pb.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    for make in (lambda: pb.BloomFilter(10, 0.05), lambda: pb.CountMinSketch(width=10, depth=2), lambda: pb.CuckooFilter(),
                 lambda: pb.hashes.default_fnv_1a("x", 2)):
        with pytest.raises(pb.NoDeviceError, match="no CPU fallback"):
            make()


def test_product_never_imports_the_oracle():
    for p in (ROOT / "pyprobables_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h") and p.is_file():
            txt = p.read_text()
            for needle in ("import oracle", "from oracle", "liboracle", "pb_oracle", "orc_"):
                assert needle not in txt, f"{p} references the test oracle ({needle})"


def test_fnv_hook_matches_golden(native, golden):
    L = native.lib()

    def fnv(b: bytes, seed: int) -> int:
        buf = (C.c_uint8 * max(len(b), 1)).from_buffer_copy(b or b"\0")
        return L.pbt_fnv1a(buf, len(b), seed)

    exp = [4040040117721899264, 3916497180155386777, 468410530588793106, 13781401791305604595, 321382271269641900]
    assert [fnv(b"this is a test", s) for s in range(5)] == exp  # hashes_test.py:27-46
    for hx, want in golden["fnv_bytes"].items():
        assert [fnv(bytes.fromhex(hx), s) for s in range(4)] == want
    for seed, want in golden["fnv_seed_big"].items():
        assert fnv(b"seed test", int(seed)) == want
    for i, hx in golden["keys"].items():
        g = int(i)
        lo, hi = L.pbt_sm64(0xB200 + 2 * g), L.pbt_sm64(0xB200 + 2 * g + 1)
        assert (lo.to_bytes(8, "little") + hi.to_bytes(8, "little")).hex() == hx


def test_fastmod_exact(native, golden):
    L = native.lib()
    rng = np.random.default_rng(7)
    moduli = [1, 2, 3, 7, 63, 64, 97, 1000, 2**20, 2**20 + 7, 9585059, 2**32 - 1, 2**32, 2**32 + 1, 9585058424, 143775874672,
              2**63 - 1, 2**63, 2**63 + 1, 2**64 - 1]
    moduli += [int(x) for x in rng.integers(1, 2**63, size=40, dtype=np.uint64)]
    for m in moduli:
        hs = [0, 1, m - 1, m, m + 1, 2 * m - 1, 2 * m, 2**64 - 1, 2**64 - 2, 2**63, (2**64 // m) * m - 1, (2**64 // m) * m]
        hs += [int(x) for x in rng.integers(0, 2**64, size=200, dtype=np.uint64)]
        for h in hs:
            h %= 2**64
            assert L.pbt_fastmod(h, m) == h % m, (h, m)
    for name, kat in golden["bloom_index_kat"].items():
        for hashes, bits in zip(golden.get("fnv_key0_k7") and [golden["fnv_key0_k7"]] or [], kat["bits"][:1]):
            if kat["k"] == 7:
                assert [L.pbt_fastmod(h, kat["m"]) for h in hashes] == bits


def test_mod_fast33_exact(native):
    """host twin of the partition kernels' modulo for m >= 2^33 (signed mad.wide form) against Python's %"""
    L = native.lib()
    rng = np.random.default_rng(33)
    moduli = [2**33, 2**33 + 1, 9585058424, 143775874672, 2**34 - 1, 2**40 + 12345, 2**63 - 1, 2**63, 2**63 + 1, 2**64 - 1]
    moduli += [int(x) for x in rng.integers(2**33, 2**64 - 1, size=50, dtype=np.uint64)]
    for m in moduli:
        hs = [0, 1, m - 1, m, m + 1, 2 * m - 1, 2 * m, 2**64 - 1, 2**64 - 2, 2**63, (2**64 // m) * m - 1, (2**64 // m) * m]
        hs += [int(x) for x in rng.integers(0, 2**64 - 1, size=500, dtype=np.uint64)]
        for h in hs:
            h %= 2**64
            assert L.pbt_mod_fast33(h, m) == h % m, (h, m)


def test_cuckoo_index_hook(native, golden, orc):
    L = native.lib()
    keys = orc.uniform_keys(0, 8)
    h0 = orc.default_fnv_1a_many(orc.pack(keys), 1)[:, 0]
    for cap, exp in golden["cuckoo_info"].items():
        got = []
        for h in h0:
            fp, i1, i2 = C.c_uint32(), C.c_uint64(), C.c_uint64()
            L.pbt_cuckoo_info(int(h), 32, int(cap), C.byref(fp), C.byref(i1), C.byref(i2))
            got.append([i1.value, i2.value, fp.value])
        assert got == exp
    # test_utilities.py:39-83 style: the low-bits mask for narrow fingerprints
    fp, i1, i2 = C.c_uint32(), C.c_uint64(), C.c_uint64()
    for bits in (1, 7, 8, 20, 31, 32):
        L.pbt_cuckoo_info(0xFFFFFFFFFFFFFFFF, bits, 1000, C.byref(fp), C.byref(i1), C.byref(i2))
        assert fp.value == (1 << bits) - 1
        assert i1.value == fp.value % 1000
    # decimal-string hash of every digit count
    import sys

    sys.path.insert(0, str(ROOT))
    for v in (0, 5, 10, 99, 100, 4294967295, 1000000000, 999999999, 2887966554):
        L.pbt_cuckoo_info(v, 32, 2**28, C.byref(fp), C.byref(i1), C.byref(i2))
        assert i2.value == orc.fnv_1a(str(v)) % 2**28


def test_pick_group(native):
    L = native.lib()
    for k in range(1, 40):
        g = L.pbt_pick_group(k)
        assert 1 <= g <= 8 and (k <= 8 and g == k or k > 8)


def test_key_packing_rules():
    from pyprobables_b200.keys import pack_keys

    kb = pack_keys([b"abcd", "wxyz"])
    assert (kb.n, kb.c.stride, kb.c.sym_width, bool(kb.c.offsets)) == (2, 4, 1, False)
    assert bytes(kb._keep[0]) == b"abcdwxyz"
    kb = pack_keys(["a", b"", "ccc"])
    assert kb.c.sym_width == 1 and kb._keep[1].tolist() == [0, 1, 1, 4]
    kb = pack_keys(["é"])  # latin-1 range: one byte 0xE9, not its two utf-8 bytes (hashes.py:98)
    assert kb.c.sym_width == 1 and bytes(kb._keep[0]) == b"\xe9"
    kb = pack_keys(["日本", b"ab"])  # code points > 255: whole batch as u32 symbols
    assert kb.c.sym_width == 4 and kb._keep[0].tolist() == [0x65E5, 0x672C, 0x61, 0x62] and kb._keep[1].tolist() == [0, 2, 4]
    arr = np.arange(32, dtype=np.uint8).reshape(2, 16)
    kb = pack_keys(arr)
    assert (kb.n, kb.c.stride, kb.on_device) == (2, 16, False)
    kb = pack_keys((np.frombuffer(b"abcdef", dtype=np.uint8), np.array([0, 2, 2, 6], dtype=np.uint64)))
    assert kb.n == 3 and kb.c.stride == 0
    assert pack_keys("single").n == 1 and pack_keys([]).n == 0
    with pytest.raises(TypeError):
        pack_keys([1, 2, 3])
    with pytest.raises(ValueError):
        pack_keys((np.zeros(3, dtype=np.uint8), np.array([0, 5], dtype=np.uint64)))


def test_shim_validation_matches_reference_messages(native):
    """constructor validation happens before any device work, so the messages can be checked on CPU
    (bloom_test.py:395-473, countminsketch_test.py:435-561, cuckoo_test.py:353-447)"""
    import pyprobables_b200 as pb
    from pyprobables_b200.bloom import optimized_params

    assert optimized_params(10, 0.05) == (0.05000000074505806, 4, 63)
    assert optimized_params(16_000_000, 0.001)[1:] == (10, 230041400)
    with pytest.raises(pb.InitializationError, match="estimated elements must be greater than 0"):
        optimized_params(0, 0.1)
    with pytest.raises(pb.InitializationError, match="false positive rate must be between 0.0 and 1.0"):
        optimized_params(10, 1.5)
    with pytest.raises(pb.InitializationError, match="Number hashes is zero"):
        optimized_params(100, 0.999)
    if pb.device_count() == 0:
        with pytest.raises(pb.InitializationError, match="capacity, bucket_size, and max_swaps"):
            pb.CuckooFilter(capacity=0)
    e = pb.CuckooFilterFullError("The CuckooFilter is currently full")
    assert e.message == "The CuckooFilter is currently full" and isinstance(e, pb.ProbablesBaseException)
    # host hash plugins (hashes_test.py:64-146 vectors)
    assert pb.hashes.default_md5("this is a test", 3)[0] == 12174049463882854484
    assert pb.hashes.default_sha256("this is a test", 1)[0] == 10244166640140130606


def test_constructor_errors_come_before_device_work():
    """argument validation needs no GPU: the reference's messages (bloom_test.py:395-473,
    countminsketch_test.py:435-561, cuckoo_test.py:353-447) are raised even on a GPU-less machine"""
    import pyprobables_b200 as pb

    cases = [
        (lambda: pb.BloomFilter(est_elements=100, false_positive_rate=1.1), pb.InitializationError,
         "Bloom: false positive rate must be between 0.0 and 1.0"),
        (lambda: pb.BloomFilter(est_elements=0, false_positive_rate=0.1), pb.InitializationError,
         "Bloom: estimated elements must be greater than 0"),
        (lambda: pb.BloomFilter(est_elements=100, false_positive_rate="1.1"), pb.InitializationError,
         "Bloom: false positive rate must be between 0.0 and 1.0"),
        (lambda: pb.BloomFilter(est_elements=[0], false_positive_rate=0.1), pb.InitializationError,
         "Bloom: estimated elements must be greater than 0"),
        (lambda: pb.BloomFilter(est_elements=10, false_positive_rate=0.999), pb.InitializationError,
         "Bloom: Number hashes is zero; unusable parameters provided"),
        (lambda: pb.BloomFilter(), pb.InitializationError, "Insufecient parameters to set up the Bloom Filter"),
        (lambda: pb.CountMinSketch(width=0, depth=5), pb.InitializationError, "CountMinSketch: width and depth must be greater than 0"),
        (lambda: pb.CountMinSketch(width=10, depth=-1), pb.InitializationError, "CountMinSketch: width and depth must be greater than 0"),
        (lambda: pb.CountMinSketch(confidence=-1, error_rate=0.1), pb.InitializationError,
         "CountMinSketch: width and depth must be greater than 0"),
        (lambda: pb.CountMinSketch(), pb.InitializationError, "Must provide one of the following to initialize the Count-Min Sketch:"),
        (lambda: pb.CuckooFilter(capacity=0), pb.InitializationError,
         "CuckooFilter: capacity, bucket_size, and max_swaps must be an integer greater than 0"),
        (lambda: pb.CuckooFilter(bucket_size=0), pb.InitializationError,
         "CuckooFilter: capacity, bucket_size, and max_swaps must be an integer greater than 0"),
        (lambda: pb.CuckooFilter(max_swaps="a"), pb.InitializationError,
         "CuckooFilter: capacity, bucket_size, and max_swaps must be an integer greater than 0"),
        (lambda: pb.CuckooFilter(finger_size=5), ValueError, "CuckooFilter: fingerprint size must be between 1 and 4"),
        (lambda: pb.CuckooFilter(filepath="/nonexistent/file.cko"), pb.InitializationError, "CuckooFilter: failed to load provided file"),
    ]
    for make, exc, msg in cases:
        with pytest.raises(exc) as ei:
            make()
        assert str(ei.value).startswith(msg), (str(ei.value), msg)


def test_shard_plan_is_window_aligned():
    from pyprobables_b200.sharded import ShardPlan

    for m, g in ((9585058424, 1), (19170116848, 2), (76680467392, 8), (143775874672, 8), (479252922, 2), (63, 4)):
        p = ShardPlan.make(m, g)
        assert p.shard_bits == p.windows_per_rank << p.window_log2
        assert p.total_windows == p.windows_per_rank * g <= ShardPlan.MAX_WINDOWS
        assert (p.total_windows << p.window_log2) >= m
        assert sum(p.active_windows(r) for r in range(g)) == -(-m // (1 << p.window_log2))
        for r in range(g):
            lo, hi = p.bounds(r)
            assert lo % (1 << p.window_log2) == 0 or lo == m


def test_slice_batch_views_the_same_keys():
    """keys.slice_batch (block-wise hashing in the stateful wrappers): a slice of a packed batch is the same keys, for
    the fixed-stride and the offsets layout, u8 and u32 symbols"""
    import ctypes as C

    from pyprobables_b200.keys import pack_keys, slice_batch

    def unpack(kb):
        sw = int(kb.c.sym_width)
        out = []
        if kb.c.offsets:
            o = np.frombuffer((C.c_uint64 * (kb.n + 1)).from_address(kb.c.offsets), dtype=np.uint64)
            for a, b in zip(o[:-1], o[1:]):
                out.append(bytes((C.c_uint8 * (int(b - a) * sw)).from_address(kb.c.data + int(a) * sw)) if b > a else b"")
        else:
            step = int(kb.c.stride) * sw
            for i in range(kb.n):
                out.append(bytes((C.c_uint8 * step).from_address(kb.c.data + i * step)))
        return out

    ragged = [f"key-{i}-{'z' * (i % 7)}" for i in range(100)]
    fixed = np.arange(100 * 16, dtype=np.uint8).reshape(100, 16)
    wide = [f"clé-{i}-中" for i in range(50)]
    for batch in (ragged, fixed, wide):
        kb = pack_keys(batch)
        whole = unpack(kb)
        for lo, hi in ((0, 100), (10, 35), (99, 100), (40, 40), (60, 1000)):
            part = slice_batch(kb, lo, hi)
            assert part.n == max(0, min(hi, kb.n) - min(lo, kb.n)) and unpack(part) == whole[lo:hi]
            assert part.on_device == kb.on_device and int(part.c.sym_width) == int(kb.c.sym_width)
