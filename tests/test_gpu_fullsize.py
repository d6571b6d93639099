"""BASELINE.json-size checks on one B200 (gpu-marked; ~1 minute): the full-size structures (1.2 GB Bloom bit
array with its > 2^32 modulus, the 2^20 x 5 Count-Min table, a 2^26 x 4 Cuckoo table) exercised through
size-independent properties plus an oracle comparison on a sample that the host finishes in seconds.
The complete 1e9-key runs with bit-exact oracle parity are benchmarks/full_configs.py (results in profiles/)."""

import ctypes as C
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def md5(a):
    return hashlib.md5(memoryview(np.ascontiguousarray(a))).hexdigest()


@pytest.fixture(scope="module")
def env():
    import torch

    import pyprobables_b200 as pb

    assert pb.device_count() >= 1
    ctx = pb.default_context()
    return torch, pb, ctx


def _device_keys(torch, ctx, first, n):
    t = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
    ctx.gen_uniform_keys(first, n, t.data_ptr())
    ctx.synchronize()
    return t


def test_bloom_config2_size(env, orc):
    torch, pb, ctx = env
    n = 60_000_000
    keys = _device_keys(torch, ctx, 0, n)
    whole = pb.BloomFilter(10**9, 0.01)
    assert (whole.number_bits, whole.number_hashes, whole.bloom_length) == (9585058424, 7, 1198132303)
    whole.add_many(keys)  # partitioned path (bitmap >> L2, big batch), overlapped passes
    ctx.synchronize()
    bits = whole._cnt_number_bits_set()
    # oracle on the same keys, whole 1.2 GB bit array
    ob = orc.Bloom(whole.number_bits, whole.number_hashes)
    for lo in range(0, n, 20_000_000):
        ob.add(orc.pack(orc.uniform_keys(lo, 20_000_000)))
    assert bits == ob.popcount()
    assert md5(whole.bloom_numpy()) == md5(ob.bloom)
    # idempotence: inserting everything again changes nothing
    whole.add_many(keys)
    ctx.synchronize()
    assert whole._cnt_number_bits_set() == bits
    # linearity: two halves through the DIRECT kernel, united on the device, equal the whole
    ctx.set_option("bloom_insert_mode", 1)
    try:
        a, b = pb.BloomFilter(10**9, 0.01), pb.BloomFilter(10**9, 0.01)
        a.add_many(keys[: n // 2])
        b.add_many(keys[n // 2 :])
        ctx.synchronize()
    finally:
        ctx.set_option("bloom_insert_mode", 0)
    u = a.union(b)
    counts = (C.c_uint64 * 2)()
    from pyprobables_b200 import _native

    _native.call("pb_bloom_pair_popcounts", u._h, whole._h, counts)
    assert counts[0] == counts[1] == bits  # |u OR whole| == |u AND whole| == |whole|  <=>  identical bit arrays
    # membership: everything inserted is present (30 M device keys: the auto mode samples the hit rate and takes the
    # partitioned query); absent probes agree with the oracle and with the direct kernel
    assert bool(whole.check_many(keys[:30_000_000]).all())
    mixed = torch.cat([keys[:10_000_000], _device_keys(torch, ctx, 10**9, 10_000_000), keys[-5:]])
    ctx.set_option("bloom_check_mode", 2)
    try:
        part = whole.check_many(mixed)
    finally:
        ctx.set_option("bloom_check_mode", 1)
    try:
        direct = whole.check_many(mixed)
    finally:
        ctx.set_option("bloom_check_mode", 0)
    assert bool((part == direct).all()) and bool(part[:10_000_000].all()) and bool(part[-5:].all())
    assert int(part[10_000_000:-5].sum()) < 100  # the filter is 4 % full: essentially no false positives
    probe = orc.uniform_keys(10**9, 2_000_000)
    assert (whole.check_many(probe) == ob.check(orc.pack(probe))).all()
    for f in (whole, a, b, u):
        f.close()


def test_cms_config3_size(env, orc):
    torch, pb, ctx = env
    n = 30_000_000
    ranks = np.random.default_rng(3).zipf(1.1, n).astype(np.uint64)
    keys = orc.rank_keys(ranks)
    whole = pb.CountMinSketch(width=1 << 20, depth=5)
    whole.add_many(keys)
    oc = orc.CMS(1 << 20, 5)
    oc.add_parallel(orc.pack(keys))
    assert md5(whole.bins_numpy()) == md5(oc.bins) and whole.elements_added == n
    # linearity: sketches of two halves joined == sketch of the whole; sum of row 0 == number of adds
    a, b = pb.CountMinSketch(width=1 << 20, depth=5), pb.CountMinSketch(width=1 << 20, depth=5)
    a.add_many(keys[: n // 3])
    b.add_many(keys[n // 3 :])
    a.join(b)
    assert md5(a.bins_numpy()) == md5(oc.bins) and a.elements_added == n
    assert int(whole.bins_numpy().reshape(5, -1).sum(axis=1).min()) == n == int(whole.bins_numpy().reshape(5, -1).sum(axis=1).max())
    top = orc.rank_keys(np.arange(1, 100_001, dtype=np.uint64))
    est = whole.check_many(top)
    assert (est == oc.check(orc.pack(top))).all()
    assert (np.diff(est[:50]) <= 0).all()  # the heaviest ranks come out in order


def test_cuckoo_large_table(env, orc):
    torch, pb, ctx = env
    cap = 1 << 26
    f = pb.CuckooFilter(capacity=cap, bucket_size=4, max_swaps=500, auto_expand=False)
    n = int(cap * 4 * 0.90)
    keys = _device_keys(torch, ctx, 0, n)
    f.add_many(keys)
    assert abs(f.load_factor() - 0.90) < 0.03 and f.elements_added <= n
    present = f.check_many(keys[:20_000_000])
    assert bool(present.all())
    # re-adding is a no-op (dedupe, cuckoo.py:300-302)
    before = f.elements_added
    f.add_many(keys[:10_000_000])
    assert f.elements_added == before
    # elements_added equals the number of distinct fingerprints (oracle fingerprint function, torch.unique)
    ofp = orc.Cuckoo(16, 4, 5, 32)
    pieces = []
    for lo in range(0, n, 40_000_000):
        m = min(40_000_000, n - lo)
        pieces.append(torch.unique(torch.from_numpy(ofp.fingerprint_info(orc.pack(orc.uniform_keys(lo, m)))[2].astype(np.int64)).cuda()))
    uniq = torch.unique(torch.cat(pieces))
    assert int(uniq.numel()) == f.elements_added
    f.close()
