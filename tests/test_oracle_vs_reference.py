"""Live cross-check of the CPU oracle against the real pure-Python reference on random inputs.

Runs only where the reference checkout exists (this container: /root/reference); on the GPU box it is skipped --
there the committed fixtures of tests/golden/ (generated from the same reference) do the pinning.  The product
never touches either; this is the oracle's own safety net."""

import random
import struct
import sys
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF / "probables" / "__init__.py").exists(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, str(REF))
    try:
        import probables

        yield probables
    finally:
        sys.path.remove(str(REF))


def _random_keys(rng, n, max_len=40, unicode_share=0.2):
    keys = []
    for _ in range(n):
        if rng.random() < unicode_share:
            keys.append("".join(chr(rng.choice([rng.randrange(32, 127), rng.randrange(160, 256), rng.randrange(0x400, 0x500),
                                                rng.randrange(0x1F600, 0x1F640)])) for _ in range(rng.randrange(0, 12))))
        else:
            keys.append(bytes(rng.randrange(256) for _ in range(rng.randrange(0, max_len))))
    return keys


def test_fnv_random(ref, orc):
    from probables.hashes import default_fnv_1a

    rng = random.Random(1)
    for key in _random_keys(rng, 400):
        depth = rng.randrange(1, 12)
        assert orc.default_fnv_1a(key, depth) == default_fnv_1a(key, depth), key


@pytest.mark.parametrize("est,fpr", [(50, 0.2), (1000, 0.01), (5000, 0.0005)])
def test_bloom_random(ref, orc, est, fpr):
    rng = random.Random(est)
    keys = _random_keys(rng, est)
    probes = keys[: est // 3] + _random_keys(rng, est)
    r = ref.BloomFilter(est_elements=est, false_positive_rate=fpr)
    for k in keys:
        r.add(k)
    f32, k, m, length = orc.bloom_params(est, fpr)
    assert (f32, k, m, length) == (r.false_positive_rate, r.number_hashes, r.number_bits, r.bloom_length)
    o = orc.Bloom(m, k)
    o.add(orc.pack(keys))
    assert bytes(o.bloom) == bytes(r.bloom)
    assert o.check(orc.pack(probes)).tolist() == [r.check(p) for p in probes]
    assert o.popcount() == r._cnt_number_bits_set()
    # export bytes incl. footer
    assert bytes(o.bloom) + struct.pack("QQf", est, len(keys), f32) == bytes(r)


@pytest.mark.parametrize("width,depth,query", [(64, 4, "min"), (101, 5, "mean"), (37, 6, "mean-min"), (1000, 7, "mean-min")])
def test_cms_random(ref, orc, width, depth, query):
    rng = random.Random(width * depth)
    cls = {"min": ref.CountMinSketch, "mean": ref.CountMeanSketch, "mean-min": ref.CountMeanMinSketch}[query]
    r = cls(width=width, depth=depth)
    o = orc.CMS(width, depth, query)
    pool = _random_keys(rng, 60, unicode_share=0.1)
    seq = [(rng.choice(pool), rng.choice([1, 1, 1, 2, 7, 1000, 2**20])) for _ in range(1500)]
    want = [r.add(k, n) for k, n in seq]
    got = o.add(orc.pack([k for k, _ in seq]), np.array([n for _, n in seq], dtype=np.int64), want_returns=True)
    assert got.tolist() == want
    assert o.bins.tolist() == list(r._bins) and o.elements_added == r.elements_added
    assert o.check(orc.pack(pool)).tolist() == [r.check(k) for k in pool]
    # saturation at INT32_MAX
    r.add(pool[0], 2**31)
    o.add(orc.pack([pool[0]]), 2**31)
    assert o.bins.tolist() == list(r._bins) and o.check(orc.pack(pool[:5])).tolist() == [r.check(k) for k in pool[:5]]


@pytest.mark.parametrize("capacity,bucket,fp_bytes", [(200, 4, 4), (64, 2, 2), (500, 8, 1), (97, 3, 3)])
def test_cuckoo_random(ref, orc, capacity, bucket, fp_bytes):
    rng = random.Random(capacity)
    keys = _random_keys(rng, int(capacity * bucket * 0.5), unicode_share=0.1)
    r = ref.CuckooFilter(capacity=capacity, bucket_size=bucket, max_swaps=100, finger_size=fp_bytes, auto_expand=False)
    o = orc.Cuckoo(capacity, bucket, 100, fp_bytes * 8)
    i1, i2, fp = o.fingerprint_info(orc.pack(keys))
    assert [(int(a), int(b), int(c)) for a, b, c in zip(i1, i2, fp)] == [r._generate_fingerprint_info(k) for k in keys]
    random.seed(5)
    for k in keys:
        r.add(k)
    assert len(o.add(orc.pack(keys))) == 0
    assert o.elements_added == r.elements_added
    stored = sorted(x for b in r.buckets for x in b)
    assert o.fingerprints().tolist() == stored
    probes = keys[:50] + _random_keys(rng, 300)
    assert o.check(orc.pack(probes)).tolist() == [r.check(p) for p in probes]


@pytest.mark.parametrize("est,fpr", [(20, 0.05), (300, 0.01), (2000, 0.2)])
def test_counting_bloom_random(ref, orc, est, fpr):
    """oracle head start for the next scope row (CountingBloomFilter, countingbloom.py:125-208), including the
    double increment when two of a key's hashes share an index"""
    rng = random.Random(est)
    r = ref.CountingBloomFilter(est_elements=est, false_positive_rate=fpr)
    o = orc.CountingBloom(r.number_bits, r.number_hashes)
    pool = _random_keys(rng, max(est // 2, 10), unicode_share=0.1)
    for _ in range(6):
        batch = [rng.choice(pool) for _ in range(est)]
        n = rng.choice([1, 1, 2, 5])
        assert o.add(orc.pack(batch), n).tolist() == [r.add(k, n) for k in batch]
        assert o.bloom.tolist() == list(r._bloom) and o.elements_added == r.elements_added
        rem = [rng.choice(pool) for _ in range(est // 3)]
        m = rng.choice([1, 3])
        assert o.remove(orc.pack(rem), m).tolist() == [r.remove(k, m) for k in rem]
        assert o.bloom.tolist() == list(r._bloom) and o.elements_added == r.elements_added
    probes = pool + _random_keys(rng, 100)
    assert o.check(orc.pack(probes)).tolist() == [r.check(p) for p in probes]
    # export bytes: uint32 counters + the Bloom footer
    assert o.bloom.tobytes() + struct.pack("QQf", est, o.elements_added, r.false_positive_rate) == bytes(r)
