"""CountingBloomFilter on the device (reference: probables/blooms/countingbloom.py:125-208) against
the reference's own test literal, tests/golden/golden_r2.json (recorded from the pure-Python reference by
make_golden_r2.py) and the CPU oracle.  Counters are integers: bit-exact everywhere."""

import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()


@pytest.fixture(scope="module")
def pb():
    import pyprobables_b200 as p

    assert p.device_count() >= 1
    return p


def test_reference_literals(pb):
    # tests/countingbloom_test.py:106-144
    c = pb.CountingBloomFilter(est_elements=10, false_positive_rate=0.01)
    for w in ("test", "out", "the", "counting", "bloom", "filter", "test", "Test", "out", "test"):
        c.add(w)
    assert md5(bytes(c)) == "0b83c837da30e25f768f0527c039d341"
    assert c.check("test") == 3 and c.check("out") == 2 and c.check("bloom") == 1 and c.check("nope") == 0
    assert c.elements_added == 10
    c2 = pb.CountingBloomFilter.frombytes(bytes(c))
    assert bytes(c2) == bytes(c) and c2.check("test") == 3 and c2.elements_added == 10
    c3 = pb.CountingBloomFilter(hex_string=c.export_hex())
    assert c3.export_hex() == c.export_hex() and c3.check("out") == 2
    with pytest.raises(pb.InitializationError):
        pb.CountingBloomFilter()


def test_sequence_with_colliding_hashes(pb, golden):
    """399 single add / remove / check calls: every return value, the counters (incl. the double increment when two
    hashes of a key share a counter), elements_added, the export bytes and the statistics string"""
    g = golden["cbloom_seq"]
    c = pb.CountingBloomFilter(est_elements=g["est"], false_positive_rate=g["fpr"])
    assert (c.number_bits, c.number_hashes) == (g["num_bits"], g["k"])
    got = []
    for op, key, n in g["ops"]:
        got.append(c.add(key, n) if op == "add" else c.remove(key, n) if op == "remove" else c.check(key))
    assert got == g["returns"]
    assert c.bloom_numpy().tolist() == g["counters"] and c.elements_added == g["elements_added"]
    assert md5(bytes(c)) == g["export_md5"] and c.export_hex() == g["export_hex"]
    assert str(c) == g["str"]
    # the same adds as batches (grouped by num_els) leave the same counters as the sequential loop
    adds = [(k, n) for op, k, n in g["ops"] if op == "add"]
    c2 = pb.CountingBloomFilter(est_elements=g["est"], false_positive_rate=g["fpr"])
    c3 = pb.CountingBloomFilter(est_elements=g["est"], false_positive_rate=g["fpr"])
    for k, n in adds:
        c2.add(k, n)
    for n in sorted({n for _, n in adds}):
        c3.add_many([k for k, m in adds if m == n], n)
    assert (c2.bloom_numpy() == c3.bloom_numpy()).all() and c2.elements_added == c3.elements_added


def test_batches_vs_golden_and_oracle(pb, orc, golden):
    g = golden["cbloom_batch"]
    c = pb.CountingBloomFilter(est_elements=g["est"], false_positive_rate=g["fpr"])
    keys = orc.uniform_keys(0, g["n_add"])
    c.add_many(keys)
    assert md5(c.bloom_numpy().tobytes()) == g["counters_md5_after_add"]
    c.remove_many(keys[: g["n_remove"]])
    assert md5(c.bloom_numpy().tobytes()) == g["counters_md5"] and c.elements_added == g["elements_added"]
    assert c.check_many(keys[::100]).tolist() == g["checks_0_300_step100"]
    assert c._cnt_number_bits_set() == g["nonzero"] and c.estimate_elements() == g["estimate_elements"]
    # removals that run counters dry: the optimistic parallel pass must notice and replay in order
    gc = golden["cbloom_clamped"]
    c = pb.CountingBloomFilter(est_elements=gc["est"], false_positive_rate=gc["fpr"])
    c.add_many(orc.uniform_keys(0, 500))
    c.remove_many(np.concatenate([orc.uniform_keys(0, 500)] * 3), 2)
    assert c.bloom_numpy().tolist() == gc["counters"] and c.elements_added == gc["elements_added"]
    # larger, vs the oracle: ragged str keys with repeats, weights, interleaved removes
    rng = np.random.default_rng(9)
    c = pb.CountingBloomFilter(est_elements=200_000, false_positive_rate=0.02)
    o = orc.CountingBloom(c.number_bits, c.number_hashes)
    pool = [f"key-{i}-{'z' * (i % 11)}" for i in range(150_000)]
    for rnd in range(3):
        batch = [pool[i] for i in rng.integers(0, len(pool), size=120_000)]
        n = int(rng.integers(1, 4))
        c.add_many(batch, n)
        o.add(orc.pack(batch), n)
        assert (c.bloom_numpy() == o.bloom).all() and c.elements_added == o.elements_added
        rem = [pool[i] for i in rng.integers(0, len(pool), size=60_000)]
        c.remove_many(rem, 1 + rnd)
        o.remove(orc.pack(rem), 1 + rnd)
        assert (c.bloom_numpy() == o.bloom).all() and c.elements_added == o.elements_added, rnd
    probes = pool[:5000] + [f"absent-{i}" for i in range(5000)]
    assert (c.check_many(probes) == o.check(orc.pack(probes))).all()
    # 16-byte keys take the fused kernels
    k16 = orc.uniform_keys(50, 300_000)
    c.add_many(k16, 2)
    o.add(orc.pack(k16), 2)
    assert (c.bloom_numpy() == o.bloom).all()
    assert (c.check_many(k16[:4000]) == o.check(orc.pack(k16[:4000]))).all()


def test_saturation(pb, orc):
    c = pb.CountingBloomFilter(est_elements=10, false_positive_rate=0.05)
    assert c.add("a", 2**32 - 5) == 2**32 - 5
    assert c.add("a", 3) == 2**32 - 2
    assert c.add("a", 100) == 2**32 - 1  # countingbloom.py:147-149
    assert c.check("a") == 2**32 - 1
    assert c.remove("a") == 2**32 - 1  # :199-200: saturated counters are never decremented
    assert c.check("a") == 2**32 - 1
    c.add_many(["a", "b", "a"], 2**31)  # batch on the exact (CAS) path
    assert c.check("a") == 2**32 - 1 and c.check("b") >= 2**31


def test_set_algebra_and_plugin(pb, golden):
    g = golden["cbloom_algebra"]
    a = pb.CountingBloomFilter(est_elements=100, false_positive_rate=0.05)
    b = pb.CountingBloomFilter(est_elements=100, false_positive_rate=0.05)
    for i in range(60):
        a.add(f"a{i % 40}")
        b.add(f"a{i % 25 + 30}", 2)
    u, x = a.union(b), a.intersection(b)
    assert u.bloom_numpy().tolist() == g["union"] and x.bloom_numpy().tolist() == g["intersection"]
    assert a.jaccard_index(b) == g["jaccard"]
    assert u.elements_added == g["union_elements_added"] and x.elements_added == g["intersection_elements_added"]
    with pytest.raises(TypeError):
        a.union(pb.BloomFilter(100, 0.05))
    with pytest.raises(pb.SimilarityError):
        a.union(pb.CountingBloomFilter(est_elements=101, false_positive_rate=0.05))
    # custom hash_function -> host hashes -> pre-hashed kernels
    c = pb.CountingBloomFilter(est_elements=500, false_positive_rate=0.02, hash_function=pb.hashes.default_md5)
    words = [f"w{i % 77}" for i in range(400)]
    c.add_many(words)
    ref = np.zeros(c.number_bits, dtype=np.int64)
    for w in words:
        for h in pb.hashes.default_md5(w, c.number_hashes):
            ref[h % c.number_bits] += 1
    assert (c.bloom_numpy() == ref).all()
    assert c.check("w5") == min(ref[h % c.number_bits] for h in pb.hashes.default_md5("w5", c.number_hashes))
    assert c.remove("w5") == c.check("w5")


@pytest.mark.parametrize("est,fpr,wl_bits", [(300_000, 0.01, 20), (200_000, 0.0001, 22), (40_000, 0.2, 16)])
def test_partitioned_add_equals_direct_and_oracle(pb, orc, est, fpr, wl_bits):
    """the two-pass add (indices binned by window of counters, then added while the window is L2 resident) against the
    oracle: several windows, repeated keys (one index listed twice counts twice), amounts > 1, a skewed batch that
    overflows its sublists into the side list, ragged keys, and the saturating fallback once a counter could overflow"""
    ctx = pb.default_context()
    keys = orc.uniform_keys(0, 400_000)
    keys[::9] = keys[:len(keys[::9])]
    skew = np.repeat(orc.uniform_keys(7, 2), 150_000, axis=0)
    words = [f"w{i}-{'y' * (i % 11)}" for i in range(50_000)]
    try:
        ctx.set_option("bloom_insert_mode", 2)
        ctx.set_option("bloom_window_log2_bits", wl_bits)  # windows of 2^(bits-5) counters
        c = pb.CountingBloomFilter(est, fpr)
        o = orc.CountingBloom(c.number_bits, c.number_hashes)
        c.add_many(keys), o.add(orc.pack(keys))
        assert (c.bloom_numpy() == o.bloom).all()
        c.add_many(keys[:100_000], 3), o.add(orc.pack(keys[:100_000]), 3)
        c.add_many(skew), o.add(orc.pack(skew))
        c.add_many(words, 2), o.add(orc.pack(words), 2)
        assert (c.bloom_numpy() == o.bloom).all() and c.elements_added == o.elements_added
        probes = np.concatenate([keys[:5000], orc.uniform_keys(9_000_000, 5000)])
        assert (c.check_many(probes) == o.check(orc.pack(probes))).all()
        before = c.bloom_numpy().astype(np.uint64)
        c.add_many(keys[:50], 2**31)  # from here on a counter could reach 2^32: the saturating direct kernel takes over
        c.add_many(keys[:50], 2**31)
        inc = orc.CountingBloom(c.number_bits, c.number_hashes)
        inc.add(orc.pack(keys[:50]), 1)  # how often each counter is hit by one pass over these 50 keys
        want = np.minimum(before + inc.bloom.astype(np.uint64) * 2**32, 2**32 - 1)  # countingbloom.py:147-149
        assert (c.bloom_numpy() == want).all() and int(c.bloom_numpy().max()) == 2**32 - 1
    finally:
        ctx.set_option("bloom_insert_mode", 0)
        ctx.set_option("bloom_window_log2_bits", 27)
