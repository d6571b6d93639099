"""Stateful wrappers on the device: ExpandingBloomFilter / RotatingBloomFilter batches against the
reference's test literals, vectors recorded from the pure-Python reference (tests/golden/make_golden_r2_wrappers.py)
and the oracle's one-key-at-a-time loop.  A batch must reach exactly the state the reference reaches adding the keys
one by one: bitmaps, per-filter counts, number of filters and the exported bytes."""

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


def stack_state(f) -> dict:
    return {"n_blooms": len(f._blooms), "per_bloom_added": [b.elements_added for b in f._blooms],
            "per_bloom_md5": [md5(b.bloom_numpy().tobytes()) for b in f._blooms], "elements_added": f.elements_added,
            "export_md5": md5(bytes(f))}


def oracle_state(o) -> dict:
    return {"n_blooms": len(o.blooms), "per_bloom_added": [b.elements_added for b in o.blooms],
            "per_bloom_md5": [md5(b.bloom.tobytes()) for b in o.blooms], "elements_added": o.elements_added,
            "export_md5": md5(o.export())}


def test_expanding_reference_literals(pb, golden):
    # tests/expandingbloom_test.py:25-54: one batch == 120 single adds
    e = pb.ExpandingBloomFilter(est_elements=10, false_positive_rate=0.05)
    assert (e.expansions, e.false_positive_rate, e.estimated_elements, e.elements_added) == (0, 0.05, 10, 0)
    e.add_many([f"{i}" for i in range(120)])
    assert e.expansions == 8 and e.elements_added == 120 and stack_state(e) == golden["ebf_120_no_force"]
    one = pb.ExpandingBloomFilter(est_elements=10, false_positive_rate=0.05)
    for i in range(120):
        one.add(f"{i}")
    assert stack_state(one) == golden["ebf_120_no_force"]
    f = pb.ExpandingBloomFilter(est_elements=10, false_positive_rate=0.05)
    f.add_many([f"{i}" for i in range(100)], force=True)
    assert f.expansions == 9 and stack_state(f) == golden["ebf_100_force"]
    # :56-83 check / contains
    c = pb.ExpandingBloomFilter(est_elements=30, false_positive_rate=0.05)
    c.add_many([f"{i}" for i in range(100)])
    c.add("this is a test")
    c.add("this is another test")
    assert c.expansions > 1 and c.elements_added == 102
    assert c.check("this is a test") and "this is another test" in c
    assert not c.check("this is yet another test!") and "this is not another test" not in c
    # :85-97 push; :99-137 wire format
    p = pb.ExpandingBloomFilter(est_elements=25, false_positive_rate=0.05)
    assert md5(bytes(p)) == "eb5769ae9babdf7b37d6ce64d58812bc"
    p.push(), p.push(), p.push()
    assert p.expansions == 3 and p.elements_added == 0
    back = pb.ExpandingBloomFilter.frombytes(bytes(c))
    assert bytes(back) == bytes(c) and back.check("this is a test") and back.expansions == c.expansions
    assert back.elements_added == 102 and not back.check("this is yet another test!")


def test_expanding_stream_vs_reference_and_oracle(pb, orc, golden):
    g = golden["ebf_stream"]
    keys = np.concatenate([orc.uniform_keys(0, g["n_unique"]), orc.uniform_keys(0, g["n_repeat"])])
    e = pb.ExpandingBloomFilter(est_elements=g["est"], false_positive_rate=g["fpr"])
    for q in range(4):
        e.add_many(keys[q * 6250 : (q + 1) * 6250])
        assert stack_state(e) == g["quarters"][q]
    hits = np.flatnonzero(e.check_many(orc.uniform_keys(100000, 2000))) + 100000
    assert hits.tolist() == g["probe_hits_100000_102000"]
    # the whole stream as ONE batch, from a CUDA tensor, reaches the same state
    import torch

    one = pb.ExpandingBloomFilter(est_elements=g["est"], false_positive_rate=g["fpr"])
    one.add_many(torch.from_numpy(keys).cuda())
    assert stack_state(one) == g["quarters"][3]
    res = one.check_many(torch.from_numpy(keys[:5000]).cuda())
    assert res.is_cuda and bool(res.all())


@pytest.mark.parametrize("est,fpr,n,dup_every", [(50_000, 0.01, 400_000, 7), (1000, 0.2, 30_000, 0), (3, 0.3, 500, 2)])
def test_expanding_batch_equals_one_at_a_time(pb, orc, est, fpr, n, dup_every):
    """larger batches against the oracle loop: many growth steps inside one batch, high false-positive rates (keys
    skipped because of earlier keys of the same batch), repeated keys"""
    keys = orc.uniform_keys(10_000, n)
    if dup_every:
        keys[::dup_every] = keys[0 : len(keys[::dup_every])]  # repeats of early keys all over the batch
    e = pb.ExpandingBloomFilter(est_elements=est, false_positive_rate=fpr)
    o = orc.ExpandingBloom(est, fpr)
    half = n // 3
    e.add_many(keys[:half]), o.add(orc.pack(keys[:half]))
    assert stack_state(e) == oracle_state(o)
    e.add_many(keys[half:]), o.add(orc.pack(keys[half:]))
    assert stack_state(e) == oracle_state(o)
    probes = orc.uniform_keys(5_000_000, 20_000)
    assert (e.check_many(probes) == o.check(orc.pack(probes))).all()


def test_expanding_ragged_str_and_plugin_hash(pb, orc):
    words = [f"word-{i}-{'x' * (i % 13)}" for i in range(5000)]
    e = pb.ExpandingBloomFilter(est_elements=400, false_positive_rate=0.05)
    e.add_many(words)
    o = orc.ExpandingBloom(400, 0.05)
    o.add(orc.pack(words))
    assert stack_state(e) == oracle_state(o)

    def shifted(key, depth):  # a user hash_function: the k-seed FNV of the key with a suffix
        return pb.hashes.default_fnv_1a(key + "#", depth)

    p = pb.ExpandingBloomFilter(est_elements=400, false_positive_rate=0.05, hash_function=shifted)
    p.add_many(words)
    o = orc.ExpandingBloom(400, 0.05)
    o.add(orc.pack([w + "#" for w in words]))
    assert stack_state(p) == oracle_state(o)
    assert p.check(words[17]) and p.check_alt(shifted(words[17], p._blooms[0].number_hashes))
    p.add_alt(shifted("brand new", p._blooms[0].number_hashes))
    assert p.check("brand new") and p.elements_added == 5001


def test_rotating_stream_vs_reference(pb, orc, golden):
    g = golden["rbf_stream"]
    r = pb.RotatingBloomFilter(est_elements=g["est"], false_positive_rate=g["fpr"], max_queue_size=g["queue"])
    assert r.max_queue_size == g["queue"] and r.current_queue_size == 1
    for h in range(2):
        r.add_many(orc.uniform_keys(h * 3000, 3000))
        assert stack_state(r) == g["halves"][h]
    present = np.flatnonzero(r.check_many(orc.uniform_keys(0, 6000)[::10])) * 10
    assert present.tolist() == g["present_step10"]
    r.push()
    assert stack_state(r) == g["after_push"]
    r.pop()
    assert stack_state(r) == g["after_pop"]
    r.add_many(orc.uniform_keys(6000, 500), force=True)
    assert stack_state(r) == g["after_force_500"]
    back = pb.RotatingBloomFilter.frombytes(bytes(r), max_queue_size=g["queue"])
    assert bytes(back) == bytes(r) and back.current_queue_size == r.current_queue_size
    # one batch == the two halves
    one = pb.RotatingBloomFilter(est_elements=g["est"], false_positive_rate=g["fpr"], max_queue_size=g["queue"])
    one.add_many(orc.uniform_keys(0, 6000))
    assert stack_state(one) == g["halves"][1]


def test_rotating_reference_literals(pb):
    # tests/expandingbloom_test.py:168-200
    blm = pb.RotatingBloomFilter(est_elements=10, false_positive_rate=0.05, max_queue_size=5)
    blm.add("test")
    assert blm.expansions == 0
    blm.add_many([f"{i}" for i in range(10)], force=True)
    assert blm.expansions == 1 and blm.current_queue_size == 2 and blm.check("test")
    for lo, q in ((10, 3), (20, 4), (30, 5)):
        blm.add_many([f"{i}" for i in range(lo, lo + 10)], force=True)
        assert blm.check("test") and blm.current_queue_size == q
    blm.add_many([f"{i}" for i in range(40, 50)], force=True)
    assert not blm.check("test") and blm.current_queue_size == 5 and blm.elements_added == 51
    # :202-252 push / pop
    blm = pb.RotatingBloomFilter(est_elements=10, false_positive_rate=0.05, max_queue_size=5)
    blm.add("test")
    for q in (2, 3, 4, 5):
        blm.push()
        assert blm.current_queue_size == q and "test" in blm
    blm.push()
    assert blm.current_queue_size == 5 and "test" not in blm
    blm.add("that")
    for q in (4, 3, 2, 1):
        blm.pop()
        assert blm.current_queue_size == q and "that" in blm
    with pytest.raises(pb.RotatingBloomFilterError) as ex:
        blm.pop()
    assert str(ex.value) == "Popping a Bloom Filter will result in an unusable system!"


@pytest.mark.parametrize("queue", [1, 3])
def test_rotating_batch_equals_one_at_a_time(pb, orc, queue):
    keys = orc.uniform_keys(77, 60_000)
    keys[::5] = keys[0:12_000]
    r = pb.RotatingBloomFilter(est_elements=2500, false_positive_rate=0.1, max_queue_size=queue)
    o = orc.ExpandingBloom(2500, 0.1, max_queue_size=queue)
    r.add_many(keys[:25_000]), o.add(orc.pack(keys[:25_000]))
    assert stack_state(r) == oracle_state(o)
    r.add_many(keys[25_000:]), o.add(orc.pack(keys[25_000:]))
    assert stack_state(r) == oracle_state(o)


# ---------------------------------------------------------------- CountingCuckooFilter (cuckoo/countingcuckoo.py)
def test_counting_cuckoo_vs_reference(pb, orc, golden):
    g = golden["ccf"]
    keys = orc.uniform_keys(0, 1300)
    f = pb.CountingCuckooFilter(capacity=g["capacity"], bucket_size=g["bucket_size"], max_swaps=g["max_swaps"], auto_expand=False)
    f.add_many(keys[g["draws"]])
    assert f.bins() == [tuple(x) for x in g["bins"]]
    assert (f.elements_added, f.unique_elements) == (g["elements_added"], g["unique_elements"])
    assert f.load_factor() == g["load_factor"]
    assert f.check_many(keys).tolist() == g["check_0_1300"]
    assert f.remove_many(keys[0:1300:3]).tolist() == g["removed_step3"]
    assert f.bins() == [tuple(x) for x in g["bins_after_remove"]]
    assert (f.elements_added, f.unique_elements) == (g["elements_added_after_remove"], g["unique_after_remove"])
    # wire format round trip
    back = pb.CountingCuckooFilter.frombytes(bytes(f))
    assert back.bins() == f.bins() and back.capacity == f.capacity
    assert (back.elements_added, back.unique_elements) == (f.elements_added, f.unique_elements)
    assert (back.check_many(keys) == f.check_many(keys)).all()


def test_counting_cuckoo_reference_literals(pb, golden):
    # tests/countingcuckoo_test.py style: single adds, counts, contains, remove down to zero
    f = pb.CountingCuckooFilter(capacity=100, max_swaps=5)
    for w, n in (("this is a test", 3), ("this is another test", 1)):
        for _ in range(n):
            f.add(w)
    assert f.check("this is a test") == 3 and f.check("this is another test") == 1 and f.check("nope") == 0
    assert "this is a test" in f and "nope" not in f
    assert f.elements_added == 4 and f.unique_elements == 2
    assert f.remove("this is a test") and f.check("this is a test") == 2 and f.elements_added == 3 and f.unique_elements == 2
    assert f.remove("this is another test") and f.check("this is another test") == 0 and f.unique_elements == 1
    assert not f.remove("this is another test") and not f.remove("nope") and f.elements_added == 2
    b = f.buckets
    assert sum(len(x) for x in b) == 1 and [x for x in b if x][0][0].count == 2
    # export bytes where no eviction happens (slot order = first insertion order), inserted in key order
    e = golden["ccf_export"]
    ctx = pb.default_context()
    ctx.set_option("cuckoo_serial", 1)
    try:
        s = pb.CountingCuckooFilter(capacity=e["capacity"], bucket_size=4, max_swaps=5)
        s.add_many([str(i % 400) for i in range(600)])
    finally:
        ctx.set_option("cuckoo_serial", 0)
    assert len(bytes(s)) == e["export_len"] and md5(bytes(s)) == e["export_md5"]
    assert (s.elements_added, s.unique_elements) == (e["elements_added"], e["unique_elements"])


def test_counting_cuckoo_batch_equals_one_at_a_time(pb, orc):
    """a skewed stream at high load: many repeats inside one batch, evictions, removals that run counts dry"""
    rng = np.random.default_rng(5)
    pool = orc.uniform_keys(900_000, 60_000)
    draws = np.minimum(rng.zipf(1.3, 150_000) - 1, 59_999)
    stream = pool[draws]
    cap = 1 << 14  # 65 536 slots; ~29 000 distinct keys
    f = pb.CountingCuckooFilter(capacity=cap, bucket_size=4, max_swaps=500, auto_expand=False)
    o = orc.CountingCuckoo(cap, 4, 500)
    f.add_many(stream[:100_000]), o.add(orc.pack(stream[:100_000]))
    f.add_many(stream[100_000:]), o.add(orc.pack(stream[100_000:]))
    assert f.bins() == o.bins() and (f.elements_added, f.unique_elements) == (o.elements_added, o.unique_elements)
    probes = np.concatenate([pool[:5000], orc.uniform_keys(5_000_000, 5000)])
    assert (f.check_many(probes) == o.check(orc.pack(probes))).all()
    rem = pool[np.minimum(rng.zipf(1.3, 40_000) - 1, 59_999)]  # popular keys are removed more often than they were added
    got, want = f.remove_many(rem), o.remove(orc.pack(rem))
    assert (got == want).all()
    assert f.bins() == o.bins() and (f.elements_added, f.unique_elements) == (o.elements_added, o.unique_elements)
    assert (f.check_many(probes) == o.check(orc.pack(probes))).all()


def test_counting_cuckoo_expand_full_and_plugin_hash(pb, orc):
    keys = orc.uniform_keys(0, 3000)
    f = pb.CountingCuckooFilter(capacity=100, bucket_size=2, max_swaps=20)  # auto_expand: 200 slots -> grows
    f.add_many(np.concatenate([keys, keys[:1000]]))
    assert f.capacity > 100 and f.unique_elements == 3000 and f.elements_added == 4000
    c = f.check_many(keys)
    assert (c[:1000] == 2).all() and (c[1000:] == 1).all()
    full = pb.CountingCuckooFilter(capacity=100, bucket_size=2, max_swaps=20, auto_expand=False)
    with pytest.raises(pb.CuckooFilterFullError):
        full.add_many(keys)
    assert full.unique_elements == len(full.bins()) <= 200 and full.elements_added == sum(c for _, c in full.bins())

    def md5_int(key):
        if isinstance(key, str):
            key = key.encode("utf-8")
        return int.from_bytes(hashlib.md5(key).digest()[:8], "big")

    words = [f"w{i % 700}" for i in range(2000)]
    p = pb.CountingCuckooFilter(capacity=400, bucket_size=4, max_swaps=100, hash_function=md5_int)
    p.add_many(words)
    assert p.unique_elements == 700 and p.elements_added == 2000
    assert p.check("w5") == 3 and p.check("w699") == 2 and p.check("zzz") == 0
    assert p.remove("w699") and p.remove("w699") and not p.remove("w699") and p.unique_elements == 699


# ---------------------------------------------------------------- HeavyHitters / StreamThreshold (countminsketch.py:532-831)
def _names(orc, ranks):
    return [bytes(k).hex() for k in orc.rank_keys(np.asarray(ranks, dtype=np.uint64))]


def test_heavy_hitters_vs_reference(pb, orc, golden):
    import struct

    h = golden["heavy_hitters"]
    names = _names(orc, h["ranks"])
    hh = pb.HeavyHitters(num_hitters=h["num_hitters"], width=h["width"], depth=h["depth"])
    rets = hh.add_many(names)
    assert md5(struct.pack(f"<{len(rets)}q", *rets.tolist())) == h["returns_md5"]
    assert hh.heavy_hitters == h["heavy_hitters"] and md5(hh.bins_numpy().tobytes()) == h["bins_md5"]
    assert hh.elements_added == len(names) and hh.number_heavy_hitters == h["num_hitters"]
    # the same stream in uneven pieces and single adds
    two = pb.HeavyHitters(num_hitters=h["num_hitters"], width=h["width"], depth=h["depth"])
    two.add_many(names[:7]), two.add(names[7]), two.add_many(names[8:12_345]), two.add_many(names[12_345:])
    assert two.heavy_hitters == h["heavy_hitters"] and md5(two.bins_numpy().tobytes()) == h["bins_md5"]
    with pytest.raises(pb.NotSupportedError):
        hh.remove(names[0])
    with pytest.raises(pb.NotSupportedError):
        hh.join(two)
    assert "Heavy Hitters Count-Min Sketch" in str(hh) and f"Number Recorded: {h['num_hitters']}" in str(hh)
    back = pb.HeavyHitters.frombytes(bytes(hh), num_hitters=5)
    assert back.check(names[0]) == hh.check(names[0]) and back.heavy_hitters == {}
    hh.clear()
    assert hh.heavy_hitters == {} and hh.elements_added == 0


def test_stream_threshold_vs_reference(pb, orc, golden):
    import struct

    t = golden["stream_threshold"]
    names = _names(orc, golden["heavy_hitters"]["ranks"])
    st = pb.StreamThreshold(threshold=t["threshold"], width=t["width"], depth=t["depth"])
    rets = st.add_many(names)
    assert md5(struct.pack(f"<{len(rets)}q", *rets.tolist())) == t["returns_md5"]
    assert st.meets_threshold == t["meets_threshold"] and st.threshold == t["threshold"]
    # remove below the threshold drops the key again (:818-831)
    top = max(st.meets_threshold, key=st.meets_threshold.get)
    cnt = st.meets_threshold[top]
    assert st.remove(top, cnt - t["threshold"] + 1) == t["threshold"] - 1 and top not in st.meets_threshold
    assert st.add(top, 5) == t["threshold"] + 4 and st.meets_threshold[top] == t["threshold"] + 4


@pytest.mark.parametrize("query_type", ["min", "mean", "mean-min"])
def test_add_many_returns_equals_one_at_a_time(pb, orc, query_type):
    """every key's own return value inside one batch: heavy collisions (narrow table), per-key amounts, saturation"""
    rng = np.random.default_rng(11)
    ranks = np.minimum(rng.zipf(1.2, 200_000), 1 << 40).astype(np.uint64)
    keys = orc.rank_keys(ranks)
    amounts = rng.integers(0, 5, keys.shape[0]).astype(np.int64)
    amounts[1000] = (1 << 31) - 50  # drives the counters of one key to INT32_MAX
    c = pb.CountMinSketch(width=997, depth=4 if query_type == "mean-min" else 5)
    c.query_type = query_type
    o = orc.CMS(997, c.depth, query_type)
    got = c.add_many_returns(keys[:50_000])
    want = o.add(orc.pack(keys[:50_000]), 1, want_returns=True)
    assert (got.cpu().numpy() == want).all() and (c.bins_numpy() == o.bins).all()
    got = c.add_many_returns(keys[50_000:], amounts[50_000:])
    want = o.add(orc.pack(keys[50_000:]), amounts[50_000:], want_returns=True)
    assert (got.cpu().numpy() == want).all() and (c.bins_numpy() == o.bins).all()
    assert c.elements_added == o.elements_added
    import torch

    dev = c.add_many_returns(torch.from_numpy(keys[:1000]).cuda(), 2)
    assert (dev.cpu().numpy() == o.add(orc.pack(keys[:1000]), 2, want_returns=True)).all()


def test_heavy_hitters_large_batch_vs_oracle(pb, orc):
    rng = np.random.default_rng(3)
    ranks = np.minimum(rng.zipf(1.1, 300_000), 1 << 40).astype(np.uint64)
    keys = orc.rank_keys(ranks)
    names = [k.tobytes() for k in keys]
    hh = pb.HeavyHitters(num_hitters=50, width=1 << 12, depth=5)
    o = orc.HeavyHitters(50, 1 << 12, 5)
    got = hh.add_many(keys)  # a numpy batch: dictionary keys are the rows' bytes
    want = o.add_tracked(names, orc.pack(keys))
    assert (got == want).all() and hh.heavy_hitters == o.top_x
    st = pb.StreamThreshold(threshold=500, width=1 << 12, depth=5)
    os_ = orc.StreamThreshold(500, 1 << 12, 5)
    st.add_many(keys), os_.add_tracked(names, orc.pack(keys))
    assert st.meets_threshold == os_.meets and len(os_.meets) > 10


# ---------------------------------------------------------------- the reference's own test literals for the wrappers
class _serial:
    """key-order insertion (one device thread): reproduces the reference's slot order, which its export md5s pin"""

    def __init__(self, pb):
        self.ctx = pb.default_context()

    def __enter__(self):
        self.ctx.set_option("cuckoo_serial", 1)

    def __exit__(self, *a):
        self.ctx.set_option("cuckoo_serial", 0)


def test_counting_cuckoo_reference_test_file(pb, tmp_path):
    # tests/countingcuckoo_test.py:24-46 constructor properties
    cko = pb.CountingCuckooFilter()
    assert (cko.capacity, cko.bucket_size, cko.max_swaps, cko.expansion_rate, cko.auto_expand) == (10000, 4, 500, 2, True)
    cko = pb.CountingCuckooFilter(capacity=100, bucket_size=2, max_swaps=5, expansion_rate=4, auto_expand=False)
    assert (cko.capacity, cko.bucket_size, cko.max_swaps, cko.expansion_rate, cko.auto_expand) == (100, 2, 5, 4, False)
    # :100-128 lots / full
    cko = pb.CountingCuckooFilter(capacity=100, bucket_size=2, max_swaps=100)
    cko.add_many([str(i) for i in range(125)])
    assert cko.elements_added == 125
    with pytest.raises(pb.CuckooFilterFullError) as ex:
        full = pb.CountingCuckooFilter(capacity=100, bucket_size=2, max_swaps=100, auto_expand=False)
        for i in range(175):
            full.add(str(i))
    assert str(ex.value) == "The CountingCuckooFilter is currently full"
    # :177-197 load factor
    cko = pb.CountingCuckooFilter(capacity=100, bucket_size=2, max_swaps=10)
    assert cko.load_factor() == 0.0
    cko.add_many([str(i) for i in range(50)])
    assert cko.load_factor() == 0.25
    cko.add_many([str(i + 50) for i in range(50)])
    assert cko.load_factor() == (0.25 if cko.capacity == 200 else 0.50)
    cko.add_many([str(i) for i in range(100)])
    assert cko.load_factor() == (0.25 if cko.capacity == 200 else 0.50) and cko.elements_added == 200
    # :199-253 export / bytes / frombytes / load
    with _serial(pb):
        cko = pb.CountingCuckooFilter(capacity=1000, bucket_size=2, auto_expand=False)
        for i in range(100):
            cko.add(str(i))
    assert md5(bytes(cko)) == "6a98c2df1ec9fbb4f75f8e6392696b9b"
    path = tmp_path / "c.cck"
    cko.export(path)
    assert md5(path.read_bytes()) == "6a98c2df1ec9fbb4f75f8e6392696b9b"
    cko2 = pb.CountingCuckooFilter.frombytes(bytes(cko))
    assert bytes(cko2) == bytes(cko) and all(cko2.check(str(i)) for i in range(100)) and not cko2.check("999")
    ckf = pb.CountingCuckooFilter(filepath=path)
    assert (ckf.check_many([str(i) for i in range(100)]) == 1).all()
    assert (ckf.capacity, ckf.bucket_size, ckf.max_swaps, ckf.load_factor()) == (1000, 2, 500, 0.05)
    # :255-273 expand
    cko = pb.CountingCuckooFilter()
    cko.add_many([str(i) for i in range(200)])
    cko.expand()
    assert (cko.check_many([str(i) for i in range(200)]) > 0).all() and cko.capacity == 20000
    cko = pb.CountingCuckooFilter(capacity=100, bucket_size=2, max_swaps=100)
    for i in range(375):
        cko.add(str(i))
    assert cko.capacity == 400 and cko.elements_added == 375 and (cko.check_many([str(i) for i in range(375)]) > 0).all()
    # :275-300 bin repr / str
    cko = pb.CountingCuckooFilter(capacity=1, bucket_size=2, max_swaps=100)
    cko.add("this is a test")
    assert str(cko.buckets[0]) == "[(fingerprint:4280557824 count:1)]"
    cko = pb.CountingCuckooFilter(capacity=100, bucket_size=2, max_swaps=100)
    cko.add_many([str(i) for i in range(75)])
    assert str(cko) == ("CountingCuckooFilter:\n\tCapacity: 100\n\tTotal Bins: 200\n\tLoad Factor: 37.5%\n\tInserted Elements: 75\n"
                        "\tMax Swaps: 100\n\tExpansion Rate: 2\n\tAuto Expand: True")


def test_counting_cuckoo_error_rate_reference_test_file(pb, tmp_path):
    # tests/countingcuckoo_test.py:302-420
    cko = pb.CountingCuckooFilter.init_error_rate(0.00001)
    assert (cko.capacity, cko.bucket_size, cko.max_swaps, cko.expansion_rate, cko.auto_expand) == (10000, 4, 500, 2, True)
    assert (cko.fingerprint_size, cko.fingerprint_size_bits, cko.error_rate) == (3, 20, 0.00001)
    for w in ("this is a test", "this is another test", "this is yet another test"):
        cko.add(w)
    assert cko.elements_added == 3 and cko.check("this is a test") and "this is yet another test" in cko
    assert not cko.check("this is not another test") and "this is not a test" not in cko
    with _serial(pb):
        cko = pb.CountingCuckooFilter.init_error_rate(0.00001)
        for i in range(1000):
            cko.add(str(i))
        assert md5(bytes(cko)) == "f68767bd97b21426f5d2315fb38961ad"
        cko = pb.CountingCuckooFilter.init_error_rate(0.00001)
        for i in range(1000):
            cko.add(str(i))
            if i % 2 == 1:
                cko.add(str(i))
    path = tmp_path / "c.cko"
    cko.export(path)
    assert md5(path.read_bytes()) == "88bc3a08bfc967f9ba60e9d57c21207f"
    ckf = pb.CountingCuckooFilter.load_error_rate(error_rate=0.00001, filepath=path)
    assert ckf.check_many([str(i) for i in range(1000)]).tolist() == [(i % 2) + 1 for i in range(1000)]
    assert (ckf.capacity, ckf.bucket_size, ckf.max_swaps, ckf.expansion_rate, ckf.auto_expand) == (10000, 4, 500, 2, True)
    assert (ckf.fingerprint_size_bits, ckf.fingerprint_size, ckf.error_rate, ckf.load_factor()) == (20, 3, 0.00001, 0.025)
    cko = pb.CountingCuckooFilter.init_error_rate(0.00001, capacity=3000)
    cko.add_many([str(i) for i in range(1000)])
    cko2 = pb.CountingCuckooFilter.frombytes(bytes(cko), error_rate=0.00001)
    assert bytes(cko2) == bytes(cko) and (cko2.check_many([str(i) for i in range(1000)]) > 0).all()
    assert not cko2.check("9999") and cko2.capacity == 3000


def test_heavy_hitters_reference_test_file(pb, tmp_path):
    # tests/countminsketch_test.py:567-735
    for kw in ({"width": 1000, "depth": 5}, {"confidence": 0.96875, "error_rate": 0.002}):
        hh1 = pb.HeavyHitters(num_hitters=1000, **kw)
        assert (hh1.width, hh1.depth, hh1.confidence, hh1.error_rate, hh1.elements_added) == (1000, 5, 0.96875, 0.002, 0)
        assert hh1.heavy_hitters == {} and hh1.number_heavy_hitters == 1000
    hh1 = pb.HeavyHitters(num_hitters=2, width=1000, depth=5)
    assert [hh1.add("this is a test") for _ in range(3)] == [1, 2, 3]
    assert hh1.add("this is also a test") == 1 and hh1.add("this is not a test") == 1 and hh1.add("this is not a test") == 2
    assert hh1.heavy_hitters == {"this is a test": 3, "this is not a test": 2}
    assert [hh1.add("this is also a test") for _ in range(3)] == [2, 3, 4]
    assert hh1.heavy_hitters == {"this is a test": 3, "this is also a test": 4}
    hh1 = pb.HeavyHitters(num_hitters=2, width=1000, depth=5)
    assert hh1.add("this is a test", 3) == 3 and hh1.add("this is also a test") == 1 and hh1.add("this is not a test", 2) == 2
    assert hh1.heavy_hitters == {"this is a test": 3, "this is not a test": 2}
    assert hh1.add("this is also a test", 3) == 4
    assert hh1.heavy_hitters == {"this is a test": 3, "this is also a test": 4}
    assert [hh1.add("this is not a test", 2) for _ in range(4)] == [4, 6, 8, 10]
    assert hh1.heavy_hitters == {"this is not a test": 10, "this is also a test": 4}
    # the same sequence as ONE ordered batch with per-key amounts
    hb = pb.HeavyHitters(num_hitters=2, width=1000, depth=5)
    seq = ["this is a test", "this is also a test", "this is not a test", "this is also a test"] + ["this is not a test"] * 4
    assert hb.add_many(seq, np.array([3, 1, 2, 3, 2, 2, 2, 2])).tolist() == [3, 1, 2, 4, 4, 6, 8, 10]
    assert hb.heavy_hitters == {"this is not a test": 10, "this is also a test": 4}
    with pytest.raises(pb.NotSupportedError) as ex:
        hh1.remove("this is a test")
    assert str(ex.value) == ("Unable to remove elements in the HeavyHitters class as it is an un supported action (and does not"
                             "make sense)!")
    hh1 = pb.HeavyHitters(num_hitters=1000, width=1000, depth=5)
    assert hh1.add("this is a test", 100) == 100 and hh1.elements_added == 100 and hh1.heavy_hitters == {"this is a test": 100}
    assert md5(bytes(hh1)) == "fb1c39dd1a73f1ef0d7fc79f60fc028e"
    path = tmp_path / "h.cms"
    hh1.export(path)
    assert md5(path.read_bytes()) == "fb1c39dd1a73f1ef0d7fc79f60fc028e"
    hh2 = pb.HeavyHitters(num_hitters=1000, filepath=path)
    assert (hh2.width, hh2.depth, hh2.elements_added, hh2.check("this is a test"), hh2.heavy_hitters) == (1000, 5, 100, 100, {})
    assert hh2.add("this is a test", 1) == 101 and hh2.heavy_hitters == {"this is a test": 101}
    hh3 = pb.HeavyHitters.frombytes(bytes(hh1), num_hitters=500)
    assert (hh3.width, hh3.depth, hh3.number_heavy_hitters, hh3.elements_added) == (1000, 5, 500, 100)
    assert bytes(hh3) == bytes(hh1) and hh3.check("this is a test") == 100
    hs = pb.HeavyHitters(num_hitters=2, width=1000, depth=5)
    hs.add("this is a test", 100)
    assert str(hs) == ("Heavy Hitters Count-Min Sketch:\n\tWidth: 1000\n\tDepth: 5\n\tConfidence: 0.96875\n\tError Rate: 0.002\n"
                       "\tElements Added: 100\n\tNumber Hitters: 2\n\tNumber Recorded: 1")
    hh1.clear()
    assert hh1.elements_added == 0 and hh1.heavy_hitters == {}


def test_stream_threshold_reference_test_file(pb, tmp_path):
    # tests/countminsketch_test.py:759-944
    for kw in ({"width": 1000, "depth": 5}, {"confidence": 0.96875, "error_rate": 0.002}):
        st1 = pb.StreamThreshold(threshold=1000, **kw)
        assert (st1.width, st1.depth, st1.confidence, st1.error_rate, st1.elements_added) == (1000, 5, 0.96875, 0.002, 0)
        assert st1.meets_threshold == {} and st1.threshold == 1000
    st1 = pb.StreamThreshold(threshold=2, width=1000, depth=5)
    assert st1.add("this is a test") == 1 and st1.meets_threshold == {}
    assert st1.add("this is a test") == 2 and st1.meets_threshold == {"this is a test": 2}
    assert st1.add("this is not a test") == 1 and st1.add("this is a test") == 3
    assert st1.add("this is not a test") == 2 and st1.add("this is still not a test") == 1
    assert st1.meets_threshold == {"this is a test": 3, "this is not a test": 2} and st1.elements_added == 6
    st1 = pb.StreamThreshold(threshold=10, width=1000, depth=5)
    got = st1.add_many(["this is a test", "this is a test", "this is not a test", "this is a test", "this is not a test"],
                       np.array([5, 5, 9, 20, 2]))
    assert got.tolist() == [5, 10, 9, 30, 11]
    assert st1.meets_threshold == {"this is a test": 30, "this is not a test": 11} and st1.elements_added == 41
    assert st1.remove("this is a test") == 29 and st1.meets_threshold == {"this is a test": 29, "this is not a test": 11}
    assert st1.remove("this is not a test") == 10 and st1.remove("this is not a test") == 9
    assert st1.meets_threshold == {"this is a test": 29} and st1.elements_added == 38
    st1 = pb.StreamThreshold(threshold=10, width=1000, depth=5)
    assert st1.add("this is a test", 30) == 30 and st1.add("this is not a test", 11) == 11 and st1.elements_added == 41
    assert st1.remove("this is not a test", 2) == 9 and st1.meets_threshold == {"this is a test": 30} and st1.elements_added == 39
    st1.clear()
    assert st1.meets_threshold == {} and st1.elements_added == 0
    st1 = pb.StreamThreshold(threshold=10, width=1000, depth=5)
    assert st1.add("this is a test", 100) == 100 and md5(bytes(st1)) == "fb1c39dd1a73f1ef0d7fc79f60fc028e"
    path = tmp_path / "s.cms"
    st1.export(path)
    st2 = pb.StreamThreshold(threshold=10, filepath=path)
    assert (st2.width, st2.depth, st2.elements_added, st2.check("this is a test"), st2.meets_threshold) == (1000, 5, 100, 100, {})
    assert st2.add("this is a test", 1) == 101 and st2.meets_threshold == {"this is a test": 101}
    st3 = pb.StreamThreshold.frombytes(bytes(st1), threshold=10)
    assert (st3.threshold, st3.elements_added, st3.check("this is a test")) == (10, 100, 100) and bytes(st3) == bytes(st1)
    assert str(st1) == ("Stream Threshold Count-Min Sketch:\n\tWidth: 1000\n\tDepth: 5\n\tConfidence: 0.96875\n\tError Rate: 0.002\n"
                        "\tElements Added: 100\n\tThreshold: 10\n\tNumber Meeting Threshold: 1")
    with pytest.raises(pb.NotSupportedError):
        st1.join(pb.StreamThreshold(threshold=1000, width=1000, depth=5))


def test_expanding_rotating_reference_test_file_io(pb, tmp_path):
    # tests/expandingbloom_test.py:111-160
    blm = pb.ExpandingBloomFilter(est_elements=25, false_positive_rate=0.05)
    blm.add_many([str(i) for i in range(105)])
    blm2 = pb.ExpandingBloomFilter.frombytes(bytes(blm))
    assert (blm2.expansions, blm2.false_positive_rate, blm2.estimated_elements, blm2.elements_added) == (3, 0.05000000074505806, 25, 105)
    assert bytes(blm2) == bytes(blm) and blm.check_many([str(i) for i in range(105)]).all()
    path = tmp_path / "e.ebf"
    pb.ExpandingBloomFilter(est_elements=25, false_positive_rate=0.05).export(path)
    assert md5(path.read_bytes()) == "eb5769ae9babdf7b37d6ce64d58812bc"
    assert [b.elements_added for b in pb.ExpandingBloomFilter(filepath=path)._blooms] == [0]
    blm = pb.ExpandingBloomFilter(est_elements=25, false_positive_rate=0.05)
    for i in range(15):
        blm.add(f"{i}")
        blm.push()
    blm.export(path)
    blm2 = pb.ExpandingBloomFilter(filepath=path)
    assert blm2.expansions == 15 and all(f"{i}" in blm2 for i in range(15)) and not any(f"{i}" in blm2 for i in range(99, 125))
    # :266-318
    rbf = pb.RotatingBloomFilter(est_elements=25, false_positive_rate=0.05, max_queue_size=3)
    rbf.add_many([str(i) for i in range(105)])
    rbf2 = pb.RotatingBloomFilter.frombytes(bytes(rbf), max_queue_size=3)
    assert (rbf2.expansions, rbf2.false_positive_rate, rbf2.estimated_elements, rbf2.elements_added) == (2, 0.05000000074505806, 25, 105)
    assert rbf2.current_queue_size == 3 and bytes(rbf2) == bytes(rbf)
    keys = [str(i) for i in range(105)]
    assert (rbf.check_many(keys) == rbf2.check_many(keys)).all()
    rpath = tmp_path / "r.rbf"
    pb.RotatingBloomFilter(est_elements=25, false_positive_rate=0.05).export(rpath)
    assert md5(rpath.read_bytes()) == "eb5769ae9babdf7b37d6ce64d58812bc"
    rbf = pb.RotatingBloomFilter(est_elements=25, false_positive_rate=0.05)
    for i in range(15):
        rbf.add(f"{i}")
        rbf.push()
    rbf.export(rpath)
    rbf2 = pb.RotatingBloomFilter(filepath=rpath)
    assert not any(f"{i}" in rbf2 for i in range(5)) and all(f"{i}" in rbf2 for i in range(6, 15))
    assert (rbf2.current_queue_size, rbf2.expansions, rbf2.elements_added) == (10, 9, 15)


def test_counting_cuckoo_many_repeats_small_table(pb, orc):
    """a long stream of repeats into a small table: the count map is reserved chunk by chunk (a quarter of the map at a
    time), never by the size of the batch"""
    rng = np.random.default_rng(9)
    pool = orc.uniform_keys(123, 150)
    stream = pool[rng.integers(0, 150, 1_500_000)]
    f = pb.CountingCuckooFilter(capacity=64, bucket_size=4, max_swaps=100, auto_expand=False)
    f.add_many(stream)
    counts = f.check_many(pool)
    uniq, freq = np.unique(stream.view(np.dtype((np.void, 16))).ravel(), return_counts=True)
    by_key = {bytes(u): int(c) for u, c in zip(uniq, freq)}
    assert [int(c) for c in counts] == [by_key.get(k.tobytes(), 0) for k in pool]
    assert f.elements_added == 1_500_000 and f.unique_elements == len(by_key) == 150
    assert f.remove_many(stream[:1000]).all() and f.elements_added == 1_499_000
