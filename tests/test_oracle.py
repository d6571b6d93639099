"""Pins the CPU oracle (oracle/pb_oracle.c) to the reference: literals from the reference's own tests
(cited per test) and fixtures generated from the pure-Python reference (tests/golden/golden.json)."""

import hashlib
import struct

import numpy as np


def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()


# ---------------------------------------------------------------- hashes
def test_fnv_reference_kats(orc):
    # /root/reference/tests/hashes_test.py:27-46 (str) and :148-167 (bytes)
    exp1 = [4040040117721899264, 3916497180155386777, 468410530588793106, 13781401791305604595, 321382271269641900]
    assert orc.default_fnv_1a("this is a test", 5) == exp1
    assert orc.default_fnv_1a(b"this is a test", 5) == exp1
    # hashes_test.py:48-55: equal at seed 0, different at seeds 1..4
    a = orc.default_fnv_1a("gMPflVXtwGDXbIhP73TX", 5)
    b = orc.default_fnv_1a("LtHf1prlU1bCeYZEdqWf", 5)
    assert a[0] == b[0] and all(x != y for x, y in zip(a[1:], b[1:]))


def test_fnv_golden(orc, golden):
    for s, exp in golden["fnv_str"].items():
        assert orc.default_fnv_1a(s, 4) == exp, s
    for hx, exp in golden["fnv_bytes"].items():
        assert orc.default_fnv_1a(bytes.fromhex(hx), 4) == exp
    for seed, exp in golden["fnv_seed_big"].items():
        assert orc.fnv_1a("seed test", int(seed)) == exp
    # str is hashed per code point, not per UTF-8 byte
    assert orc.fnv_1a("é") == orc.fnv_1a("é".encode("latin1")) != orc.fnv_1a("é".encode("utf-8"))


def test_key_generators(orc, golden):
    for i, hx in golden["keys"].items():
        assert orc.uniform_keys(int(i), 1).tobytes().hex() == hx
    assert orc.default_fnv_1a(bytes.fromhex(golden["keys"]["0"]), 7) == golden["fnv_key0_k7"]
    keys = orc.pack(orc.uniform_keys(0, 8))
    for name, kat in golden["bloom_index_kat"].items():
        h = orc.default_fnv_1a_many(keys, kat["k"])
        assert (h % np.uint64(kat["m"])).tolist() == kat["bits"], name
    assert orc.rank_keys(np.array([golden["cms_zipf"]["ranks_first5"][0]])).tobytes().hex() == golden["cms_zipf"]["key0"]


# ---------------------------------------------------------------- bloom
def test_bloom_sizing(orc, golden):
    # bloom_test.py:27-36 and :134-140
    assert orc.bloom_params(10, 0.05) == (0.05000000074505806, 4, 63, 8)
    assert orc.bloom_params(16_000_000, 0.001)[2:] == (230041400, 28755175)
    for key, (fpr, k, m) in golden["bloom_sizing"].items():
        e, f = key.split("/")
        assert orc.bloom_params(int(e), float(f))[:3] == (fpr, k, m)


def test_bloom_ten_keys_hex(orc, golden):
    # bloom_test.py:256-265: export_hex of 10 keys
    fpr, k, m, _ = orc.bloom_params(10, 0.05)
    b = orc.Bloom(m, k)
    b.add(orc.pack([f"this is a test {i}" for i in range(10)]))
    hx = b.bloom.tobytes().hex() + struct.pack(">QQf", 10, 10, fpr).hex()
    assert hx == "6da491461a6bba4d000000000000000a000000000000000a3d4ccccd" == golden["bloom_10"]["export_hex"]
    # bloom_test.py:323-341: one key, export md5
    b = orc.Bloom(m, k)
    b.add(orc.pack(["this is a test"]))
    assert md5(b.bloom.tobytes() + struct.pack("QQf", 10, 1, fpr)) == "8d27e30e1c5875b0edcf7413c7bdb221"


def test_bloom_check_reference(orc):
    # bloom_test.py:56-74
    fpr, k, m, _ = orc.bloom_params(10, 0.05)
    b = orc.Bloom(m, k)
    b.add(orc.pack(["this is a test", "this is another test"]))
    got = b.check(orc.pack(["this is a test", "this is another test", "this is yet another test", "this is not another test"]))
    assert got.tolist() == [True, True, False, False]


def test_bloom_variable_and_unicode(orc, golden):
    g = golden["bloom_var"]
    _, k, m, _ = orc.bloom_params(g["est"], g["fpr"])
    b = orc.Bloom(m, k)
    b.add(orc.pack([bytes.fromhex(x) for x in g["keys"]]))
    assert b.bloom.tobytes().hex() == g["bitmap_hex"]
    assert b.popcount() == g["bits_set"]
    assert b.check(orc.pack([bytes.fromhex(x) for x in g["probes"]])).tolist() == g["check_probes"]
    g = golden["bloom_unicode"]
    _, k, m, _ = orc.bloom_params(g["est"], g["fpr"])
    b = orc.Bloom(m, k)
    b.add(orc.pack(g["keys"]))
    assert b.bloom.tobytes().hex() == g["bitmap_hex"]
    assert b.check(orc.pack([s.encode("utf-8") for s in g["keys"]])).tolist() == g["check_utf8_bytes"]


def test_bloom_config1_full_state(orc, golden):
    """BASELINE config 1: 1 M uniform keys; the whole 1.2 MB bitmap must equal the reference's."""
    g = golden["config1"]
    _, k, m, length = orc.bloom_params(10**6, 0.01)
    assert (m, k, length) == (g["num_bits"], g["k"], g["bloom_length"])
    b = orc.Bloom(m, k)
    b.add(orc.pack(orc.uniform_keys(0, 10**6)))
    assert md5(b.bloom) == g["bitmap_md5"]
    assert b.popcount() == g["bits_set"]
    assert md5(b.bloom.tobytes() + struct.pack("QQf", 10**6, 10**6, np.float32(0.01))) == g["export_md5"]
    assert b.check(orc.pack(orc.uniform_keys(0, 10**6))).all()
    fp = np.nonzero(b.check(orc.pack(orc.uniform_keys(10**6, 10**6))))[0] + 10**6
    assert len(fp) == g["false_positives"]
    assert md5(fp.astype("<u8").tobytes()) == g["false_positive_idx_md5"]
    # pre-hashed path (add_alt / check_alt)
    b2 = orc.Bloom(m, k)
    keys = orc.pack(orc.uniform_keys(0, 50_000))
    b2.add_hashes(orc.default_fnv_1a_many(keys, k))
    b3 = orc.Bloom(m, k)
    b3.add(keys)
    assert (b2.bloom == b3.bloom).all()
    assert b2.check_hashes(orc.default_fnv_1a_many(keys, k)).all()


# ---------------------------------------------------------------- count-min
def test_cms_reference_kats(orc):
    # countminsketch_test.py:76-92
    c = orc.CMS(1000, 5)
    assert c.add(orc.pack(["this is a test"] * 4), 1, want_returns=True).tolist() == [1, 2, 3, 4]
    assert c.elements_added == 4
    c = orc.CMS(1000, 5)
    assert c.add(orc.pack(["this is a test"] * 4), 4, want_returns=True).tolist() == [4, 8, 12, 16]
    # countminsketch_test.py:187-203: export md5
    c = orc.CMS(1000, 5)
    c.add(orc.pack(["this is a test"]), 100)
    assert md5(c.bins.tobytes() + struct.pack("IIq", 1000, 5, 100)) == "fb1c39dd1a73f1ef0d7fc79f60fc028e"
    # countminsketch_test.py:270-278: saturation (num_els clamped to INT64_MAX by the caller)
    c = orc.CMS(1000, 5)
    c.add(orc.pack(["this is a test"]), 2**63 - 1)
    assert c.check(orc.pack(["this is a test"])).tolist() == [2**31 - 1]
    assert c.elements_added == 2**63 - 1


def test_cms_query_types_reference(orc):
    # countminsketch_test.py:111-185: 255/189/16/5 weighted adds, the three query types
    for qt, exp in (("min", [255, 189, 16, 5]), ("mean", None), ("mean-min", None)):
        c = orc.CMS(1000, 5, qt)
        names = ["this is a test", "this is another test", "this is also a test", "this is something to test"]
        for nm, n in zip(names, (255, 189, 16, 5)):
            c.add(orc.pack([nm]), n)
        got = c.check(orc.pack(names)).tolist()
        if exp:
            assert got == exp
        assert c.elements_added == 255 + 189 + 16 + 5


def test_cms_small_golden(orc, golden):
    for name, g in golden["cms_small"].items():
        depth = int(name[1])
        qt = name.split("-", 1)[1]
        c = orc.CMS(97, depth, qt)
        keys = orc.pack([k for k, _ in g["seq"]])
        rets = c.add(keys, np.array([n for _, n in g["seq"]], dtype=np.int64), want_returns=True)
        assert rets.tolist() == g["returns"], name
        assert c.bins.tolist() == g["bins"]
        assert c.elements_added == g["elements_added"]
        assert c.check(orc.pack([f"key-{i}" for i in range(50)])).tolist() == g["checks"]
    g = golden["cms_saturation"]
    c = orc.CMS(1000, 5)
    r = c.add(orc.pack(["this is a test", "this is a test", "other"]), np.array([2**31 - 10, 100, 7]), want_returns=True)
    assert r.tolist() == g["returns"]
    assert md5(c.bins) == g["bins_md5"] and c.elements_added == g["elements_added"]


def test_cms_zipf_golden(orc, golden):
    g = golden["cms_zipf"]
    ranks = np.random.default_rng(0xB200).zipf(1.1, g["n"]).astype(np.int64)
    assert md5(ranks.astype("<i8").tobytes()) == g["ranks_md5"], "numpy zipf stream changed: regenerate golden"
    keys = orc.pack(orc.rank_keys(ranks.astype(np.uint64)))
    c = orc.CMS(g["width"], g["depth"])
    c.add(keys)
    assert md5(c.bins) == g["bins_md5"]
    assert md5(c.bins.tobytes() + struct.pack("IIq", g["width"], g["depth"], g["n"])) == g["export_md5"]
    assert int(np.count_nonzero(c.bins)) == g["nonzero_bins"]
    probe = orc.pack(orc.rank_keys(np.arange(1, 1001, dtype=np.uint64)))
    for qt, exp in g["estimates_1_1000"].items():
        c.query_type = qt
        assert c.check(probe).tolist() == exp, qt
    # the order-free parallel add used by the CPU baseline gives the same table
    c2 = orc.CMS(g["width"], g["depth"])
    c2.add_parallel(keys)
    assert (c2.bins == c.bins).all()


# ---------------------------------------------------------------- cuckoo
def test_cuckoo_export_kats(orc, golden):
    # cuckoo_test.py:248-266: 1000 keys str(i) into the default filter (no eviction happens)
    c = orc.Cuckoo(10000, 4, 500, 32)
    c.add(orc.pack([str(i) for i in range(1000)]))
    assert md5(c.export_slots().tobytes() + struct.pack("II", 4, 500)) == "1371760d4ee9ccbe83e0144919750140"
    # cuckoo_test.py:489-498: error-rate constructor -> 20-bit fingerprints
    g = golden["cuckoo_1000_err"]
    c = orc.Cuckoo(g["capacity"], 4, 500, g["fp_bits"])
    c.add(orc.pack([str(i) for i in range(1000)]))
    assert md5(c.export_slots().tobytes() + struct.pack("II", 4, 500)) == "3c693508d1a3acd819310fd0c11dc906"
    assert c.elements_added == g["elements_added"]


def test_cuckoo_info_and_dedupe(orc, golden):
    keys = orc.pack(orc.uniform_keys(0, 8))
    for cap, exp in golden["cuckoo_info"].items():
        i1, i2, fp = orc.Cuckoo(16, 4, 5, 32).fingerprint_info(keys)  # index math only
        c = orc.Cuckoo.__new__(orc.Cuckoo)
        c.capacity, c.fp_bits, c.h = int(cap), 32, None
        i1, i2, fp = orc.Cuckoo.fingerprint_info(c, keys)
        assert [[int(a), int(b), int(f)] for a, b, f in zip(i1, i2, fp)] == exp
    # cuckoo_test.py:221-231: re-adding does not grow elements_added
    c = orc.Cuckoo(100, 2, 5, 32)
    c.add(orc.pack(["this is a test", "this is another test", "this is yet another test"] * 2))
    assert c.elements_added == 3


def test_cuckoo_95_load_membership(orc, golden):
    g = golden["cuckoo_95"]
    c = orc.Cuckoo(g["capacity"], 4, 500, 32, rng_seed=7)
    keys = orc.uniform_keys(0, g["keys_consumed"])
    failed = c.add(orc.pack(keys))
    assert len(failed) == 0
    assert c.elements_added == g["elements_added"]
    assert md5(c.fingerprints().astype("<u4").tobytes()) == g["sorted_fp_md5"]
    assert c.check(orc.pack(keys)).all()
    pos = np.nonzero(c.check(orc.pack(orc.uniform_keys(10_000_000, 1_000_000))))[0] + 10_000_000
    assert pos.tolist() == g["probe_positive_idx"]
    # placement differs with the RNG seed, membership does not
    c2 = orc.Cuckoo(g["capacity"], 4, 500, 32, rng_seed=99)
    c2.add(orc.pack(keys))
    assert (c2.fingerprints() == c.fingerprints()).all()


def test_cuckoo_full(orc, golden):
    # cuckoo_test.py:126-135 behaviour: a full filter reports the homeless fingerprint
    c = orc.Cuckoo(100, 2, 100, 32)
    failed = c.add(orc.pack(orc.uniform_keys(0, 400)))
    assert c.n_failed > 0 and golden["cuckoo_full"]["type"] == "CuckooFilterFullError"


# ---------------------------------------------------------------- counting bloom (round 2 golden: make_golden_r2.py)
def test_counting_bloom_sequences_match_reference(orc, golden):
    """every return value and the final counters of a 399-step add / remove / check sequence recorded from the
    reference, on a table small enough that hashes of one key collide (the double-increment quirk)"""
    g = golden["cbloom_seq"]
    assert g["keys_with_colliding_hashes"] > 0
    o = orc.CountingBloom(g["num_bits"], g["k"])
    got = []
    for op, key, n in g["ops"]:
        kb = orc.pack([key])
        if op == "add":
            got.append(int(o.add(kb, n)[0]))
        elif op == "remove":
            got.append(int(o.remove(kb, n)[0]))
        else:
            got.append(int(o.check(kb)[0]))
    assert got == g["returns"]
    assert o.bloom.tolist() == g["counters"] and o.elements_added == g["elements_added"]
    gb = golden["cbloom_batch"]
    o = orc.CountingBloom(gb["num_bits"], gb["k"])
    o.add(orc.pack(orc.uniform_keys(0, gb["n_add"])))
    assert hashlib.md5(o.bloom.tobytes()).hexdigest() == gb["counters_md5_after_add"]
    o.remove(orc.pack(orc.uniform_keys(0, gb["n_remove"])))
    assert hashlib.md5(o.bloom.tobytes()).hexdigest() == gb["counters_md5"] and o.elements_added == gb["elements_added"]
    gc = golden["cbloom_clamped"]
    from pyprobables_b200.bloom import optimized_params

    _, k, m = optimized_params(gc["est"], gc["fpr"])
    o = orc.CountingBloom(m, k)
    o.add(orc.pack(orc.uniform_keys(0, 500)))
    rets = o.remove(orc.pack(np.concatenate([orc.uniform_keys(0, 500)] * 3)), 2)
    assert rets[:20].tolist() == gc["returns_head"]
    assert hashlib.md5(struct.pack("<1500Q", *[int(x) for x in rets])).hexdigest() == gc["returns_md5"]
    assert o.bloom.tolist() == gc["counters"] and o.elements_added == gc["elements_added"]
