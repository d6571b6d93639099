"""CPU: the oracle's one-key-at-a-time restatement of the stateful wrappers against vectors recorded from the reference
(tests/golden/make_golden_r2_wrappers.py)."""

import hashlib

import numpy as np


def md5(b) -> str:
    return hashlib.md5(bytes(b)).hexdigest()


def stack_state(o) -> dict:
    return {"n_blooms": len(o.blooms), "per_bloom_added": [b.elements_added for b in o.blooms],
            "per_bloom_md5": [md5(b.bloom.tobytes()) for b in o.blooms], "elements_added": o.elements_added,
            "export_md5": md5(o.export())}


def test_expanding_reference_literals(orc, golden):
    o = orc.ExpandingBloom(10, 0.05)
    o.add(orc.pack([f"{i}" for i in range(120)]))
    assert stack_state(o) == golden["ebf_120_no_force"] and len(o.blooms) - 1 == 8  # tests/expandingbloom_test.py:47-54
    o = orc.ExpandingBloom(10, 0.05)
    o.add(orc.pack([f"{i}" for i in range(100)]), force=True)
    assert stack_state(o) == golden["ebf_100_force"] and len(o.blooms) - 1 == 9  # :33-38
    assert md5(orc.ExpandingBloom(25, 0.05).export()) == "eb5769ae9babdf7b37d6ce64d58812bc"  # :99-109


def test_expanding_stream(orc, golden):
    g = golden["ebf_stream"]
    keys = np.concatenate([orc.uniform_keys(0, g["n_unique"]), orc.uniform_keys(0, g["n_repeat"])])
    o = orc.ExpandingBloom(g["est"], g["fpr"])
    for q in range(4):
        o.add(orc.pack(keys[q * 6250 : (q + 1) * 6250]))
        assert stack_state(o) == g["quarters"][q]
    hits = np.flatnonzero(o.check(orc.pack(orc.uniform_keys(100000, 2000)))) + 100000
    assert hits.tolist() == g["probe_hits_100000_102000"]


def test_rotating_stream(orc, golden):
    g = golden["rbf_stream"]
    o = orc.ExpandingBloom(g["est"], g["fpr"], max_queue_size=g["queue"])
    for h in range(2):
        o.add(orc.pack(orc.uniform_keys(h * 3000, 3000)))
        assert stack_state(o) == g["halves"][h]
    present = np.flatnonzero(o.check(orc.pack(orc.uniform_keys(0, 6000)[::10]))) * 10
    assert present.tolist() == g["present_step10"]
    o.push()
    assert stack_state(o) == g["after_push"]
    o.blooms.pop(0)  # RotatingBloomFilter.pop, expandingbloom.py:332-341
    assert stack_state(o) == g["after_pop"]
    o.add(orc.pack(orc.uniform_keys(6000, 500)), force=True)
    assert stack_state(o) == g["after_force_500"]


def test_counting_cuckoo_vs_reference(orc, golden):
    g = golden["ccf"]
    keys = orc.uniform_keys(0, 1300)
    o = orc.CountingCuckoo(g["capacity"], g["bucket_size"], g["max_swaps"])
    assert o.add(orc.pack(keys[g["draws"]])) == []
    assert o.bins() == [tuple(x) for x in g["bins"]]  # whatever victims the eviction walk picked
    assert (o.elements_added, o.unique_elements) == (g["elements_added"], g["unique_elements"])
    assert o.check(orc.pack(keys)).tolist() == g["check_0_1300"]
    assert o.remove(orc.pack(keys[0:1300:3])).tolist() == g["removed_step3"]
    assert o.bins() == [tuple(x) for x in g["bins_after_remove"]]
    assert (o.elements_added, o.unique_elements) == (g["elements_added_after_remove"], g["unique_after_remove"])
    e = golden["ccf_export"]
    o = orc.CountingCuckoo(e["capacity"], 4, 5)
    o.add(orc.pack([str(i % 400) for i in range(600)]))
    assert md5(o.export()) == e["export_md5"] and len(o.export()) == e["export_len"]


def test_heavy_hitters_and_stream_threshold_vs_reference(orc, golden):
    import struct

    h = golden["heavy_hitters"]
    names = [bytes(k).hex() for k in orc.rank_keys(np.array(h["ranks"], dtype=np.uint64))]
    o = orc.HeavyHitters(h["num_hitters"], h["width"], h["depth"])
    rets = o.add_tracked(names, orc.pack(names))
    assert md5(struct.pack(f"<{len(rets)}q", *rets.tolist())) == h["returns_md5"]
    assert o.top_x == h["heavy_hitters"] and md5(o.bins.tobytes()) == h["bins_md5"]
    t = golden["stream_threshold"]
    o = orc.StreamThreshold(t["threshold"], t["width"], t["depth"])
    rets = o.add_tracked(names, orc.pack(names))
    assert md5(struct.pack(f"<{len(rets)}q", *rets.tolist())) == t["returns_md5"] and o.meets == t["meets_threshold"]
