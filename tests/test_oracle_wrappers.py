"""CPU: the oracle's one-key-at-a-time restatement of the stateful wrappers against vectors recorded from the reference
(tests/golden/make_golden_r2_wrappers.py)."""

import hashlib

import numpy as np
import pytest


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


def _bare_heavy_hitters(num_hitters):
    """a HeavyHitters object without its device sketch: only the dictionary bookkeeping is exercised (CPU)"""
    from pyprobables_b200.countminsketch import HeavyHitters

    hh = object.__new__(HeavyHitters)
    hh._top_x, hh._top_x_size, hh._num_hitters, hh._smallest = {}, 0, num_hitters, 0
    return hh


def test_heavy_hitters_bookkeeping_replays_the_reference(orc, golden):
    """countminsketch.py:644-660 through HeavyHitters._track, fed with the reference's own return values"""
    h = golden["heavy_hitters"]
    names = [bytes(k).hex() for k in orc.rank_keys(np.array(h["ranks"], dtype=np.uint64))]
    o = orc.HeavyHitters(h["num_hitters"], h["width"], h["depth"])
    rets = o.add_tracked(names, orc.pack(names))
    hh = _bare_heavy_hitters(h["num_hitters"])
    for key, res in zip(names, rets.tolist()):
        hh._track(key, res)
    assert hh.heavy_hitters == h["heavy_hitters"] and list(hh.heavy_hitters) == list(o.top_x)  # same insertion order too


@pytest.mark.parametrize("seed,num_hitters,n_keys,zipf", [(1, 5, 40, 1.3), (2, 20, 500, 1.1), (3, 1, 10, 2.0), (4, 50, 60, 1.05), (5, 8, 3000, 1.2)])
def test_heavy_hitters_bulk_replay_equals_one_by_one(seed, num_hitters, n_keys, zipf):
    """HeavyHitters._replay_rows (tracked keys updated in bulk between the pairs that change WHICH keys are tracked)
    against _track pair by pair, on random streams whose per-key values rise like a sketch's return values do; several
    runs back to back so that the dictionary, its stale `smallest` and the candidate filter carry over"""
    import torch

    rng = np.random.default_rng(seed)
    pool = rng.integers(0, 256, (n_keys, 12), dtype=np.uint8)
    counts = np.zeros(n_keys, dtype=np.int64)
    bulk, single = _bare_heavy_hitters(num_hitters), _bare_heavy_hitters(num_hitters)
    for run in range(6):
        m = int(rng.integers(1, 20_000))
        ids = np.minimum(rng.zipf(zipf, m) - 1, n_keys - 1)
        vals = np.empty(m, dtype=np.int64)
        noise = rng.integers(0, 3, m)  # collisions only ever add to an estimate
        for j, i in enumerate(ids):
            counts[i] += 1
            vals[j] = counts[i] + noise[j] * (counts[i] > 3)
            counts[i] = vals[j]  # keep each key's values non-decreasing, as add()'s return values are
        keep = vals >= bulk._smallest  # the candidate filter add_many applies per slice
        rows, v = pool[ids[keep]], vals[keep]
        if rows.shape[0]:
            bulk._replay_rows(torch.from_numpy(rows), torch.from_numpy(v))
        for i, x in zip(ids.tolist(), vals.tolist()):
            single._track(pool[i].tobytes(), x)
        assert bulk.heavy_hitters == single.heavy_hitters and list(bulk.heavy_hitters) == list(single.heavy_hitters)
        assert (bulk._smallest, bulk._top_x_size) == (single._smallest, single._top_x_size)


def test_counting_cuckoo_removal_flags_go_to_the_first_occurrences():
    """countingcuckoo._first_occurrences_win: whatever occurrences the device let succeed, the flags end up where the
    one-key-at-a-time loop puts them -- on the first `count` occurrences of each fingerprint"""
    from pyprobables_b200.countingcuckoo import _first_occurrences_win

    rng = np.random.default_rng(4)
    for _ in range(50):
        n = int(rng.integers(2, 400))
        fps = rng.integers(1, 12, n).astype(np.uint32)
        stored = {int(f): int(rng.integers(0, 6)) for f in np.unique(fps)}  # count of each fingerprint before the batch
        left, sequential = dict(stored), np.zeros(n, dtype=bool)
        for i, f in enumerate(fps.tolist()):  # the reference's loop
            if left[f] > 0:
                left[f] -= 1
                sequential[i] = True
        device = np.zeros(n, dtype=bool)  # the same number of successes per fingerprint on arbitrary occurrences
        for f, c in stored.items():
            where = np.flatnonzero(fps == f)
            device[rng.permutation(where)[: min(c, where.size)]] = True
        assert (_first_occurrences_win(device, fps) == sequential).all()
