"""Round-2 API rows on the device against tests/golden/golden_r2.json (recorded from the pure-Python reference):
CuckooFilter.remove (cuckoo.py:317-330), CuckooFilter with a non-default hash_function (idx_2 from the user's function,
cuckoo.py:489), fnv_1a with any seed / fnv_1a_32 (hashes.py:86-122), BloomFilter.export_c_header (bloom.py:306-322)."""

import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import pyprobables_b200 as p

    assert p.device_count() >= 1
    return p


def md5_int_hash(key) -> int:
    """the SimpleHashT used by make_golden_r2.py: first 8 bytes of md5, big-endian"""
    if isinstance(key, str):
        key = key.encode("utf-8")
    return int.from_bytes(hashlib.md5(key).digest()[:8], "big")


def test_cuckoo_remove(pb, orc, golden):
    g = golden["cuckoo_remove"]
    ck = pb.CuckooFilter(capacity=200, bucket_size=2, max_swaps=50)
    ck.add_many([f"w{i}" for i in range(150)])
    assert ck.elements_added == 150
    assert [ck.remove(f"w{i}") for i in range(0, 300, 3)] == g["removed"]
    assert ck.elements_added == g["elements_added"]
    assert ck.check_many([f"w{i}" for i in range(150)]).tolist() == g["present_after"]
    assert ck.fingerprints().tolist() == g["fingerprints"]
    # the same as one batch, with a repeated key: only its first occurrence reports True
    ck2 = pb.CuckooFilter(capacity=200, bucket_size=2, max_swaps=50)
    ck2.add_many([f"w{i}" for i in range(150)])
    batch = [f"w{i}" for i in range(0, 300, 3)] + ["w0", "w3"]
    assert ck2.remove_many(batch).tolist() == g["removed"] + [False, False]
    assert ck2.elements_added == g["elements_added"] and ck2.fingerprints().tolist() == g["fingerprints"]
    # removed keys can be added again; large device-side batch against the oracle's fingerprint function
    keys = orc.uniform_keys(0, 400_000)
    ck3 = pb.CuckooFilter(capacity=200_000, bucket_size=4, auto_expand=False)
    ck3.add_many(keys)
    n0 = ck3.elements_added
    gone = ck3.remove_many(keys[:150_000])
    assert ck3.elements_added == n0 - int(gone.sum())
    fp_all = orc.Cuckoo(16, 4, 5, 32).fingerprint_info(orc.pack(keys))[2]
    kept = np.unique(fp_all[150_000:])
    expect = np.setdiff1d(np.unique(fp_all), np.unique(fp_all[:150_000]))
    assert (ck3.fingerprints() == expect).all()  # every fingerprint of a removed key is gone (even when a kept key shares it)
    assert not ck3.check_many(keys[:150_000]).any()
    ck3.add_many(keys[:150_000])
    assert ck3.check_many(keys).all() and kept.size <= ck3.elements_added


def test_cuckoo_custom_hash_matches_reference(pb, golden):
    g = golden["cuckoo_md5"]
    words = [f"md5-key-{i}" for i in range(900)]
    probes = [f"probe-{i}" for i in range(2000)]
    # a file written by the reference with a non-default hash loads and answers identically
    ck = pb.CuckooFilter.frombytes(bytes.fromhex(g["export_hex"]), hash_function=md5_int_hash)
    ck.fingerprint_size = 2
    assert ck.capacity == g["capacity"] and ck.elements_added == g["elements_added"]
    assert ck.check_many(words).all()
    assert np.nonzero(ck.check_many(probes))[0].tolist() == g["probe_hits"]
    assert [list(ck._generate_fingerprint_info(w)) for w in words[:10]] == g["info_head"]
    assert bytes(ck).hex() == g["export_hex"]
    assert ck.remove(words[5]) and not ck.check(words[5]) and ck.elements_added == g["elements_added"] - 1
    # built here from scratch with the same function: same fingerprint set, same membership
    ck2 = pb.CuckooFilter(capacity=300, bucket_size=4, max_swaps=100, finger_size=2, hash_function=md5_int_hash)
    ck2.add_many(words)
    assert ck2.elements_added == g["elements_added"] and ck2.fingerprints().tolist() == g["fingerprints"]
    assert ck2.check_many(words).all()
    assert np.nonzero(ck2.check_many(probes))[0].tolist() == g["probe_hits"]
    # growth with a custom hash (cuckoo.py:455-481 through the user's function)
    ck3 = pb.CuckooFilter(capacity=20, bucket_size=2, max_swaps=20, hash_function=md5_int_hash)
    ck3.add_many(words[:300])
    assert ck3.capacity > 20 and ck3.check_many(words[:300]).all()
    assert ck3.elements_added == len({md5_int_hash(w) & 0xFFFFFFFF for w in words[:300]})


def test_small_cuckoo_filters_do_not_allocate_the_claim_bitmap(pb):
    import torch

    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    filters = [pb.CuckooFilter() for _ in range(6)]
    for i, f in enumerate(filters):
        f.add(f"key {i}")
        f.add_many([f"k{j}" for j in range(1000)])
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < (128 << 20), "six default filters must not cost 6 x 512 MiB of claim bitmaps"
    assert all(f.check("k7") for f in filters)


def test_fnv_seeds_and_32bit(pb, golden):
    for seed, want in golden["fnv_1a_seeds"].items():
        assert pb.hashes.fnv_1a("this is a test", int(seed)) == want
    assert pb.hashes.fnv_1a("this is a test", -3) == pb.hashes.fnv_1a("this is a test", 2**64 - 3)

    def ref32(key, seed=0):  # hashes.py:106-122
        h = (0x811C9DC5 + 31 * seed) & 0xFFFFFFFF
        for c in (list(key) if not isinstance(key, str) else list(map(ord, key))):
            h = ((h ^ c) * 0x01000193) & 0xFFFFFFFF
        return h

    for key in ("", "a", "this is a test", b"\x00\xff bytes", "naïve café", "日本語"):
        for seed in (0, 1, 99, 2**31):
            assert pb.hashes.fnv_1a_32(key, seed) == ref32(key, seed), (key, seed)


def test_export_c_header(pb, golden, tmp_path):
    b = pb.BloomFilter(est_elements=10, false_positive_rate=0.05)
    for i in range(10):
        b.add(f"this is a test {i}")
    p = tmp_path / "b.h"
    b.export_c_header(p)
    assert p.read_text() == golden["bloom_c_header"]
