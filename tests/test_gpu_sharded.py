"""Multi-GPU parity of the range-sharded Bloom filter that also runs on a box with ONE GPU:

* `test_config5_geometry_virtual_ranks`: the exact geometry of BASELINE config 5 -- BloomFilter(1e10, 0.001):
  m = 143 775 874 672 bits, k = 10, eight shards of 34 windows of 2^29 bits -- driven through the C ABI as eight
  *virtual ranks* in one process (eight shard handles + eight mailboxes connected by pointer,
  pb_p2p_connect_local): the k = 10 partition kernel, the 64 MiB-window regime, the copy-engine exchange, the flag
  protocol and bloom_apply_sources all run exactly as they do across eight GPUs.  The 18 GB bit array is verified
  through its expected bit positions (oracle hashes): every expected bit is set and each shard's popcount equals
  the number of distinct expected positions inside it, i.e. the shards are identical to the reference's bit array.
* `test_two_ranks_share_one_gpu`: two PROCESSES on cuda:0 (gloo control plane, CUDA-IPC mailboxes): sharded.py end to
  end -- add_many in all modes incl. the skewed-batch overflow route, the index-routing check_many, Count-Min merge.
* `test_sharded_bloom_and_cms_match_oracle`: the same script over NCCL with one rank per GPU when the box has
  several GPUs (skipped on a 1-GPU box).
"""

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _run_check(world: int, extra: list[str], port: int, timeout: int = 900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "benchmarks" / "sharded_check.py")] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=str(ROOT))
    assert "SHARDED PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_sharded_bloom_and_cms_match_oracle():
    import pyprobables_b200 as pb

    n = pb.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    _run_check(world, ["--keys", "500000"], 29533)


def test_two_ranks_share_one_gpu():
    _run_check(2, ["--same-gpu", "--keys", "300000", "--chunk", "120000", "--est", "20000000"], 29534)


class _VirtualRanks:
    """R ranks of the p2p insert inside one process (C ABI only)"""

    def __init__(self, pb, num_bits, k, world, chunk_keys):
        from pyprobables_b200 import _native
        from pyprobables_b200.sharded import ShardPlan

        self.nat, self.world, self.m, self.k = _native, world, num_bits, k
        self.plan = ShardPlan.make(num_bits, world)
        self.send_ctx = [_native.Context(0) for _ in range(world)]
        self.apply_ctx = [_native.Context(0) for _ in range(world)]
        n_sub, sub_cap = C.c_uint32(), C.c_uint32()
        _native.call("pb_bloom_partition_layout", self.send_ctx[0].handle, chunk_keys, k, num_bits, self.plan.window_log2,
                     self.plan.total_windows, C.byref(n_sub), C.byref(sub_cap))
        self.shards, self.boxes = [], []
        for r in range(world):
            lo, hi = self.plan.bounds(r)
            h = C.c_void_p()
            _native.call("pb_bloom_create_shard", self.apply_ctx[r].handle, num_bits, k, lo, hi, C.byref(h))
            self.shards.append(h)
            p = C.c_void_p()
            _native.call("pb_p2p_create", self.send_ctx[r].handle, world, r, self.plan.windows_per_rank, n_sub.value, sub_cap.value,
                         C.byref(p))
            self.boxes.append(p)
        arr = (C.c_void_p * world)(*[b.value for b in self.boxes])
        for r in range(world):
            _native.call("pb_p2p_connect_local", self.boxes[r], arr)

    def insert(self, torch, keys_per_rank, chunk_keys):
        from pyprobables_b200.keys import pack_keys

        nat = self.nat
        ovf = [torch.empty(1 << 16, dtype=torch.int64, device="cuda") for _ in range(self.world)]
        ovf_n = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(self.world)]
        torch.cuda.synchronize()
        n_chunks = max(-(-int(t.shape[0]) // chunk_keys) for t in keys_per_rank)
        for ci in range(n_chunks):
            # producers of a chunk are launched before its consumers (safe with any stream-to-queue mapping)
            for r in range(self.world):
                t = keys_per_rank[r][ci * chunk_keys : (ci + 1) * chunk_keys]
                nat.call("pb_p2p_partition_send", self.boxes[r], pack_keys(t, sync=False).ref(), self.m, self.k, self.plan.window_log2,
                         C.c_void_p(ovf[r].data_ptr()), ovf[r].numel(), C.c_void_p(ovf_n[r].data_ptr()))
            for r in range(self.world):
                nat.call("pb_p2p_apply", self.boxes[r], self.shards[r], self.plan.active_windows(r), self.plan.window_log2)
        for c in self.send_ctx + self.apply_ctx:
            c.synchronize()
        for r in range(self.world):
            aborted = C.c_int(0)
            nat.call("pb_p2p_check", self.boxes[r], C.byref(aborted))
            assert aborted.value == 0, f"rank {r}: a flag wait timed out"
            assert int(ovf_n[r].item()) == 0, "uniform keys must not overflow a sublist"

    def close(self):
        for p in self.boxes:
            self.nat.lib().pb_p2p_destroy(p)
        for h in self.shards:
            self.nat.lib().pb_bloom_destroy(h)


@pytest.mark.parametrize("est,fpr,world,n_per_rank,chunk", [
    (10**10, 0.001, 8, 900_000, 400_000),   # BASELINE config 5 exactly: m = 143 775 874 672, k = 10, 8 shards
    (3 * 10**9, 0.01, 3, 500_000, 200_000),  # k = 7, short last shard, 512-key tiles (216 global windows)
])
def test_config5_geometry_virtual_ranks(orc, golden, est, fpr, world, n_per_rank, chunk):
    import torch

    import pyprobables_b200 as pb
    from pyprobables_b200 import _native
    from pyprobables_b200.bloom import optimized_params

    assert pb.device_count() >= 1
    _, k, m = optimized_params(est, fpr)
    if est == 10**10:
        kat = golden["bloom_index_kat"]["cfg5"]
        assert (m, k) == (kat["m"], kat["k"]) == (143775874672, 10)
    vr = _VirtualRanks(pb, m, k, world, chunk)
    try:
        host = [orc.uniform_keys(r * n_per_rank, n_per_rank + 777 * r) for r in range(world)]  # ragged on purpose
        vr.insert(torch, [torch.from_numpy(h).cuda() for h in host], chunk)
        hashes = np.concatenate([orc.default_fnv_1a_many(orc.pack(h), k).reshape(-1) for h in host])
        pos = np.unique(hashes % np.uint64(m))
        if est == 10**10:  # key 0's golden bit positions (generated by the pure-Python reference) are among them
            assert np.isin(np.array(kat["bits"][0], dtype=np.uint64), pos).all()
        total_pc = 0
        for r in range(world):
            lo, hi = vr.plan.bounds(r)
            mine = torch.from_numpy(pos[(pos >= lo) & (pos < hi)].astype(np.int64)).cuda()
            out = torch.zeros(mine.numel(), dtype=torch.uint8, device="cuda")
            _native.call("pb_bloom_test_bit_indices", vr.shards[r], C.c_void_p(mine.data_ptr()), mine.numel(), C.c_void_p(out.data_ptr()))
            pc = C.c_uint64()
            _native.call("pb_bloom_popcount", vr.shards[r], C.byref(pc))
            assert bool(out.all().item()), f"shard {r}: an expected bit is missing"
            assert pc.value == mine.numel(), f"shard {r}: {pc.value} bits set, {mine.numel()} expected"
            total_pc += pc.value
        assert total_pc == pos.size
    finally:
        vr.close()
