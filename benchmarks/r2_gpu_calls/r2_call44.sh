#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_counting.py -q -m gpu -x > gpurun_out/c44_pytest.log 2>&1
tail -25 gpurun_out/c44_pytest.log
timeout 300 python - <<PY 2>&1 | tail -6
import sys, time
sys.path.insert(0, ".")
import torch
import pyprobables_b200 as pb
n = 125_000_000
stream = torch.cuda.Stream()
ctx = pb.Context(0, stream=stream.cuda_stream)
with torch.cuda.stream(stream):
    keys = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
    ctx.gen_uniform_keys(0, n, keys.data_ptr())
    for mode in (1, 0):
        ctx.set_option("bloom_insert_mode", mode)
        f = pb.CountingBloomFilter(2 * 10**8, 0.01, context=ctx)
        f.add_many(keys[: 1 << 20])
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        f.add_many(keys)
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1)
        import hashlib
        print("mode", mode, "counting bloom add", round(n / ms / 1e6, 2), "G keys/s", round(ms, 2), "ms", "md5", hashlib.md5(f.bloom_numpy().tobytes()).hexdigest()[:12])
        f.close()
PY
