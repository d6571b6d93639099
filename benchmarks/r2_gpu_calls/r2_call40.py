"""kernel-time breakdown of ExpandingBloomFilter.add_many (1e8 keys, filters of 1.25e7)"""
import sys, time
sys.path.insert(0, ".")
import torch
import pyprobables_b200 as pb

n = 100_000_000
stream = torch.cuda.Stream()
ctx = pb.Context(0, stream=stream.cuda_stream)
with torch.cuda.stream(stream):
    keys = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
    ctx.gen_uniform_keys(0, n, keys.data_ptr())
    f = pb.ExpandingBloomFilter(est_elements=n // 8, false_positive_rate=0.01, context=ctx)
    f.add_many(keys[: 1 << 16])
    f.close()
    f = pb.ExpandingBloomFilter(est_elements=n // 8, false_positive_rate=0.01, context=ctx)
    ctx.set_option("kernel_timing", 1)
    ctx.synchronize()
    t0 = time.perf_counter()
    f.add_many(keys)
    ctx.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    kt = ctx.kernel_times()
    print("wall ms", wall)
    tot = 0
    for k, v in sorted(kt.items(), key=lambda kv: -kv[1]["total_ms"] if isinstance(kv[1], dict) else 0):
        print(k, v)
        tot += v["total_ms"] if isinstance(v, dict) else 0
    print("sum of timed kernels ms", tot)
