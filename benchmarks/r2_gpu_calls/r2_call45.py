"""Counting Bloom partitioned add: window size sweep with per-kernel times (125 M keys into 1.9e9 counters)"""
import sys
sys.path.insert(0, ".")
import torch
import pyprobables_b200 as pb

n = 125_000_000
stream = torch.cuda.Stream()
ctx = pb.Context(0, stream=stream.cuda_stream)
with torch.cuda.stream(stream):
    keys = torch.empty((n, 16), dtype=torch.uint8, device="cuda")
    ctx.gen_uniform_keys(0, n, keys.data_ptr())
    ctx.set_option("kernel_timing", 1)
    for mode, bits, tile in ((1, 27, 0), (2, 27, 0), (2, 28, 0), (2, 29, 0), (2, 28, 256), (2, 29, 256)):
        ctx.set_option("bloom_insert_mode", mode)
        ctx.set_option("bloom_window_log2_bits", bits)
        ctx.set_option("bloom_part_tile", tile)
        f = pb.CountingBloomFilter(2 * 10**8, 0.01, context=ctx)
        f.add_many(keys[: 1 << 20])
        ctx.synchronize()
        ctx.kernel_times()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        f.add_many(keys)
        e1.record(stream)
        stream.synchronize()
        print("mode", mode, "window bits", bits, "tile", tile, round(n / e0.elapsed_time(e1) / 1e6, 2), "G keys/s", round(e0.elapsed_time(e1), 2), "ms", ctx.kernel_times())
        f.close()
