// pb_ctx.cu -- context lifetime, buffers, the host-batch copy/compute pipeline, key generators,
// batch hashing (hashes.py:71-83) and the host-callable self-test hooks.
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <new>

#include "pb_common.cuh"
#include "pb_hash.cuh"
#include "pb_keys.cuh"

namespace pb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int scratch_reserve(pb_ctx *ctx, pb_scratch &s, size_t bytes) {
    (void)ctx;
    if (bytes <= s.cap) return PB_OK;
    if (s.p) {
        // buffers may still be in flight on the streams of this context
        PB_CUDA(cudaStreamSynchronize(ctx->stream));
        PB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        if (ctx->aux_stream) PB_CUDA(cudaStreamSynchronize(ctx->aux_stream));
        PB_CUDA(cudaFree(s.p));
        s.p = nullptr;
        s.cap = 0;
    }
    size_t want = bytes + (bytes >> 3) + 256;  // grow with slack so repeated calls settle
    PB_CUDA(cudaMalloc(&s.p, want));
    s.cap = want;
    return PB_OK;
}

void scratch_release(pb_scratch &s) {
    if (s.p) cudaFree(s.p);
    s.p = nullptr;
    s.cap = 0;
}

int validate_keys(const pb_keys *keys) {
    PB_REQUIRE(keys != nullptr, "keys is NULL");
    PB_REQUIRE(keys->sym_width == 1 || keys->sym_width == 4, "sym_width must be 1 or 4 (got %u)", keys->sym_width);
    if (keys->n == 0) return PB_OK;
    if (keys->offsets == nullptr) {
        PB_REQUIRE(keys->stride == 0 || keys->data != nullptr, "keys->data is NULL");
    }
    if (keys->sym_width == 4) PB_REQUIRE(((uintptr_t)keys->data & 3u) == 0, "u32 symbols must be 4-byte aligned");
    return PB_OK;
}

static int ensure_pinned(pb_ctx *ctx, int slot, size_t bytes) {
    if (ctx->pinned_cap[slot] >= bytes) return PB_OK;
    if (ctx->pinned[slot]) {
        PB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
        PB_CUDA(cudaFreeHost(ctx->pinned[slot]));
        ctx->pinned[slot] = nullptr;
        ctx->pinned_cap[slot] = 0;
    }
    PB_CUDA(cudaHostAlloc(&ctx->pinned[slot], bytes, cudaHostAllocDefault));
    ctx->pinned_cap[slot] = bytes;
    return PB_OK;
}

static bool is_pinned_host(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// copy [src, src+bytes) from the host into dst on the copy stream (slot-ordered)
static int stage_h2d(pb_ctx *ctx, int slot, void *dst, const void *src, size_t bytes, bool pinned, size_t bounce_off) {
    if (bytes == 0) return PB_OK;
    if (pinned) {
        PB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    } else {
        memcpy((uint8_t *)ctx->pinned[slot] + bounce_off, src, bytes);
        PB_CUDA(cudaMemcpyAsync(dst, (uint8_t *)ctx->pinned[slot] + bounce_off, bytes, cudaMemcpyHostToDevice,
                                ctx->copy_stream));
    }
    return PB_OK;
}

int for_each_chunk(pb_ctx *ctx, const pb_keys *keys, chunk_fn fn, void *user, uint64_t max_chunk_keys) {
    PB_TRY(validate_keys(keys));
    const uint64_t n = keys->n;
    if (n == 0) return PB_OK;
    const uint32_t sw = keys->sym_width;
    if (keys->on_device) {
        uint64_t step = max_chunk_keys ? max_chunk_keys : n;
        for (uint64_t c0 = 0; c0 < n; c0 += step) {
            uint64_t cn = std::min(step, n - c0);
            DevKeys dk;
            dk.sym_width = sw;
            dk.stride = keys->stride;
            dk.n = cn;
            dk.base_symbol = 0;
            if (keys->offsets) {
                dk.data = (const uint8_t *)keys->data;
                dk.offsets = keys->offsets + c0;
                dk.total_bytes = 0;
            } else {
                dk.data = (const uint8_t *)keys->data + c0 * (uint64_t)keys->stride * sw;
                dk.offsets = nullptr;
                dk.total_bytes = cn * (uint64_t)keys->stride * sw;
            }
            PB_TRY(fn(ctx, dk, c0, (int)((c0 / step) & 1), user));
        }
        return PB_OK;
    }
    // host batch: chunked, double buffered
    uint64_t step = (uint64_t)ctx->h2d_chunk_keys;
    if (max_chunk_keys && max_chunk_keys < step) step = max_chunk_keys;
    if (step < 1) step = 1;
    const bool pinned = is_pinned_host(keys->data) && (keys->offsets == nullptr || is_pinned_host(keys->offsets));
    uint64_t chunk_idx = 0;
    for (uint64_t c0 = 0; c0 < n; c0 += step, ++chunk_idx) {
        const int slot = (int)(chunk_idx & 1);
        const uint64_t cn = std::min(step, n - c0);
        uint64_t sym0, sym1;
        if (keys->offsets) {
            sym0 = keys->offsets[c0];
            sym1 = keys->offsets[c0 + cn];
            PB_REQUIRE(sym1 >= sym0, "offsets must be non-decreasing");
        } else {
            sym0 = c0 * (uint64_t)keys->stride;
            sym1 = sym0 + cn * (uint64_t)keys->stride;
        }
        const size_t data_bytes = (size_t)(sym1 - sym0) * sw;
        const size_t off_bytes = keys->offsets ? (size_t)(cn + 1) * sizeof(uint64_t) : 0;
        const size_t data_pad = (data_bytes + 15) & ~(size_t)15;
        // device staging must be free: the kernels of chunk-2 (same slot) are done
        if (chunk_idx >= 2) PB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[slot], 0));
        PB_TRY(scratch_reserve(ctx, ctx->key_stage[slot], data_pad + 16));
        if (off_bytes) PB_TRY(scratch_reserve(ctx, ctx->off_stage[slot], off_bytes));
        if (!pinned) {
            if (chunk_idx >= 2) PB_CUDA(cudaEventSynchronize(ctx->ev_copy[slot]));  // bounce buffer free again
            PB_TRY(ensure_pinned(ctx, slot, data_pad + off_bytes + 64));
        }
        PB_TRY(stage_h2d(ctx, slot, ctx->key_stage[slot].p, (const uint8_t *)keys->data + sym0 * sw, data_bytes, pinned, 0));
        if (off_bytes)
            PB_TRY(stage_h2d(ctx, slot, ctx->off_stage[slot].p, keys->offsets + c0, off_bytes, pinned, data_pad));
        PB_CUDA(cudaEventRecord(ctx->ev_copy[slot], ctx->copy_stream));
        PB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[slot], 0));
        DevKeys dk;
        dk.data = (const uint8_t *)ctx->key_stage[slot].p;
        dk.offsets = off_bytes ? (const uint64_t *)ctx->off_stage[slot].p : nullptr;
        dk.n = cn;
        dk.stride = keys->stride;
        dk.sym_width = sw;
        dk.base_symbol = keys->offsets ? sym0 : 0;
        dk.total_bytes = data_bytes;
        PB_TRY(fn(ctx, dk, c0, slot, user));
        PB_CUDA(cudaEventRecord(ctx->ev_done[slot], ctx->stream));
    }
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

// ---------------------------------------------------------------- kernels: generators + batch hashing
__global__ void __launch_bounds__(256) gen_uniform_kernel(uint64_t seed, uint64_t first, uint64_t n, ulonglong2 *out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g = first + i;
        out[i] = make_ulonglong2(sm64(seed + 2 * g), sm64(seed + 2 * g + 1));
    }
}

__global__ void __launch_bounds__(256) gen_rank_kernel(const uint64_t *ranks, uint64_t n, ulonglong2 *out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = ranks[i];
        out[i] = make_ulonglong2(r, sm64(r));
    }
}

// Zipf(a) ranks on the device (bench utility): the rejection sampler numpy's Generator.zipf uses
// (Devroye, "Non-Uniform Random Variate Generation", p. 551), driven by a counter-based splitmix64 stream so that
// rank i depends only on (seed, i).  Not reference code; the oracle is fed the very same buffer.
__global__ void __launch_bounds__(256) gen_zipf_kernel(uint64_t seed, uint64_t first, uint64_t n, double a, uint64_t *out) {
    const double am1 = a - 1.0, b = pow(2.0, am1);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t ctr = sm64(seed ^ sm64(first + i));
        uint64_t x = 1;
        for (int it = 0; it < 64; ++it) {
            const double u = 1.0 - (double)(sm64(ctr++) >> 11) * (1.0 / 9007199254740992.0);  // (0, 1]
            const double v = (double)(sm64(ctr++) >> 11) * (1.0 / 9007199254740992.0);
            const double xf = floor(pow(u, -1.0 / am1));
            if (xf > 9.2e18 || xf < 1.0) continue;
            const double t = pow(1.0 + 1.0 / xf, am1);
            if (v * xf * (t - 1.0) / (b - 1.0) <= t / b) {
                x = (uint64_t)xf;
                break;
            }
        }
        out[i] = x;
    }
}

template <int KG>
__global__ void __launch_bounds__(256) hash_fixed16_kernel(const uint4 *keys, uint64_t n, uint32_t depth, uint64_t *out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 w = keys[i];
        for (uint32_t s0 = 0; s0 < depth; s0 += KG) {
            uint64_t h[KG];
            fnv_group_16<KG>(w, s0, h);
#pragma unroll
            for (int j = 0; j < KG; ++j)
                if (s0 + j < depth) out[i * depth + s0 + j] = h[j];
        }
    }
}

template <int KG, int SYMW>
__global__ void __launch_bounds__(kTileKeys) hash_staged_kernel(DevKeys dk, uint32_t depth, uint64_t *out) {
    __shared__ TileSmem sm;
    if (threadIdx.x == 0) mbar_init(&sm.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t parity = 0;
    const uint64_t tiles = (dk.n + kTileKeys - 1) / kTileKeys;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint64_t first = tile * kTileKeys;
        const uint32_t count = (uint32_t)min((uint64_t)kTileKeys, dk.n - first);
        KeyRef kr = stage_tile<SYMW>(dk, first, count, sm, parity);
        if (threadIdx.x < count) {
            const uint64_t i = first + threadIdx.x;
            for (uint32_t s0 = 0; s0 < depth; s0 += KG) {
                uint64_t h[KG];
                fnv_group_ptr<KG, SYMW>(kr.p, kr.len, s0, h);
#pragma unroll
                for (int j = 0; j < KG; ++j)
                    if (s0 + j < depth) out[i * depth + s0 + j] = h[j];
            }
        }
    }
}

// fnv_1a(key, seed) / fnv_1a_32(key, seed) for an arbitrary seed (hashes.py:86-122): one hash per key from an explicit
// start value.  Not a hot path (the filters use seeds 0..k-1 through the kernels above): one thread walks one key.
template <int SYMW>
__global__ void __launch_bounds__(256) hash_from_kernel(DevKeys dk, uint64_t h0, int bits32, uint64_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < dk.n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t beg, end;
        if (dk.offsets) {
            beg = dk.offsets[i] - dk.base_symbol;
            end = dk.offsets[i + 1] - dk.base_symbol;
        } else {
            beg = i * (uint64_t)dk.stride;
            end = beg + dk.stride;
        }
        uint64_t h = h0;
        for (uint64_t s = beg; s < end; ++s) {
            const uint32_t sym = SYMW == 4 ? reinterpret_cast<const uint32_t *>(dk.data)[s] : (uint32_t)dk.data[s];
            if (bits32) h = (uint32_t)(((uint32_t)h ^ sym) * 0x01000193u);  // hashes.py:116-121
            else h = fnv_step(h, sym);                                    // hashes.py:99-102
        }
        out[i] = h;
    }
}

struct HashFromArgs {
    uint64_t h0;
    int bits32;
    uint64_t *out_dev, *out_host;
};

static int hash_from_chunk(pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) {
    HashFromArgs *a = (HashFromArgs *)user;
    uint64_t *out = a->out_dev ? a->out_dev + first : nullptr;
    if (!out) {
        PB_TRY(scratch_reserve(ctx, ctx->out_stage[slot], dk.n * sizeof(uint64_t)));
        out = (uint64_t *)ctx->out_stage[slot].p;
    }
    const int grid = grid_for(ctx, dk.n, 256, 8);
    if (dk.sym_width == 4) hash_from_kernel<4><<<grid, 256, 0, ctx->stream>>>(dk, a->h0, a->bits32, out);
    else hash_from_kernel<1><<<grid, 256, 0, ctx->stream>>>(dk, a->h0, a->bits32, out);
    PB_TRY(check_launch(ctx, "hash_from"));
    if (a->out_host) PB_CUDA(cudaMemcpyAsync(a->out_host + first, out, dk.n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    return PB_OK;
}

struct HashArgs {
    uint32_t depth;
    uint64_t *out_dev;   // device output for the whole batch (device-out) or nullptr
    uint64_t *out_host;  // host output
};

template <int KG>
static int launch_hash(pb_ctx *ctx, const DevKeys &dk, uint32_t depth, uint64_t *out) {
    if (is_fixed16(dk)) {
        int grid = grid_for(ctx, dk.n, 256, 8);
        hash_fixed16_kernel<KG><<<grid, 256, 0, ctx->stream>>>((const uint4 *)dk.data, dk.n, depth, out);
    } else {
        uint64_t tiles = (dk.n + kTileKeys - 1) / kTileKeys;
        int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->num_sms * 4);
        if (dk.sym_width == 4)
            hash_staged_kernel<KG, 4><<<grid, kTileKeys, 0, ctx->stream>>>(dk, depth, out);
        else
            hash_staged_kernel<KG, 1><<<grid, kTileKeys, 0, ctx->stream>>>(dk, depth, out);
    }
    return check_launch(ctx, "hash_keys");
}

// default_fnv_1a rows for device-resident keys of any layout: out[i*depth + s] (stream-ordered)
int hash_dev_keys(pb_ctx *ctx, const DevKeys &dk, uint32_t depth, uint64_t *out) {
    switch (pick_group(depth)) {
        case 1: return launch_hash<1>(ctx, dk, depth, out);
        case 2: return launch_hash<2>(ctx, dk, depth, out);
        case 3: return launch_hash<3>(ctx, dk, depth, out);
        case 4: return launch_hash<4>(ctx, dk, depth, out);
        case 5: return launch_hash<5>(ctx, dk, depth, out);
        case 6: return launch_hash<6>(ctx, dk, depth, out);
        case 7: return launch_hash<7>(ctx, dk, depth, out);
        default: return launch_hash<8>(ctx, dk, depth, out);
    }
}

static int hash_chunk(pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) {
    HashArgs *a = (HashArgs *)user;
    uint64_t *out = a->out_dev ? a->out_dev + first * a->depth : nullptr;
    if (!out) {
        PB_TRY(scratch_reserve(ctx, ctx->out_stage[slot], dk.n * a->depth * sizeof(uint64_t)));
        out = (uint64_t *)ctx->out_stage[slot].p;
    }
    const int st = hash_dev_keys(ctx, dk, a->depth, out);
    PB_TRY(st);
    if (a->out_host)
        PB_CUDA(cudaMemcpyAsync(a->out_host + first * a->depth, out, dk.n * a->depth * sizeof(uint64_t),
                                cudaMemcpyDeviceToHost, ctx->stream));
    return PB_OK;
}

}  // namespace pb

using namespace pb;

extern "C" {

int pb_version(void) { return PB200_VERSION; }
const char *pb_last_error(void) { return pb::g_err; }

int pb_device_count(int *out) {
    PB_REQUIRE(out, "out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *out = n;
    return PB_OK;
}

int pb_ctx_create(int device, void *stream, pb_ctx **out) {
    PB_REQUIRE(out, "out is NULL");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible: libpb200 has no CPU fallback");
        return PB_ERR_NO_DEVICE;
    }
    PB_REQUIRE(device >= 0 && device < n, "device %d out of range (0..%d)", device, n - 1);
    PB_CUDA(cudaSetDevice(device));
    pb_ctx *c = new (std::nothrow) pb_ctx();
    if (!c) return PB_ERR_OOM;
    c->device = device;
    cudaDeviceProp prop;
    PB_CUDA(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        PB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    PB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    PB_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        PB_CUDA(cudaEventCreateWithFlags(&c->ev_part[i], cudaEventDisableTiming));
        PB_CUDA(cudaEventCreateWithFlags(&c->ev_apply[i], cudaEventDisableTiming));
        PB_CUDA(cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming));
        PB_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
    PB_CUDA(cudaHostAlloc(&c->pinned_small, 4096, cudaHostAllocDefault));
    *out = c;
    return PB_OK;
}

int pb_ctx_destroy(pb_ctx *ctx) {
    if (!ctx) return PB_OK;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev_part[i]) cudaEventDestroy(ctx->ev_part[i]);
        if (ctx->ev_apply[i]) cudaEventDestroy(ctx->ev_apply[i]);
        scratch_release(ctx->key_stage[i]);
        scratch_release(ctx->off_stage[i]);
        scratch_release(ctx->aux_stage[i]);
        scratch_release(ctx->out_stage[i]);
        if (ctx->pinned[i]) cudaFreeHost(ctx->pinned[i]);
        if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    }
    scratch_release(ctx->part_stage);
    scratch_release(ctx->part_cursors);
    scratch_release(ctx->small);
    scratch_release(ctx->flush);
    scratch_release(ctx->claim_set);
    scratch_release(ctx->claim_bitmap);
    if (ctx->pinned_small) cudaFreeHost(ctx->pinned_small);
    for (const pb_timed_launch &t : ctx->timed) {
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
    }
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->pending_e0) cudaEventDestroy(ctx->pending_e0);
    cudaStreamDestroy(ctx->copy_stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PB_OK;
}

int pb_ctx_synchronize(pb_ctx *ctx) {
    PB_REQUIRE(ctx, "ctx is NULL");
    DeviceGuard g(ctx->device);
    PB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->aux_stream));
    return PB_OK;
}

int pb_ctx_stream(pb_ctx *ctx, void **out_stream) {
    PB_REQUIRE(ctx && out_stream, "NULL argument");
    *out_stream = (void *)ctx->stream;
    return PB_OK;
}

int pb_ctx_launch_count(pb_ctx *ctx, uint64_t *out) {
    PB_REQUIRE(ctx && out, "NULL argument");
    *out = ctx->launches;
    return PB_OK;
}

// Per-kernel device time of the launches bracketed since the last call ("kernel_timing" option):
// lines of "name launches total_ms" into buf.  Synchronizes the compute stream.
int pb_ctx_kernel_times(pb_ctx *ctx, char *buf, size_t cap) {
    PB_REQUIRE(ctx && buf && cap > 0, "NULL argument");
    DeviceGuard g(ctx->device);
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    struct Acc {
        const char *name;
        uint64_t n;
        double ms;
    };
    std::vector<Acc> acc;
    for (const pb_timed_launch &t : ctx->timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.e0, t.e1) != cudaSuccess) {
            cudaGetLastError();
            ms = 0.f;
        }
        bool found = false;
        for (Acc &a : acc)
            if (!strcmp(a.name, t.name)) {
                a.n++;
                a.ms += ms;
                found = true;
                break;
            }
        if (!found) acc.push_back(Acc{t.name, 1, (double)ms});
        ctx->event_pool.push_back(t.e0);
        ctx->event_pool.push_back(t.e1);
    }
    ctx->timed.clear();
    size_t off = 0;
    buf[0] = 0;
    for (const Acc &a : acc) {
        int w = snprintf(buf + off, cap - off, "%s %llu %.6f\n", a.name, (unsigned long long)a.n, a.ms);
        if (w < 0 || (size_t)w >= cap - off) break;
        off += (size_t)w;
    }
    return PB_OK;
}

static int64_t *option_slot(pb_ctx *ctx, const char *name) {
    if (!strcmp(name, "bloom_insert_mode")) return &ctx->bloom_insert_mode;
    if (!strcmp(name, "bloom_window_log2_bits")) return &ctx->bloom_window_log2_bits;
    if (!strcmp(name, "stage_bytes")) return &ctx->stage_bytes;
    if (!strcmp(name, "bloom_apply_cpw_per_sm")) return &ctx->bloom_apply_cpw_per_sm;
    if (!strcmp(name, "bloom_min_chunks")) return &ctx->bloom_min_chunks;
    if (!strcmp(name, "bloom_check_mode")) return &ctx->bloom_check_mode;
    if (!strcmp(name, "bloom_part_ctas_per_sm")) return &ctx->bloom_part_ctas_per_sm;
    if (!strcmp(name, "bloom_overlap")) return &ctx->bloom_overlap;
    if (!strcmp(name, "bloom_part_tile")) return &ctx->bloom_part_tile;
    if (!strcmp(name, "h2d_chunk_keys")) return &ctx->h2d_chunk_keys;
    if (!strcmp(name, "cms_aggregate")) return &ctx->cms_aggregate;
    if (!strcmp(name, "cms_hot_cache")) return &ctx->cms_hot_cache;
    if (!strcmp(name, "cuckoo_serial")) return &ctx->cuckoo_serial;
    if (!strcmp(name, "p2p_timeout_ms")) return &ctx->p2p_timeout_ms;
    if (!strcmp(name, "p2p_copy_lanes")) return &ctx->p2p_copy_lanes;
    if (!strcmp(name, "kernel_timing")) return &ctx->kernel_timing;
    return nullptr;
}

int pb_ctx_set_option(pb_ctx *ctx, const char *name, int64_t value) {
    PB_REQUIRE(ctx && name, "NULL argument");
    int64_t *s = option_slot(ctx, name);
    PB_REQUIRE(s, "unknown option '%s'", name);
    if (!strcmp(name, "bloom_window_log2_bits")) PB_REQUIRE(value >= 10 && value <= 40, "bloom_window_log2_bits out of range");
    if (!strcmp(name, "h2d_chunk_keys")) PB_REQUIRE(value >= 1, "h2d_chunk_keys must be >= 1");
    if (!strcmp(name, "bloom_insert_mode")) PB_REQUIRE(value >= 0 && value <= 2, "bloom_insert_mode must be 0, 1 or 2");
    *s = value;
    return PB_OK;
}

int pb_ctx_get_option(pb_ctx *ctx, const char *name, int64_t *out) {
    PB_REQUIRE(ctx && name && out, "NULL argument");
    int64_t *s = option_slot(ctx, name);
    PB_REQUIRE(s, "unknown option '%s'", name);
    *out = *s;
    return PB_OK;
}

int pb_host_alloc(size_t bytes, void **out) {
    PB_REQUIRE(out, "out is NULL");
    PB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return PB_OK;
}
int pb_host_free(void *p) {
    if (p) PB_CUDA(cudaFreeHost(p));
    return PB_OK;
}
int pb_dev_alloc(pb_ctx *ctx, size_t bytes, void **out) {
    PB_REQUIRE(ctx && out, "NULL argument");
    DeviceGuard g(ctx->device);
    PB_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return PB_OK;
}
int pb_dev_free(pb_ctx *ctx, void *p) {
    PB_REQUIRE(ctx, "ctx is NULL");
    DeviceGuard g(ctx->device);
    if (p) {
        PB_CUDA(cudaStreamSynchronize(ctx->stream));
        PB_CUDA(cudaFree(p));
    }
    return PB_OK;
}
int pb_memcpy_h2d(pb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
    PB_REQUIRE(ctx, "ctx is NULL");
    DeviceGuard g(ctx->device);
    PB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}
int pb_memcpy_d2h(pb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
    PB_REQUIRE(ctx, "ctx is NULL");
    DeviceGuard g(ctx->device);
    PB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}
int pb_memset_dev(pb_ctx *ctx, void *dst_dev, int value, size_t bytes) {
    PB_REQUIRE(ctx, "ctx is NULL");
    DeviceGuard g(ctx->device);
    PB_CUDA(cudaMemsetAsync(dst_dev, value, bytes, ctx->stream));
    return PB_OK;
}
int pb_flush_l2(pb_ctx *ctx) {
    PB_REQUIRE(ctx, "ctx is NULL");
    DeviceGuard g(ctx->device);
    size_t bytes = std::max<size_t>(ctx->l2_bytes * 2, (size_t)256 << 20);
    PB_TRY(scratch_reserve(ctx, ctx->flush, bytes));
    PB_CUDA(cudaMemsetAsync(ctx->flush.p, 0xA5, bytes, ctx->stream));
    return PB_OK;
}

int pb_hash_keys(pb_ctx *ctx, const pb_keys *keys, uint32_t depth, uint64_t *out, int out_on_device) {
    PB_REQUIRE(ctx && out, "NULL argument");
    PB_REQUIRE(depth >= 1, "depth must be >= 1");
    DeviceGuard g(ctx->device);
    HashArgs a;
    a.depth = depth;
    a.out_dev = out_on_device ? out : nullptr;
    a.out_host = out_on_device ? nullptr : out;
    PB_TRY(for_each_chunk(ctx, keys, hash_chunk, &a));
    if (!out_on_device) PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

int pb_hash_keys_from(pb_ctx *ctx, const pb_keys *keys, uint64_t start_value, int bits, uint64_t *out, int out_on_device) {
    PB_REQUIRE(ctx && out, "NULL argument");
    PB_REQUIRE(bits == 64 || bits == 32, "bits must be 64 or 32");
    DeviceGuard g(ctx->device);
    HashFromArgs a{start_value, bits == 32 ? 1 : 0, out_on_device ? out : nullptr, out_on_device ? nullptr : out};
    PB_TRY(for_each_chunk(ctx, keys, hash_from_chunk, &a));
    if (!out_on_device) PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

int pb_gen_uniform_keys(pb_ctx *ctx, uint64_t seed, uint64_t first, uint64_t n, void *out_dev) {
    PB_REQUIRE(ctx && (out_dev || n == 0), "NULL argument");
    PB_REQUIRE(((uintptr_t)out_dev & 15u) == 0, "out_dev must be 16-byte aligned");
    if (n == 0) return PB_OK;
    DeviceGuard g(ctx->device);
    gen_uniform_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(seed, first, n, (ulonglong2 *)out_dev);
    return check_launch(ctx, "gen_uniform_keys");
}

int pb_gen_zipf_ranks(pb_ctx *ctx, uint64_t seed, uint64_t first, uint64_t n, double a, uint64_t *out_ranks_dev) {
    PB_REQUIRE(ctx && (out_ranks_dev || n == 0), "NULL argument");
    PB_REQUIRE(a > 1.0, "the Zipf exponent must be > 1");
    if (n == 0) return PB_OK;
    DeviceGuard g(ctx->device);
    gen_zipf_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(seed, first, n, a, out_ranks_dev);
    return check_launch(ctx, "gen_zipf_ranks");
}

int pb_gen_rank_keys(pb_ctx *ctx, const uint64_t *ranks_dev, uint64_t n, void *out_dev) {
    PB_REQUIRE(ctx && ((ranks_dev && out_dev) || n == 0), "NULL argument");
    PB_REQUIRE(((uintptr_t)out_dev & 15u) == 0, "out_dev must be 16-byte aligned");
    if (n == 0) return PB_OK;
    DeviceGuard g(ctx->device);
    gen_rank_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(ranks_dev, n, (ulonglong2 *)out_dev);
    return check_launch(ctx, "gen_rank_keys");
}

// ---- host-callable checks of the shared integer arithmetic (no GPU needed; used by the CPU test-suite
// to pin fastmod / FNV / cuckoo index math of the exact code the kernels compile) ----------------------
uint64_t pbt_fnv1a(const uint8_t *p, uint64_t len, uint64_t seed) {
    uint64_t h = fnv_init(seed);
    for (uint64_t i = 0; i < len; ++i) h = fnv_step(h, p[i]);
    return h;
}
uint64_t pbt_fastmod(uint64_t h, uint64_t m) {
    FastMod f = make_fastmod(m);
    return fastmod(h, f);
}
void pbt_cuckoo_info(uint64_t h, uint32_t fp_bits, uint64_t capacity, uint32_t *fp, uint64_t *i1, uint64_t *i2) {
    FastMod f = make_fastmod(capacity);
    *fp = cuckoo_fingerprint(h, fp_bits);
    *i1 = fastmod(*fp, f);
    *i2 = fastmod(fnv_of_decimal(*fp), f);
}
// host twin of mod_fast33 (pb_bloom_part.cuh): the same steps in portable C, for m >= 2^33
uint64_t pbt_mod_fast33(uint64_t h, uint64_t m) {
    const uint32_t r32 = (uint32_t)make_fastmod(m).recip;
    const int32_t m_lo_s = (int32_t)(uint32_t)m;
    const uint32_t m_hi_adj = (uint32_t)(m >> 32) + (uint32_t)((m >> 31) & 1u);
    const uint32_t t_hi = (uint32_t)(((uint64_t)(uint32_t)h * r32) >> 32);
    const uint64_t u = (uint64_t)(uint32_t)(h >> 32) * r32 + t_hi;
    const int32_t nq = (int32_t)(0u - (uint32_t)(u >> 32));
    const uint64_t r = h + (uint64_t)((int64_t)nq * (int64_t)m_lo_s);
    const uint32_t r_hi = (uint32_t)(r >> 32) + (uint32_t)nq * m_hi_adj;
    const uint64_t rr = ((uint64_t)r_hi << 32) | (uint32_t)r;
    return rr >= m ? rr - m : rr;
}
uint64_t pbt_sm64(uint64_t x) { return sm64(x); }
int pbt_pick_group(uint32_t k) { return pick_group(k); }

}  // extern "C"
