// pb_cms.cu -- CountMinSketch.add / check for whole batches
// (reference: probables/countminsketch/countminsketch.py:257-288, :323-340, :429-453).
//
// State: int32[depth][width] row-major; bin of row r = (h_r % width) + r*width (:275).
// add: bins saturate at INT32_MAX (:280-282).  For non-negative adds the saturating result is
// order-free (min(sum, INT32_MAX)), which is what lets a batch run as concurrent atomics:
//   * "safe" launches (the handle has provably added < 2^31 in total, so no bin can overflow) use
//     plain RED.ADD.S32;
//   * otherwise every update first looks at the bin: below 2^30 with a small addend it is still a RED
//     (at most resident_threads * 2048 < 2^30 can be in flight past that test), else an exact
//     saturating CAS loop.
// Skewed streams (Zipf) hammer a handful of bins: equal keys inside a warp are combined with
// __match_any_sync first, so a warp issues one atomic per distinct key and row.
#include <algorithm>
#include <new>

#include "pb_common.cuh"
#include "pb_hash.cuh"
#include "pb_keys.cuh"

using namespace pb;

struct pb_cms {
    pb_ctx *ctx = nullptr;
    uint32_t width = 0, depth = 0;
    int32_t *bins = nullptr;
    uint64_t count = 0;     // width * depth
    uint64_t abs_added = 0;  // upper bound of sum |num_els| ever applied (decides the safe path)
    FastMod fm;
};

namespace pb {

constexpr int kMaxDepthLocal = 64;
constexpr int64_t kI32Max = 2147483647LL;
constexpr int64_t kI32Min = -2147483647LL - 1;
constexpr int64_t kI64Max = 9223372036854775807LL;

struct CmsDev {
    int32_t *bins;
    FastMod fm;
    uint32_t width, depth;
};

// countminsketch.py:276-284 for one bin, concurrent-safe
template <bool SAFE>
__device__ __forceinline__ void cms_bin_add(int32_t *bin, int64_t n) {
    if (n == 0) return;
    if (SAFE) {
        atomicAdd(bin, (int32_t)n);  // RED.ADD
        return;
    }
    if (n > 0 && n <= 2048) {
        const int32_t cur = *reinterpret_cast<volatile int32_t *>(bin);
        if (cur < (1 << 30)) {
            atomicAdd(bin, (int32_t)n);
            return;
        }
    }
    int32_t old = *reinterpret_cast<volatile int32_t *>(bin);
    while (true) {
        int64_t v = (int64_t)old + n;
        if (v > kI32Max) v = kI32Max;  // :280-282
        if (v < kI32Min) v = kI32Min;  // array('i') cannot go lower (the reference raises here)
        if ((int32_t)v == old) return;
        const int32_t prev = atomicCAS(bin, old, (int32_t)v);
        if (prev == old) return;
        old = prev;
    }
}

// Combine equal keys of a warp: returns the group's total in the leader lane, 0 elsewhere.
// key identity = the full 128 bits (fixed16) or the first two seeds' hashes (generic) -- see callers.
__device__ __forceinline__ int64_t warp_combine(uint64_t id0, uint64_t id1, int64_t n, bool active, bool small_n) {
    const uint32_t live = __ballot_sync(0xffffffffu, active);
    if (!active) return 0;
    // all lanes of `live` reach here together
    const uint32_t all_small = __all_sync(live, small_n);
    if (!all_small) return n;
    uint32_t m = __match_any_sync(live, id0);
    m &= __match_any_sync(live, id1);
    const int32_t total = __reduce_add_sync(m, (int32_t)n);
    const uint32_t leader = __ffs(m) - 1;
    return (threadIdx.x & 31u) == leader ? (int64_t)total : 0;
}

// Per-CTA write-back cache for hot counters.  Under a skewed stream the few counters of the heaviest keys take a
// RED from almost every warp, and L2 applies atomics on one address one after the other: the round-1 profile of the
// Zipf(1.1) workload (rank 1 = 9.45 % of all keys) had the kernel time equal to that serial chain, not to the L2
// atomic throughput.  Counters are therefore first looked up in a small direct-mapped table in shared memory
// (tag = flat bin index); a hit or a claimed empty slot is a shared-memory atomic, everything else goes to L2 as
// before, and the table is flushed with one RED per occupied slot when the CTA is done.  Only used on the "safe"
// path, where the sum of everything ever added fits int32, so neither the cached counts nor their flush can overflow.
constexpr int kHotLog2 = 11;
constexpr uint32_t kHotEmpty = 0xFFFFFFFFu;

struct HotTable {
    uint32_t tag[1 << kHotLog2];
    int32_t cnt[1 << kHotLog2];
};

__device__ __forceinline__ void hot_add(HotTable &t, int32_t *bins, uint32_t idx, int32_t add) {
    const uint32_t slot = (idx * 0x9E3779B1u) >> (32 - kHotLog2);
    uint32_t tag = t.tag[slot];
    if (tag == kHotEmpty) {
        const uint32_t prev = atomicCAS(&t.tag[slot], kHotEmpty, idx);
        tag = prev == kHotEmpty ? idx : prev;
    }
    if (tag == idx) atomicAdd(&t.cnt[slot], add);
    else atomicAdd(bins + idx, add);
}

template <int KG, bool SAFE, bool AGG>
__global__ void __launch_bounds__(256) cms_add_fixed16(const uint4 *__restrict__ keys, uint64_t n, const int64_t *__restrict__ num_els,
                                                       int64_t scalar, CmsDev c, int use_hot) {
    __shared__ HotTable hot;
    const bool hot_on = SAFE && use_hot;
    if (hot_on) {
        for (uint32_t q = threadIdx.x; q < (1u << kHotLog2); q += blockDim.x) {
            hot.tag[q] = kHotEmpty;
            hot.cnt[q] = 0;
        }
        __syncthreads();
    }
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + stride - 1) / stride;
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint4 nextw = make_uint4(0, 0, 0, 0);
    if (i < n) nextw = __ldcs(keys + i);
    for (uint64_t r = 0; r < rounds; ++r, i += stride) {
        const bool active = i < n;
        const uint4 w = nextw;  // loaded one round ahead: the hash never waits for HBM
        if (i + stride < n) nextw = __ldcs(keys + i + stride);
        int64_t add = 0;
        if (active) add = num_els ? num_els[i] : scalar;
        if (AGG) {
            const bool small_n = add > -(1 << 24) && add < (1 << 24);
            add = warp_combine(((uint64_t)w.y << 32) | w.x, ((uint64_t)w.w << 32) | w.z, add, active, small_n);
        }
        if (!active || add == 0) continue;
        for (uint32_t s0 = 0; s0 < c.depth; s0 += KG) {
            uint64_t h[KG];
            fnv_group_16<KG>(w, s0, h);
#pragma unroll
            for (int j = 0; j < KG; ++j) {
                if (s0 + j < c.depth) {
                    const uint64_t idx = fastmod(h[j], c.fm) + (uint64_t)(s0 + j) * c.width;
                    if (hot_on) hot_add(hot, c.bins, (uint32_t)idx, (int32_t)add);
                    else cms_bin_add<SAFE>(c.bins + idx, add);
                }
            }
        }
    }
    if (hot_on) {
        __syncthreads();
        for (uint32_t q = threadIdx.x; q < (1u << kHotLog2); q += blockDim.x)
            if (hot.tag[q] != kHotEmpty && hot.cnt[q] != 0) atomicAdd(c.bins + hot.tag[q], hot.cnt[q]);
    }
}

// countminsketch.py:429-453 on the gathered bin values
__device__ __forceinline__ int64_t floordiv64(int64_t a, int64_t b) {
    int64_t q = a / b, r = a % b;
    if (r != 0 && ((r < 0) != (b < 0))) --q;
    return q;
}

struct CmsQuery {
    int query_type;  // 0 min, 1 mean, 2 mean-min
    int64_t elements_added;
};

// running reductions for min / mean; mean-min needs the adjusted values sorted (median)
struct QueryAcc {
    int64_t mn, mx, sum;
    int64_t vals[kMaxDepthLocal];
    __device__ __forceinline__ void init() {
        mn = kI64Max;
        mx = -kI64Max;
        sum = 0;
    }
    __device__ __forceinline__ void push(int r, int64_t v, int qt) {
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
        sum += v;
        if (qt == 2 && r < kMaxDepthLocal) vals[r] = v;
    }
    __device__ __forceinline__ int64_t finish(const CmsQuery &q, uint32_t depth, uint32_t width) {
        if (q.query_type == 0) return mn;                            // :430-432
        if (q.query_type == 1) return floordiv64(sum, (int64_t)depth);  // :434-436
        if (mn == 0 && mx == 0) return 0;                            // :440-441
        const int d = (int)(depth < (uint32_t)kMaxDepthLocal ? depth : (uint32_t)kMaxDepthLocal);
        for (int r = 0; r < d; ++r) vals[r] = vals[r] - floordiv64(q.elements_added - vals[r], (int64_t)width - 1);  // :443-445
        for (int a = 1; a < d; ++a) {  // insertion sort (:447)
            const int64_t x = vals[a];
            int b = a - 1;
            while (b >= 0 && vals[b] > x) {
                vals[b + 1] = vals[b];
                --b;
            }
            vals[b + 1] = x;
        }
        if (d % 2 == 0) return floordiv64(vals[d / 2] + vals[d / 2 - 1], 2);  // :448-450
        return vals[d / 2];
    }
};

template <int KG>
__global__ void __launch_bounds__(256)
    cms_check_fixed16(const uint4 *__restrict__ keys, uint64_t n, CmsDev c, CmsQuery q, int64_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 w = __ldcs(keys + i);
        QueryAcc acc;
        acc.init();
        for (uint32_t s0 = 0; s0 < c.depth; s0 += KG) {
            uint64_t h[KG];
            int32_t v[KG];
            fnv_group_16<KG>(w, s0, h);
#pragma unroll
            for (int j = 0; j < KG; ++j)
                v[j] = (s0 + j < c.depth) ? __ldg(c.bins + fastmod(h[j], c.fm) + (uint64_t)(s0 + j) * c.width) : 0;
#pragma unroll
            for (int j = 0; j < KG; ++j)
                if (s0 + j < c.depth) acc.push((int)(s0 + j), (int64_t)v[j], q.query_type);
        }
        out[i] = acc.finish(q, c.depth, c.width);
    }
}

// staged (generic keys) add + check
template <int KG, int SYMW, bool CHECK, bool SAFE>
__global__ void __launch_bounds__(kTileKeys) cms_staged(DevKeys dk, const int64_t *__restrict__ num_els, int64_t scalar, CmsDev c,
                                                        CmsQuery q, int64_t *__restrict__ out) {
    __shared__ TileSmem sm;
    if (threadIdx.x == 0) mbar_init(&sm.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t parity = 0;
    const uint64_t tiles = (dk.n + kTileKeys - 1) / kTileKeys;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint64_t first = tile * kTileKeys;
        const uint32_t count = (uint32_t)min((uint64_t)kTileKeys, dk.n - first);
        const KeyRef kr = stage_tile<SYMW>(dk, first, count, sm, parity);
        if (threadIdx.x < count) {
            const uint64_t i = first + threadIdx.x;
            QueryAcc acc;
            if (CHECK) acc.init();
            const int64_t add = CHECK ? 0 : (num_els ? num_els[i] : scalar);
            for (uint32_t s0 = 0; s0 < c.depth; s0 += KG) {
                uint64_t h[KG];
                fnv_group_ptr<KG, SYMW>(kr.p, kr.len, s0, h);
#pragma unroll
                for (int j = 0; j < KG; ++j) {
                    if (s0 + j < c.depth) {
                        int32_t *bin = c.bins + fastmod(h[j], c.fm) + (uint64_t)(s0 + j) * c.width;
                        if (CHECK) acc.push((int)(s0 + j), (int64_t)__ldg(bin), q.query_type);
                        else cms_bin_add<SAFE>(bin, add);
                    }
                }
            }
            if (CHECK) out[i] = acc.finish(q, c.depth, c.width);
        }
    }
}

// add_alt / check_alt with host-made hashes (:267-288, :332-340)
template <bool SAFE>
__global__ void __launch_bounds__(256) cms_add_hashes_kernel(const uint64_t *__restrict__ h, uint64_t n, const int64_t *__restrict__ num_els,
                                                             int64_t scalar, CmsDev c) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int64_t add = num_els ? num_els[i] : scalar;
        for (uint32_t r = 0; r < c.depth; ++r)
            cms_bin_add<SAFE>(c.bins + fastmod(h[i * c.depth + r], c.fm) + (uint64_t)r * c.width, add);
    }
}
__global__ void __launch_bounds__(256)
    cms_check_hashes_kernel(const uint64_t *__restrict__ h, uint64_t n, CmsDev c, CmsQuery q, int64_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        QueryAcc acc;
        acc.init();
        for (uint32_t r = 0; r < c.depth; ++r)
            acc.push((int)r, (int64_t)__ldg(c.bins + fastmod(h[i * c.depth + r], c.fm) + (uint64_t)r * c.width), q.query_type);
        out[i] = acc.finish(q, c.depth, c.width);
    }
}

// countminsketch.py:380-391
__global__ void __launch_bounds__(256) cms_join_kernel(int32_t *__restrict__ a, const int32_t *__restrict__ b, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int32_t x = a[i];
        if (x == (int32_t)kI32Min || x == (int32_t)kI32Max) continue;
        int64_t v = (int64_t)x + (int64_t)b[i];
        v = v > kI32Max ? kI32Max : (v < kI32Min ? kI32Min : v);
        a[i] = (int32_t)v;
    }
}

// multi-GPU merge: int32 table -> int64 (exact sums across ranks), and the saturated way back (:380-391 for
// non-negative tables: a chain of saturating joins equals min(sum, INT32_MAX))
__global__ void __launch_bounds__(256) cms_widen_kernel(const int32_t *__restrict__ a, long long *__restrict__ out, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = a[i];
}
__global__ void __launch_bounds__(256) cms_narrow_kernel(const long long *__restrict__ s, int32_t *__restrict__ out, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const long long v = s[i];
        out[i] = (int32_t)(v > kI32Max ? kI32Max : (v < kI32Min ? kI32Min : v));
    }
}

// sum and sum-of-abs of an int64 array (saturating), for elements_added with per-key num_els on the device
__global__ void __launch_bounds__(256) sum_i64_kernel(const int64_t *__restrict__ v, uint64_t n, long long *out_sum,
                                                      unsigned long long *out_abs) {
    double approx = 0.0;
    long long s = 0;
    unsigned long long a = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int64_t x = v[i];
        s += x;
        a += (unsigned long long)(x < 0 ? -x : x);
        approx += fabs((double)x);
    }
    (void)approx;
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        a += __shfl_xor_sync(0xffffffffu, a, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long *)out_sum, (unsigned long long)s);
        atomicAdd(out_abs, a);
    }
}

static CmsDev dev_view(const pb_cms *c) {
    CmsDev d;
    d.bins = c->bins;
    d.fm = c->fm;
    d.width = c->width;
    d.depth = c->depth;
    return d;
}

struct CmsAddArgs {
    pb_cms *c;
    const int64_t *num_els;  // host or device per keys->on_device; nullptr -> scalar
    bool num_els_on_device;
    int64_t scalar;
    bool safe;
};

template <int KG, bool SAFE>
static int launch_cms_add(pb_ctx *ctx, const DevKeys &dk, const int64_t *ne, int64_t scalar, const CmsDev &cd) {
    CmsQuery q{0, 0};
    launch_begin(ctx);
    if (is_fixed16(dk)) {
        const int grid = grid_for(ctx, dk.n, 256, 8);
        // the hot-counter cache needs 32-bit flat bin indices and int32 addends (guaranteed on the safe path)
        const int use_hot = SAFE && ctx->cms_hot_cache && (uint64_t)cd.width * cd.depth < 0xFFFFFFFFull ? 1 : 0;
        // warp-level combining of equal keys (match_any + reduce) costs more instructions than it saves once the
        // shared-memory hot-counter cache absorbs the skew (r2: 20.4 -> 33.3 G keys/s on Zipf(1.1) without it), so by
        // default ("cms_aggregate" = 2) it only runs on the exact saturating path, where the cache cannot be used
        const bool agg = ctx->cms_aggregate == 1 || (ctx->cms_aggregate == 2 && !use_hot);
        if (agg)
            cms_add_fixed16<KG, SAFE, true><<<grid, 256, 0, ctx->stream>>>((const uint4 *)dk.data, dk.n, ne, scalar, cd, use_hot);
        else
            cms_add_fixed16<KG, SAFE, false><<<grid, 256, 0, ctx->stream>>>((const uint4 *)dk.data, dk.n, ne, scalar, cd, use_hot);
    } else {
        uint64_t tiles = (dk.n + kTileKeys - 1) / kTileKeys;
        int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->num_sms * 4);
        if (dk.sym_width == 4) cms_staged<KG, 4, false, SAFE><<<grid, kTileKeys, 0, ctx->stream>>>(dk, ne, scalar, cd, q, nullptr);
        else cms_staged<KG, 1, false, SAFE><<<grid, kTileKeys, 0, ctx->stream>>>(dk, ne, scalar, cd, q, nullptr);
    }
    return check_launch(ctx, "cms_add");
}

template <int KG>
static int launch_cms_check(pb_ctx *ctx, const DevKeys &dk, const CmsDev &cd, const CmsQuery &q, int64_t *out) {
    launch_begin(ctx);
    if (is_fixed16(dk)) {
        cms_check_fixed16<KG><<<grid_for(ctx, dk.n, 256, 8), 256, 0, ctx->stream>>>((const uint4 *)dk.data, dk.n, cd, q, out);
    } else {
        uint64_t tiles = (dk.n + kTileKeys - 1) / kTileKeys;
        int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->num_sms * 4);
        if (dk.sym_width == 4) cms_staged<KG, 4, true, true><<<grid, kTileKeys, 0, ctx->stream>>>(dk, nullptr, 0, cd, q, out);
        else cms_staged<KG, 1, true, true><<<grid, kTileKeys, 0, ctx->stream>>>(dk, nullptr, 0, cd, q, out);
    }
    return check_launch(ctx, "cms_check");
}

#define PB_DISPATCH_KG(kg, CALL)      \
    switch (kg) {                     \
        case 1: st = CALL(1); break;  \
        case 2: st = CALL(2); break;  \
        case 3: st = CALL(3); break;  \
        case 4: st = CALL(4); break;  \
        case 5: st = CALL(5); break;  \
        case 6: st = CALL(6); break;  \
        case 7: st = CALL(7); break;  \
        default: st = CALL(8); break; \
    }

static int cms_add_chunk(pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) {
    CmsAddArgs *a = (CmsAddArgs *)user;
    const CmsDev cd = dev_view(a->c);
    const int64_t *ne = nullptr;
    if (a->num_els) {
        if (a->num_els_on_device) {
            ne = a->num_els + first;
        } else {
            PB_TRY(scratch_reserve(ctx, ctx->aux_stage[slot], dk.n * 8));
            PB_CUDA(cudaMemcpyAsync(ctx->aux_stage[slot].p, a->num_els + first, dk.n * 8, cudaMemcpyHostToDevice, ctx->stream));
            ne = (const int64_t *)ctx->aux_stage[slot].p;
        }
    }
    int st = PB_OK;
    const int kg = pick_group(a->c->depth);
    if (a->safe) {
#define CALL(K) launch_cms_add<K, true>(ctx, dk, ne, a->scalar, cd)
        PB_DISPATCH_KG(kg, CALL)
#undef CALL
    } else {
#define CALL(K) launch_cms_add<K, false>(ctx, dk, ne, a->scalar, cd)
        PB_DISPATCH_KG(kg, CALL)
#undef CALL
    }
    return st;
}

struct CmsCheckArgs {
    pb_cms *c;
    CmsQuery q;
    int64_t *out_dev;
    int64_t *out_host;
};

static int cms_check_chunk(pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) {
    CmsCheckArgs *a = (CmsCheckArgs *)user;
    const CmsDev cd = dev_view(a->c);
    int64_t *out = a->out_dev ? a->out_dev + first : nullptr;
    if (!out) {
        PB_TRY(scratch_reserve(ctx, ctx->out_stage[slot], dk.n * 8));
        out = (int64_t *)ctx->out_stage[slot].p;
    }
    int st = PB_OK;
    const int kg = pick_group(a->c->depth);
#define CALL(K) launch_cms_check<K>(ctx, dk, cd, a->q, out)
    PB_DISPATCH_KG(kg, CALL)
#undef CALL
    PB_TRY(st);
    if (a->out_host) PB_CUDA(cudaMemcpyAsync(a->out_host + first, out, dk.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return PB_OK;
}

// saturating int64 add as countminsketch.py:285-287 (plus the lower clamp of :318-319)
static int64_t sat_add_i64(int64_t a, int64_t b) {
    __int128 v = (__int128)a + (__int128)b;
    if (v > (__int128)kI64Max) return kI64Max;
    if (v < -(__int128)kI64Max - 1) return (int64_t)(-kI64Max - 1);
    return (int64_t)v;
}

// total and |total| of the batch's num_els (host side for host arrays, reduction kernel for device arrays)
static int batch_totals(pb_ctx *ctx, const int64_t *num_els, bool on_device, int64_t scalar, uint64_t n, int64_t *sum,
                        uint64_t *abs_sum) {
    if (!num_els) {
        __int128 s = (__int128)scalar * (__int128)n;
        *sum = s > (__int128)kI64Max ? kI64Max : (s < -(__int128)kI64Max - 1 ? (-kI64Max - 1) : (int64_t)s);
        __int128 a = s < 0 ? -s : s;
        *abs_sum = a > (__int128)kI64Max ? (uint64_t)kI64Max : (uint64_t)a;
        return PB_OK;
    }
    if (!on_device) {
        int64_t s = 0;
        __int128 a = 0;
        for (uint64_t i = 0; i < n; ++i) {
            s = sat_add_i64(s, num_els[i]);
            a += num_els[i] < 0 ? -(__int128)num_els[i] : (__int128)num_els[i];
        }
        *sum = s;
        *abs_sum = a > (__int128)kI64Max ? (uint64_t)kI64Max : (uint64_t)a;
        return PB_OK;
    }
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    long long *ds = (long long *)ctx->small.p + 16;
    unsigned long long *da = (unsigned long long *)ctx->small.p + 17;
    PB_CUDA(cudaMemsetAsync(ds, 0, 16, ctx->stream));
    sum_i64_kernel<<<grid_for(ctx, n, 256, 4), 256, 0, ctx->stream>>>(num_els, n, ds, da);
    PB_TRY(check_launch(ctx, "sum_i64"));
    PB_CUDA(cudaMemcpyAsync(ctx->pinned_small, ds, 16, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    *sum = ((int64_t *)ctx->pinned_small)[0];
    *abs_sum = ((uint64_t *)ctx->pinned_small)[1];
    return PB_OK;
}

}  // namespace pb

extern "C" {

int pb_cms_create(pb_ctx *ctx, uint32_t width, uint32_t depth, pb_cms **out) {
    PB_REQUIRE(ctx && out, "NULL argument");
    PB_REQUIRE(width >= 1 && depth >= 1, "width and depth must be >= 1");
    DeviceGuard g(ctx->device);
    pb_cms *c = new (std::nothrow) pb_cms();
    if (!c) return PB_ERR_OOM;
    c->ctx = ctx;
    c->width = width;
    c->depth = depth;
    c->count = (uint64_t)width * depth;
    c->fm = make_fastmod(width);
    cudaError_t e = cudaMalloc(&c->bins, c->count * 4);
    if (e != cudaSuccess) {
        set_error("cudaMalloc of %llu counter bytes failed: %s", (unsigned long long)(c->count * 4), cudaGetErrorString(e));
        delete c;
        return PB_ERR_OOM;
    }
    cudaMemsetAsync(c->bins, 0, c->count * 4, ctx->stream);
    *out = c;
    return PB_OK;
}

int pb_cms_destroy(pb_cms *c) {
    if (!c) return PB_OK;
    DeviceGuard g(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    cudaFree(c->bins);
    delete c;
    return PB_OK;
}

int pb_cms_clear(pb_cms *c) {
    PB_REQUIRE(c, "handle is NULL");
    DeviceGuard g(c->ctx->device);
    PB_CUDA(cudaMemsetAsync(c->bins, 0, c->count * 4, c->ctx->stream));
    c->abs_added = 0;
    return PB_OK;
}

int pb_cms_upload(pb_cms *c, const int32_t *bins, uint64_t count) {
    PB_REQUIRE(c && bins, "NULL argument");
    PB_REQUIRE(count == c->count, "expected %llu counters, got %llu", (unsigned long long)c->count, (unsigned long long)count);
    DeviceGuard g(c->ctx->device);
    PB_CUDA(cudaMemcpyAsync(c->bins, bins, count * 4, cudaMemcpyHostToDevice, c->ctx->stream));
    PB_CUDA(cudaStreamSynchronize(c->ctx->stream));
    // bound the content from the host copy: the largest |bin| bounds what any bin already holds
    uint64_t mx = 0;
    for (uint64_t i = 0; i < count; ++i) {
        uint64_t a = bins[i] < 0 ? (uint64_t)(-(int64_t)bins[i]) : (uint64_t)bins[i];
        mx = a > mx ? a : mx;
    }
    c->abs_added = mx;
    return PB_OK;
}

int pb_cms_download(pb_cms *c, int32_t *bins, uint64_t count) {
    PB_REQUIRE(c && bins, "NULL argument");
    PB_REQUIRE(count == c->count, "expected %llu counters, got %llu", (unsigned long long)c->count, (unsigned long long)count);
    DeviceGuard g(c->ctx->device);
    PB_CUDA(cudaMemcpyAsync(bins, c->bins, count * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    PB_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return PB_OK;
}

int pb_cms_device_ptr(pb_cms *c, void **out_dev, uint64_t *out_count) {
    PB_REQUIRE(c && out_dev, "NULL argument");
    *out_dev = c->bins;
    if (out_count) *out_count = c->count;
    return PB_OK;
}

int pb_cms_add_keys(pb_cms *c, const pb_keys *keys, const int64_t *num_els, int64_t scalar_num_els,
                    int64_t *elements_added_inout) {
    PB_REQUIRE(c && keys, "NULL argument");
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    PB_TRY(validate_keys(keys));
    if (keys->n == 0) return PB_OK;
    int64_t sum = 0;
    uint64_t abs_sum = 0;
    PB_TRY(batch_totals(ctx, num_els, keys->on_device != 0, scalar_num_els, keys->n, &sum, &abs_sum));
    CmsAddArgs a;
    a.c = c;
    a.num_els = num_els;
    a.num_els_on_device = keys->on_device != 0;
    a.scalar = scalar_num_els;
    const uint64_t bound = c->abs_added + abs_sum;
    a.safe = bound >= c->abs_added && bound <= (uint64_t)kI32Max;  // no bin can leave int32
    PB_TRY(for_each_chunk(ctx, keys, cms_add_chunk, &a));
    c->abs_added = bound >= c->abs_added ? bound : ~0ull;
    if (elements_added_inout) *elements_added_inout = sat_add_i64(*elements_added_inout, sum);
    return PB_OK;
}

int pb_cms_check_keys(pb_cms *c, const pb_keys *keys, int query_type, int64_t elements_added, int64_t *out, int out_on_device) {
    PB_REQUIRE(c && keys, "NULL argument");
    PB_REQUIRE(out || keys->n == 0, "out is NULL");
    PB_REQUIRE(query_type >= 0 && query_type <= 2, "query_type must be 0 (min), 1 (mean) or 2 (mean-min)");
    PB_REQUIRE(query_type != 2 || c->depth <= (uint32_t)kMaxDepthLocal, "mean-min supports depth <= %d", kMaxDepthLocal);
    PB_REQUIRE(query_type != 2 || c->width >= 2, "mean-min needs width >= 2 (the reference divides by width-1)");
    DeviceGuard g(c->ctx->device);
    CmsCheckArgs a;
    a.c = c;
    a.q.query_type = query_type;
    a.q.elements_added = elements_added;
    a.out_dev = out_on_device ? out : nullptr;
    a.out_host = out_on_device ? nullptr : out;
    PB_TRY(for_each_chunk(c->ctx, keys, cms_check_chunk, &a));
    if (!out_on_device) PB_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return PB_OK;
}

int pb_cms_add_hashes(pb_cms *c, const uint64_t *hashes, uint64_t n, int on_device, const int64_t *num_els,
                      int64_t scalar_num_els, int64_t *elements_added_inout) {
    PB_REQUIRE(c && (hashes || n == 0), "NULL argument");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    int64_t sum = 0;
    uint64_t abs_sum = 0;
    PB_TRY(batch_totals(ctx, num_els, on_device != 0, scalar_num_els, n, &sum, &abs_sum));
    const uint64_t *dh = hashes;
    const int64_t *dn = num_els;
    if (!on_device) {
        PB_TRY(scratch_reserve(ctx, ctx->aux_stage[0], n * c->depth * 8));
        PB_CUDA(cudaMemcpyAsync(ctx->aux_stage[0].p, hashes, n * c->depth * 8, cudaMemcpyHostToDevice, ctx->stream));
        dh = (const uint64_t *)ctx->aux_stage[0].p;
        if (num_els) {
            PB_TRY(scratch_reserve(ctx, ctx->aux_stage[1], n * 8));
            PB_CUDA(cudaMemcpyAsync(ctx->aux_stage[1].p, num_els, n * 8, cudaMemcpyHostToDevice, ctx->stream));
            dn = (const int64_t *)ctx->aux_stage[1].p;
        }
    }
    const uint64_t bound = c->abs_added + abs_sum;
    const bool safe = bound >= c->abs_added && bound <= (uint64_t)kI32Max;
    const int grid = grid_for(ctx, n, 256, 8);
    if (safe) cms_add_hashes_kernel<true><<<grid, 256, 0, ctx->stream>>>(dh, n, dn, scalar_num_els, dev_view(c));
    else cms_add_hashes_kernel<false><<<grid, 256, 0, ctx->stream>>>(dh, n, dn, scalar_num_els, dev_view(c));
    PB_TRY(check_launch(ctx, "cms_add_hashes"));
    if (!on_device) PB_CUDA(cudaStreamSynchronize(ctx->stream));
    c->abs_added = bound >= c->abs_added ? bound : ~0ull;
    if (elements_added_inout) *elements_added_inout = sat_add_i64(*elements_added_inout, sum);
    return PB_OK;
}

int pb_cms_check_hashes(pb_cms *c, const uint64_t *hashes, uint64_t n, int on_device, int query_type, int64_t elements_added,
                        int64_t *out, int out_on_device) {
    PB_REQUIRE(c && ((hashes && out) || n == 0), "NULL argument");
    PB_REQUIRE(query_type >= 0 && query_type <= 2, "query_type must be 0 (min), 1 (mean) or 2 (mean-min)");
    PB_REQUIRE(query_type != 2 || c->depth <= (uint32_t)kMaxDepthLocal, "mean-min supports depth <= %d", kMaxDepthLocal);
    PB_REQUIRE(query_type != 2 || c->width >= 2, "mean-min needs width >= 2 (the reference divides by width-1)");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    const uint64_t *dh = hashes;
    if (!on_device) {
        PB_TRY(scratch_reserve(ctx, ctx->aux_stage[0], n * c->depth * 8));
        PB_CUDA(cudaMemcpyAsync(ctx->aux_stage[0].p, hashes, n * c->depth * 8, cudaMemcpyHostToDevice, ctx->stream));
        dh = (const uint64_t *)ctx->aux_stage[0].p;
    }
    int64_t *o = out;
    if (!out_on_device) {
        PB_TRY(scratch_reserve(ctx, ctx->out_stage[0], n * 8));
        o = (int64_t *)ctx->out_stage[0].p;
    }
    CmsQuery q{query_type, elements_added};
    cms_check_hashes_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(dh, n, dev_view(c), q, o);
    PB_TRY(check_launch(ctx, "cms_check_hashes"));
    if (!out_on_device) {
        PB_CUDA(cudaMemcpyAsync(out, o, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}

int pb_cms_join_buffer(pb_cms *c, const int32_t *other_dev, uint64_t count) {
    PB_REQUIRE(c && other_dev, "NULL argument");
    PB_REQUIRE(count == c->count, "expected %llu counters, got %llu", (unsigned long long)c->count, (unsigned long long)count);
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    cms_join_kernel<<<grid_for(ctx, count, 256, 8), 256, 0, ctx->stream>>>(c->bins, other_dev, count);
    PB_TRY(check_launch(ctx, "cms_join"));
    c->abs_added = ~0ull;  // unknown from here on: take the careful add path
    return PB_OK;
}

/* multi-GPU merge helpers: the table widened to int64 (input of one all-reduce), and a table loaded from int64
 * sums with the reference's saturation */
int pb_cms_widen(pb_cms *c, int64_t *out_dev, uint64_t count) {
    PB_REQUIRE(c && out_dev, "NULL argument");
    PB_REQUIRE(count == c->count, "expected %llu counters, got %llu", (unsigned long long)c->count, (unsigned long long)count);
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    cms_widen_kernel<<<grid_for(ctx, count, 256, 8), 256, 0, ctx->stream>>>(c->bins, (long long *)out_dev, count);
    return check_launch(ctx, "cms_widen");
}

int pb_cms_load_sums(pb_cms *c, const int64_t *sums_dev, uint64_t count) {
    PB_REQUIRE(c && sums_dev, "NULL argument");
    PB_REQUIRE(count == c->count, "expected %llu counters, got %llu", (unsigned long long)c->count, (unsigned long long)count);
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    cms_narrow_kernel<<<grid_for(ctx, count, 256, 8), 256, 0, ctx->stream>>>((const long long *)sums_dev, c->bins, count);
    PB_TRY(check_launch(ctx, "cms_load_sums"));
    c->abs_added = ~0ull;  // unknown from here on: take the careful add path
    return PB_OK;
}

}  // extern "C"
