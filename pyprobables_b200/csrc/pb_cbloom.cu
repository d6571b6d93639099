// pb_cbloom.cu -- CountingBloomFilter.add / check / remove for whole batches
// (reference: probables/blooms/countingbloom.py:125-208).
//
// State: uint32[num_counters] (the reference's array('I'), bloom_length == number_bits, countingbloom.py:77-78);
// counter of hash i = h_i % num_counters (:143).
//
// add (:143-153): every (key, hash) pair adds num_els to its counter -- a key whose hashes collide on one counter
// increments it once per colliding hash (the reference's "this will increment indices each time it is viewed"
// quirk), which is simply what one atomic per pair does.  Counters saturate at UINT32_MAX (:147-149); for
// non-negative adds the saturating sum is order-free, so the batch runs as concurrent atomics: plain RED.ADD while
// the handle can prove no counter can reach 2^32 (upper bound of everything ever added), an exact saturating CAS
// otherwise.
//
// remove (:186-208) is order-dependent in general: a key removes min(num_els, its smallest counter) and skips
// saturated counters.  Batch semantics = the sequential loop, obtained in two steps:
//   1. optimistic: every pair does atomicSub(num_els); a pair that finds its counter below num_els or saturated raises
//      a flag.  If no flag was raised, every counter held at least the total the batch takes from it, so in ANY
//      sequential order every key sees min_val >= num_els, removes exactly num_els per pair, and the result is
//      what the atomics produced;
//   2. otherwise the subtraction is undone (atomicAdd of the same amounts: arithmetic mod 2^32 is a group) and one
//      thread replays the chunk in key order with the reference's exact rules.
#include <algorithm>
#include <new>

#include "pb_common.cuh"
#include "pb_hash.cuh"
#include "pb_keys.cuh"
#include "pb_bloom_part.cuh"

using namespace pb;

struct pb_cbloom {
    pb_ctx *ctx = nullptr;
    uint64_t n_counters = 0;
    uint32_t k = 0;
    uint32_t *counts = nullptr;
    uint64_t alloc = 0;        // counters allocated (multiple of 4)
    uint64_t added_bound = 0;  // upper bound of what any single counter can hold (decides the RED path)
    FastMod fm;
};

namespace pb {

constexpr uint32_t kU32Max = 0xFFFFFFFFu;

struct CbDev {
    uint32_t *counts;
    FastMod fm;
    uint32_t k;
};

// countingbloom.py:145-151 for one counter, concurrent-safe
template <bool SAFE>
__device__ __forceinline__ void cb_add(uint32_t *c, uint32_t n) {
    if (SAFE) {
        atomicAdd(c, n);  // RED.ADD
        return;
    }
    uint32_t old = *reinterpret_cast<volatile uint32_t *>(c);
    while (old != kU32Max) {
        const uint32_t v = old > kU32Max - n ? kU32Max : old + n;
        const uint32_t prev = atomicCAS(c, old, v);
        if (prev == old) return;
        old = prev;
    }
}

template <int KG, bool SAFE>
__global__ void __launch_bounds__(256) cbloom_add_fixed16(const uint4 *__restrict__ keys, uint64_t n, uint32_t add, CbDev d) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 w = __ldcs(keys + i);
        for (uint32_t s0 = 0; s0 < d.k; s0 += KG) {
            uint64_t h[KG];
            fnv_group_16<KG>(w, s0, h);
#pragma unroll
            for (int j = 0; j < KG; ++j)
                if (s0 + j < d.k) cb_add<SAFE>(d.counts + fastmod(h[j], d.fm), add);
        }
    }
}

template <bool SAFE>
__global__ void __launch_bounds__(256) cbloom_add_hashes(const uint64_t *__restrict__ h, uint64_t total, uint32_t add, CbDev d) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
        cb_add<SAFE>(d.counts + fastmod(__ldcs(h + i), d.fm), add);
}

// countingbloom.py:164-174: min over the key's counters
template <int KG>
__global__ void __launch_bounds__(256) cbloom_check_fixed16(const uint4 *__restrict__ keys, uint64_t n, CbDev d, uint32_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 w = __ldcs(keys + i);
        uint32_t mn = kU32Max;
        for (uint32_t s0 = 0; s0 < d.k; s0 += KG) {
            uint64_t h[KG];
            uint32_t v[KG];
            fnv_group_16<KG>(w, s0, h);
#pragma unroll
            for (int j = 0; j < KG; ++j) v[j] = (s0 + j < d.k) ? __ldg(d.counts + fastmod(h[j], d.fm)) : kU32Max;
#pragma unroll
            for (int j = 0; j < KG; ++j) mn = v[j] < mn ? v[j] : mn;
        }
        out[i] = mn;
    }
}

__global__ void __launch_bounds__(256) cbloom_check_hashes(const uint64_t *__restrict__ h, uint64_t n, CbDev d, uint32_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t mn = kU32Max;
        for (uint32_t s = 0; s < d.k; ++s) {
            const uint32_t v = __ldg(d.counts + fastmod(h[i * d.k + s], d.fm));
            mn = v < mn ? v : mn;
        }
        out[i] = mn;
    }
}

// remove, step 1 (optimistic) and its undo
__global__ void __launch_bounds__(256) cbloom_sub_hashes(const uint64_t *__restrict__ h, uint64_t total, uint32_t num, CbDev d, unsigned int *flag) {
    bool bad = false;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t old = atomicSub(d.counts + fastmod(__ldcs(h + i), d.fm), num);
        bad |= old < num || old == kU32Max;
    }
    if (bad) *flag = 1u;
}
__global__ void __launch_bounds__(256) cbloom_undo_hashes(const uint64_t *__restrict__ h, uint64_t total, uint32_t num, CbDev d) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(d.counts + fastmod(__ldcs(h + i), d.fm), num);
}
// remove, step 2: one thread, key order, countingbloom.py:196-208 verbatim
__global__ void cbloom_remove_serial(const uint64_t *__restrict__ h, uint64_t n, uint32_t num, CbDev d, unsigned long long *removed) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned long long total = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t mn = kU32Max;
        for (uint32_t s = 0; s < d.k; ++s) {
            const uint32_t v = d.counts[fastmod(h[i * d.k + s], d.fm)];
            mn = v < mn ? v : mn;
        }
        if (mn == kU32Max || mn == 0) continue;  // :199-202
        const uint32_t take = mn > num ? num : mn;  // :204
        for (uint32_t s = 0; s < d.k; ++s) {
            uint32_t *c = d.counts + fastmod(h[i * d.k + s], d.fm);
            const uint32_t v = *c;
            if (v < kU32Max) *c = v >= take ? v - take : 0u;  // :205-207 (the reference raises OverflowError below zero)
        }
        total += take;
    }
    *removed = total;
}

// whole-table passes: op 0 union (:300-326 sum), op 1 intersection (:210-243 sum where both are non-zero)
__global__ void __launch_bounds__(256) cbloom_combine_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ a,
                                                             const uint32_t *__restrict__ b, uint64_t n, int op) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = a[i], y = b[i];
        const uint32_t s = x > kU32Max - y ? kU32Max : x + y;
        dst[i] = (op == 0 || (x > 0 && y > 0)) ? s : 0u;
    }
}
// out[0] = counters non-zero in a or b, out[1] = non-zero in both (jaccard_index, :245-272)
__global__ void __launch_bounds__(256) cbloom_pair_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, uint64_t n,
                                                          unsigned long long *out) {
    unsigned long long cu = 0, ci = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = a[i], y = b[i];
        cu += (x | y) != 0;
        ci += (x != 0) & (y != 0);
    }
    for (int o = 16; o; o >>= 1) {
        cu += __shfl_xor_sync(0xffffffffu, cu, o);
        ci += __shfl_xor_sync(0xffffffffu, ci, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cu) atomicAdd(out, cu);
        if (ci) atomicAdd(out + 1, ci);
    }
}
// out[0] = non-zero counters, out[1] = sum, out[2] = max << 40-bit-free packing is avoided: max and its first index
// come from a second tiny pass (out[2] = max, out[3] = smallest index holding it) -- __str__ (:100-123)
__global__ void __launch_bounds__(256) cbloom_stats_kernel(const uint32_t *__restrict__ c, uint64_t n, unsigned long long *out) {
    unsigned long long nz = 0, sum = 0;
    unsigned int mx = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t v = c[i];
        nz += v != 0;
        sum += v;
        mx = v > mx ? v : mx;
    }
    for (int o = 16; o; o >>= 1) {
        nz += __shfl_xor_sync(0xffffffffu, nz, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const unsigned int m2 = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = m2 > mx ? m2 : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        if (nz) atomicAdd(out, nz);
        if (sum) atomicAdd(out + 1, sum);
        atomicMax(out + 2, (unsigned long long)mx);
    }
}
__global__ void __launch_bounds__(256) cbloom_argmax_kernel(const uint32_t *__restrict__ c, uint64_t n, unsigned long long *out) {
    const unsigned long long mx = out[2];
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (c[i] == (uint32_t)mx) atomicMin(out + 3, (unsigned long long)i);
}

static CbDev dev_view(const pb_cbloom *b) {
    CbDev d;
    d.counts = b->counts;
    d.fm = b->fm;
    d.k = b->k;
    return d;
}

struct CbArgs {
    pb_cbloom *b;
    int op;  // 0 add, 1 check, 2 remove
    uint32_t num;
    bool safe;
    uint32_t *out_dev;
    uint32_t *out_host;
    uint64_t removed;
};

#define PB_CB_KG(kg, CALL)      \
    switch (kg) {               \
        case 1: CALL(1); break; \
        case 2: CALL(2); break; \
        case 3: CALL(3); break; \
        case 4: CALL(4); break; \
        case 5: CALL(5); break; \
        case 6: CALL(6); break; \
        case 7: CALL(7); break; \
        default: CALL(8); break; \
    }

// remove for n hash rows already on the device (stream-ordered up to the flag read-back)
static int remove_rows(pb_ctx *ctx, pb_cbloom *b, const uint64_t *rows, uint64_t n, uint32_t num, uint64_t *removed) {
    const CbDev d = dev_view(b);
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    unsigned int *flag = (unsigned int *)((unsigned long long *)ctx->small.p + 64);
    unsigned long long *rem = (unsigned long long *)ctx->small.p + 66;
    PB_CUDA(cudaMemsetAsync(flag, 0, 4, ctx->stream));
    const uint64_t total = n * b->k;
    launch_begin(ctx);
    cbloom_sub_hashes<<<grid_for(ctx, total, 256, 8), 256, 0, ctx->stream>>>(rows, total, num, d, flag);
    PB_TRY(check_launch(ctx, "cbloom_remove"));
    unsigned int *hflag = (unsigned int *)((uint8_t *)ctx->pinned_small + 1024);
    PB_CUDA(cudaMemcpyAsync(hflag, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*hflag == 0) {
        *removed += n * (uint64_t)num;
        return PB_OK;
    }
    cbloom_undo_hashes<<<grid_for(ctx, total, 256, 8), 256, 0, ctx->stream>>>(rows, total, num, d);
    PB_TRY(check_launch(ctx, "cbloom_remove_undo"));
    cbloom_remove_serial<<<1, 32, 0, ctx->stream>>>(rows, n, num, d, rem);
    PB_TRY(check_launch(ctx, "cbloom_remove_serial"));
    unsigned long long *hrem = (unsigned long long *)((uint8_t *)ctx->pinned_small + 1032);
    PB_CUDA(cudaMemcpyAsync(hrem, rem, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    *removed += *hrem;
    return PB_OK;
}

static int cb_chunk(pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) {
    CbArgs *a = (CbArgs *)user;
    pb_cbloom *b = a->b;
    const CbDev d = dev_view(b);
    const int kg = pick_group(b->k);
    const bool f16 = is_fixed16(dk);
    const int grid = grid_for(ctx, dk.n, 256, 8);
    const uint4 *k4 = (const uint4 *)dk.data;
    uint32_t *out = nullptr;
    if (a->op == 1) {
        out = a->out_dev ? a->out_dev + first : nullptr;
        if (!out) {
            PB_TRY(scratch_reserve(ctx, ctx->out_stage[slot], dk.n * 4));
            out = (uint32_t *)ctx->out_stage[slot].p;
        }
    }
    if (f16 && a->op != 2) {
        launch_begin(ctx);
        if (a->op == 0) {
#define CALL(K)                                                                                   \
    do {                                                                                          \
        if (a->safe) cbloom_add_fixed16<K, true><<<grid, 256, 0, ctx->stream>>>(k4, dk.n, a->num, d); \
        else cbloom_add_fixed16<K, false><<<grid, 256, 0, ctx->stream>>>(k4, dk.n, a->num, d);    \
    } while (0)
            PB_CB_KG(kg, CALL)
#undef CALL
            PB_TRY(check_launch(ctx, "cbloom_add"));
        } else {
#define CALL(K) cbloom_check_fixed16<K><<<grid, 256, 0, ctx->stream>>>(k4, dk.n, d, out)
            PB_CB_KG(kg, CALL)
#undef CALL
            PB_TRY(check_launch(ctx, "cbloom_check"));
        }
    } else {
        // every other key layout (and remove): hash rows first (pb_hash_keys kernels), then the pre-hashed kernels
        PB_TRY(scratch_reserve(ctx, ctx->aux_stage[slot], dk.n * b->k * 8));
        uint64_t *rows = (uint64_t *)ctx->aux_stage[slot].p;
        PB_TRY(hash_dev_keys(ctx, dk, b->k, rows));
        const uint64_t total = dk.n * b->k;
        const int g2 = grid_for(ctx, total, 256, 8);
        if (a->op == 0) {
            launch_begin(ctx);
            if (a->safe) cbloom_add_hashes<true><<<g2, 256, 0, ctx->stream>>>(rows, total, a->num, d);
            else cbloom_add_hashes<false><<<g2, 256, 0, ctx->stream>>>(rows, total, a->num, d);
            PB_TRY(check_launch(ctx, "cbloom_add"));
        } else if (a->op == 1) {
            launch_begin(ctx);
            cbloom_check_hashes<<<grid, 256, 0, ctx->stream>>>(rows, dk.n, d, out);
            PB_TRY(check_launch(ctx, "cbloom_check"));
        } else {
            PB_TRY(remove_rows(ctx, b, rows, dk.n, a->num, &a->removed));
        }
    }
    if (a->op == 1 && a->out_host)
        PB_CUDA(cudaMemcpyAsync(a->out_host + first, out, dk.n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return PB_OK;
}

// the counter bound after adding n keys x num: each key can hit one counter up to k times
static bool note_add(pb_cbloom *b, uint64_t n, uint64_t num) {
    const unsigned __int128 add = (unsigned __int128)n * num * b->k;
    const unsigned __int128 nb = (unsigned __int128)b->added_bound + add;
    b->added_bound = nb > (unsigned __int128)0xFFFFFFFFFFFFFFFFull ? 0xFFFFFFFFFFFFFFFFull : (uint64_t)nb;
    return b->added_bound <= (uint64_t)kU32Max;
}

// ---- partitioned add (the Bloom insert's two passes, pb_bloom_part.cuh, on a counter array) ------------------------
// A random RED.ADD on a DRAM-resident counter array costs a sector read-modify-write (21.7 G/s measured).  Pass 1 bins
// the counter indices of a chunk by window of 2^wl counters (16 MB: L2 resident), pass 2 streams each window's lists and
// adds in L2.  Only while the handle can prove that no counter reaches 2^32 (plain RED.ADD, no saturation), for k <= 16.
__global__ void __launch_bounds__(256) cbloom_apply_overflow(const uint64_t *__restrict__ list, const unsigned long long *count, uint64_t cap,
                                                             uint32_t *__restrict__ counters, uint32_t amount) {
    const uint64_t n = *count < cap ? *count : cap;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(counters + list[i], amount);
}

struct CbPartPlan {
    bool use = false;
    uint32_t window_log2 = 0, n_windows = 0;
    uint64_t chunk_keys = 0;
    PartLayout lay{};
};

static CbPartPlan plan_cb_partition(const pb_cbloom *b, const pb_keys *keys, bool safe) {
    CbPartPlan pl;
    pb_ctx *ctx = b->ctx;
    const int64_t mode = ctx->bloom_insert_mode;  // 0 auto, 1 direct, 2 partitioned (shared with the Bloom insert)
    const uint64_t n = keys->n;
    if (mode == 1 || !safe || b->k > kMaxPartK || n == 0) return pl;
    const uint64_t bytes = b->n_counters * 4;
    const uint64_t l2 = ctx->l2_bytes ? ctx->l2_bytes : ((uint64_t)96 << 20);
    if (mode == 0) {
        if (bytes <= l2) return pl;                                        // the direct REDs hit L2 anyway
        if ((double)n * b->k * 64.0 < 4.0 * (double)bytes) return pl;      // one read+write of the array per chunk must pay
    }
    // the same window BYTES as the Bloom insert: 2^(bits-5) counters
    uint32_t wl = (uint32_t)std::max<int64_t>(5, std::min<int64_t>(ctx->bloom_window_log2_bits - 5, 26));
    while (((b->n_counters + ((1ull << wl) - 1)) >> wl) > (uint64_t)kMaxWindows2 && wl < 26) ++wl;
    const uint64_t nw = (b->n_counters + ((1ull << wl) - 1)) >> wl;
    if (nw > (uint64_t)kMaxWindows2) return pl;
    const bool fixed16 = keys->offsets == nullptr && keys->sym_width == 1 && keys->stride == 16;
    const uint64_t budget_entries = std::min<uint64_t>((uint64_t)ctx->stage_bytes / 4, 0xFFFFFFF0ull);
    uint64_t chunk = std::min<uint64_t>(n, 1ull << 27);
    PartLayout lay;
    for (;;) {
        lay = part_layout(ctx, chunk, b->k, b->n_counters, wl, (uint32_t)nw, fixed16, false);
        if ((uint64_t)lay.sub_cap * (uint64_t)lay.grid * nw <= budget_entries) break;
        if (chunk <= 4096) return pl;
        chunk = chunk - chunk / 4;
    }
    pl.use = true;
    pl.window_log2 = wl;
    pl.n_windows = (uint32_t)nw;
    pl.chunk_keys = chunk;
    pl.lay = lay;
    return pl;
}

struct CbPartArgs {
    pb_cbloom *b;
    CbPartPlan plan;
    uint32_t num;
};

static int cb_part_chunk(pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) {
    (void)first;
    (void)slot;
    CbPartArgs *a = (CbPartArgs *)user;
    pb_cbloom *b = a->b;
    const CbPartPlan &pl = a->plan;
    const size_t n_lists = (size_t)pl.n_windows * (size_t)pl.lay.grid;
    PB_TRY(scratch_reserve(ctx, ctx->part_stage, n_lists * pl.lay.sub_cap * 4));
    PB_TRY(scratch_reserve(ctx, ctx->part_cursors, n_lists * 4));
    PartDev pd;
    pd.stage = (uint32_t *)ctx->part_stage.p;
    pd.counts = (uint32_t *)ctx->part_cursors.p;
    pd.words = b->counts;
    part_set_modulus(pd, b->n_counters);
    pd.sub_cap = pl.lay.sub_cap;
    pd.n_sub = (uint32_t)pl.lay.grid;
    pd.window_log2 = pl.window_log2;
    pd.n_windows = pl.n_windows;
    pd.k = b->k;
    // indices that do not fit their sublist (7 sigma: practically never) go to a side list and are added one by one
    constexpr uint64_t kOvfCap = 1ull << 20;
    PB_TRY(scratch_reserve(ctx, ctx->aux_stage[1], kOvfCap * 8));
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    unsigned long long *ovf_count = (unsigned long long *)ctx->small.p + 48;
    PB_CUDA(cudaMemsetAsync(ovf_count, 0, 8, ctx->stream));
    pd.ovf_list = (uint64_t *)ctx->aux_stage[1].p;
    pd.ovf_count = ovf_count;
    pd.ovf_cap = kOvfCap;
    launch_begin(ctx);
    cudaError_t e = launch_part4(pl.lay.block, ctx->stream, dk, pd);
    if (e != cudaSuccess) {
        set_error("launch of bloom_part4 (counting bloom, k=%u) failed: %s", pd.k, cudaGetErrorString(e));
        return PB_ERR_CUDA;
    }
    PB_TRY(check_launch(ctx, "cbloom_part"));
    const uint32_t cpw = (uint32_t)ctx->num_sms * (uint32_t)std::max<int64_t>(1, std::min<int64_t>(ctx->bloom_apply_cpw_per_sm, 32));
    launch_begin(ctx);
    cbloom_apply2<<<pl.n_windows * cpw, 256, 0, ctx->stream>>>(pd, cpw, a->num);
    PB_TRY(check_launch(ctx, "cbloom_apply_windows"));
    cbloom_apply_overflow<<<64, 256, 0, ctx->stream>>>(pd.ovf_list, ovf_count, kOvfCap, b->counts, a->num);
    PB_TRY(check_launch(ctx, "cbloom_apply_overflow"));
    PB_CUDA(cudaMemcpyAsync(ctx->pinned_small, ovf_count, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint64_t spilled = *(uint64_t *)ctx->pinned_small;
    PB_REQUIRE(spilled <= kOvfCap, "%llu counter indices overflowed their sublists (side list holds %llu): counters are incomplete",
               (unsigned long long)spilled, (unsigned long long)kOvfCap);
    return PB_OK;
}

static int stage_rows(pb_ctx *ctx, const uint64_t *hashes, uint64_t count, int on_device, const uint64_t **dev) {
    if (on_device) {
        *dev = hashes;
        return PB_OK;
    }
    PB_TRY(scratch_reserve(ctx, ctx->aux_stage[0], count * 8));
    PB_CUDA(cudaMemcpyAsync(ctx->aux_stage[0].p, hashes, count * 8, cudaMemcpyHostToDevice, ctx->stream));
    *dev = (const uint64_t *)ctx->aux_stage[0].p;
    return PB_OK;
}

}  // namespace pb

extern "C" {

int pb_cbloom_create(pb_ctx *ctx, uint64_t num_counters, uint32_t k, pb_cbloom **out) {
    PB_REQUIRE(ctx && out, "NULL argument");
    PB_REQUIRE(num_counters >= 1 && k >= 1, "num_counters and k must be >= 1");
    DeviceGuard g(ctx->device);
    pb_cbloom *b = new (std::nothrow) pb_cbloom();
    if (!b) return PB_ERR_OOM;
    b->ctx = ctx;
    b->n_counters = num_counters;
    b->k = k;
    b->alloc = (num_counters + 3) & ~(uint64_t)3;
    b->fm = make_fastmod(num_counters);
    cudaError_t e = cudaMalloc(&b->counts, b->alloc * 4);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc of %llu counter bytes failed: %s", (unsigned long long)(b->alloc * 4), cudaGetErrorString(e));
        delete b;
        return PB_ERR_OOM;
    }
    e = cudaMemsetAsync(b->counts, 0, b->alloc * 4, ctx->stream);
    if (e != cudaSuccess) {
        set_error("memset failed: %s", cudaGetErrorString(e));
        cudaFree(b->counts);
        delete b;
        return PB_ERR_CUDA;
    }
    *out = b;
    return PB_OK;
}

int pb_cbloom_destroy(pb_cbloom *b) {
    if (!b) return PB_OK;
    DeviceGuard g(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    cudaFree(b->counts);
    delete b;
    return PB_OK;
}

int pb_cbloom_clear(pb_cbloom *b) {
    PB_REQUIRE(b, "handle is NULL");
    DeviceGuard g(b->ctx->device);
    PB_CUDA(cudaMemsetAsync(b->counts, 0, b->alloc * 4, b->ctx->stream));
    b->added_bound = 0;
    return PB_OK;
}

int pb_cbloom_upload(pb_cbloom *b, const uint32_t *counts, uint64_t n) {
    PB_REQUIRE(b && counts, "NULL argument");
    PB_REQUIRE(n == b->n_counters, "expected %llu counters, got %llu", (unsigned long long)b->n_counters, (unsigned long long)n);
    DeviceGuard g(b->ctx->device);
    PB_CUDA(cudaMemcpyAsync(b->counts, counts, n * 4, cudaMemcpyHostToDevice, b->ctx->stream));
    PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    b->added_bound = 0xFFFFFFFFFFFFFFFFull;  // unknown contents: adds take the exact saturating path
    return PB_OK;
}

int pb_cbloom_download(pb_cbloom *b, uint32_t *counts, uint64_t n) {
    PB_REQUIRE(b && counts, "NULL argument");
    PB_REQUIRE(n == b->n_counters, "expected %llu counters, got %llu", (unsigned long long)b->n_counters, (unsigned long long)n);
    DeviceGuard g(b->ctx->device);
    PB_CUDA(cudaMemcpyAsync(counts, b->counts, n * 4, cudaMemcpyDeviceToHost, b->ctx->stream));
    PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    return PB_OK;
}

int pb_cbloom_device_ptr(pb_cbloom *b, void **out_dev, uint64_t *out_count) {
    PB_REQUIRE(b && out_dev, "NULL argument");
    *out_dev = b->counts;
    if (out_count) *out_count = b->n_counters;
    return PB_OK;
}

int pb_cbloom_add_keys(pb_cbloom *b, const pb_keys *keys, uint64_t num_els) {
    PB_REQUIRE(b && keys, "NULL argument");
    PB_REQUIRE(num_els <= (uint64_t)kU32Max, "num_els must fit 32 bits");
    DeviceGuard g(b->ctx->device);
    if (num_els == 0 || keys->n == 0) return validate_keys(keys);
    const bool safe = note_add(b, keys->n, num_els);
    CbPartPlan pl = plan_cb_partition(b, keys, safe);
    if (pl.use) {
        CbPartArgs pa{b, pl, (uint32_t)num_els};
        return for_each_chunk(b->ctx, keys, cb_part_chunk, &pa, pl.chunk_keys);
    }
    CbArgs a{b, 0, (uint32_t)num_els, safe, nullptr, nullptr, 0};
    return for_each_chunk(b->ctx, keys, cb_chunk, &a);
}

int pb_cbloom_add_hashes(pb_cbloom *b, const uint64_t *hashes, uint64_t n, int on_device, uint64_t num_els) {
    PB_REQUIRE(b && (hashes || n == 0), "NULL argument");
    PB_REQUIRE(num_els <= (uint64_t)kU32Max, "num_els must fit 32 bits");
    if (n == 0 || num_els == 0) return PB_OK;
    pb_ctx *ctx = b->ctx;
    DeviceGuard g(ctx->device);
    const uint64_t *d;
    PB_TRY(stage_rows(ctx, hashes, n * b->k, on_device, &d));
    const bool safe = note_add(b, n, num_els);
    const uint64_t total = n * b->k;
    launch_begin(ctx);
    if (safe) cbloom_add_hashes<true><<<grid_for(ctx, total, 256, 8), 256, 0, ctx->stream>>>(d, total, (uint32_t)num_els, dev_view(b));
    else cbloom_add_hashes<false><<<grid_for(ctx, total, 256, 8), 256, 0, ctx->stream>>>(d, total, (uint32_t)num_els, dev_view(b));
    PB_TRY(check_launch(ctx, "cbloom_add"));
    if (!on_device) PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

int pb_cbloom_check_keys(pb_cbloom *b, const pb_keys *keys, uint32_t *out, int out_on_device) {
    PB_REQUIRE(b && keys, "NULL argument");
    PB_REQUIRE(out || keys->n == 0, "out is NULL");
    DeviceGuard g(b->ctx->device);
    CbArgs a{b, 1, 0, true, out_on_device ? out : nullptr, out_on_device ? nullptr : out, 0};
    PB_TRY(for_each_chunk(b->ctx, keys, cb_chunk, &a));
    if (!out_on_device) PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    return PB_OK;
}

int pb_cbloom_check_hashes(pb_cbloom *b, const uint64_t *hashes, uint64_t n, int on_device, uint32_t *out, int out_on_device) {
    PB_REQUIRE(b && ((hashes && out) || n == 0), "NULL argument");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = b->ctx;
    DeviceGuard g(ctx->device);
    const uint64_t *d;
    PB_TRY(stage_rows(ctx, hashes, n * b->k, on_device, &d));
    uint32_t *o = out;
    if (!out_on_device) {
        PB_TRY(scratch_reserve(ctx, ctx->out_stage[0], n * 4));
        o = (uint32_t *)ctx->out_stage[0].p;
    }
    launch_begin(ctx);
    cbloom_check_hashes<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(d, n, dev_view(b), o);
    PB_TRY(check_launch(ctx, "cbloom_check"));
    if (!out_on_device) {
        PB_CUDA(cudaMemcpyAsync(out, o, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}

int pb_cbloom_remove_keys(pb_cbloom *b, const pb_keys *keys, uint64_t num_els, uint64_t *removed_total) {
    PB_REQUIRE(b && keys && removed_total, "NULL argument");
    PB_REQUIRE(num_els <= (uint64_t)kU32Max, "num_els must fit 32 bits");
    DeviceGuard g(b->ctx->device);
    *removed_total = 0;
    if (num_els == 0 || keys->n == 0) return validate_keys(keys);
    CbArgs a{b, 2, (uint32_t)num_els, true, nullptr, nullptr, 0};
    const int st = for_each_chunk(b->ctx, keys, cb_chunk, &a);
    *removed_total = a.removed;
    return st;
}

int pb_cbloom_remove_hashes(pb_cbloom *b, const uint64_t *hashes, uint64_t n, int on_device, uint64_t num_els, uint64_t *removed_total) {
    PB_REQUIRE(b && removed_total && (hashes || n == 0), "NULL argument");
    PB_REQUIRE(num_els <= (uint64_t)kU32Max, "num_els must fit 32 bits");
    *removed_total = 0;
    if (n == 0 || num_els == 0) return PB_OK;
    pb_ctx *ctx = b->ctx;
    DeviceGuard g(ctx->device);
    const uint64_t *d;
    PB_TRY(stage_rows(ctx, hashes, n * b->k, on_device, &d));
    return remove_rows(ctx, b, d, n, (uint32_t)num_els, removed_total);
}

/* out[0] = non-zero counters (_cnt_number_bits_set, :328-330), out[1] = sum of all counters, out[2] = largest
 * counter, out[3] = first index holding it (__str__, :100-123) */
int pb_cbloom_stats(pb_cbloom *b, uint64_t *out4) {
    PB_REQUIRE(b && out4, "NULL argument");
    pb_ctx *ctx = b->ctx;
    DeviceGuard g(ctx->device);
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    unsigned long long *acc = (unsigned long long *)ctx->small.p + 80;
    const unsigned long long init[4] = {0, 0, 0, ~0ull};
    PB_CUDA(cudaMemcpyAsync(acc, init, 32, cudaMemcpyHostToDevice, ctx->stream));
    const int grid = grid_for(ctx, b->n_counters, 256, 8);
    cbloom_stats_kernel<<<grid, 256, 0, ctx->stream>>>(b->counts, b->n_counters, acc);
    PB_TRY(check_launch(ctx, "cbloom_stats"));
    cbloom_argmax_kernel<<<grid, 256, 0, ctx->stream>>>(b->counts, b->n_counters, acc);
    PB_TRY(check_launch(ctx, "cbloom_stats"));
    uint64_t *h = (uint64_t *)((uint8_t *)ctx->pinned_small + 2048);
    PB_CUDA(cudaMemcpyAsync(h, acc, 32, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 4; ++i) out4[i] = h[i];
    return PB_OK;
}

static int cb_same_shape(const pb_cbloom *a, const pb_cbloom *b) {
    return a->n_counters == b->n_counters && a->ctx->device == b->ctx->device;
}

int pb_cbloom_combine(pb_cbloom *dst, pb_cbloom *a, pb_cbloom *b, int op) {
    PB_REQUIRE(dst && a && b, "NULL argument");
    PB_REQUIRE(op == 0 || op == 1, "op must be 0 (union) or 1 (intersection)");
    PB_REQUIRE(cb_same_shape(dst, a) && cb_same_shape(dst, b), "Counting Bloom Filters are not similar");
    pb_ctx *ctx = dst->ctx;
    DeviceGuard g(ctx->device);
    if (a->ctx != ctx) PB_CUDA(cudaStreamSynchronize(a->ctx->stream));
    if (b->ctx != ctx) PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    cbloom_combine_kernel<<<grid_for(ctx, dst->n_counters, 256, 8), 256, 0, ctx->stream>>>(dst->counts, a->counts, b->counts,
                                                                                          dst->n_counters, op);
    dst->added_bound = 0xFFFFFFFFFFFFFFFFull;
    return check_launch(ctx, "cbloom_combine");
}

int pb_cbloom_pair_counts(pb_cbloom *a, pb_cbloom *b, uint64_t *counts) {
    PB_REQUIRE(a && b && counts, "NULL argument");
    PB_REQUIRE(cb_same_shape(a, b), "Counting Bloom Filters are not similar");
    pb_ctx *ctx = a->ctx;
    DeviceGuard g(ctx->device);
    if (b->ctx != ctx) PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    unsigned long long *acc = (unsigned long long *)ctx->small.p + 90;
    PB_CUDA(cudaMemsetAsync(acc, 0, 16, ctx->stream));
    cbloom_pair_kernel<<<grid_for(ctx, a->n_counters, 256, 8), 256, 0, ctx->stream>>>(a->counts, b->counts, a->n_counters, acc);
    PB_TRY(check_launch(ctx, "cbloom_pair_counts"));
    uint64_t *h = (uint64_t *)((uint8_t *)ctx->pinned_small + 3072);
    PB_CUDA(cudaMemcpyAsync(h, acc, 16, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    counts[0] = h[0];
    counts[1] = h[1];
    return PB_OK;
}

}  // extern "C"
