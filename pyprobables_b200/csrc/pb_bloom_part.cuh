// pb_bloom_part.cuh -- the partitioned Bloom insert (pass 1 + pass 2), shared by the single-GPU path (pb_bloom.cu)
// and the multi-GPU exchange over peer memory (pb_p2p.cu).
//
// Pass 1 (bloom_part4, pb_bloom_part4.cu) hashes a chunk of keys, reduces % m and bins the bit indices by bitmap
// *window* (2^window_log2 bits, sized so that a window stays resident in B200's 126 MB L2) as window-local u32.
// Pass 2 (bloom_apply2 / bloom_apply_sources below) walks the windows in launch order and ORs each window's
// indices in with RED.OR while its slice of the bitmap is L2 resident -- the DRAM never sees a random access.
//
// Staging layout (round 2): every CTA of the pass-1 launch owns a private *sublist* of every window,
//     stage [n_windows][n_sub][sub_cap]  u32 window-local bit indices
//     counts[n_windows][n_sub]           entries CTA s wrote to window w
// so pass 1 needs no global atomics, no reservation quotas and no sentinels at all (round 1 reserved list ranges
// with one global atomic per ~1000 entries; the round-2 profile showed the tile loop waiting on that round trip:
// 20 % of all warp samples sat at the barrier behind it).  sub_cap is the expected share of a CTA plus 7 sigma; a
// sublist that still overflows (skewed or duplicated keys) falls back to an exact slow path, so the result is
// exact for any input.
#pragma once
#include <math.h>

#include "pb_common.cuh"
#include "pb_hash.cuh"
#include "pb_keys.cuh"

namespace pb {

constexpr int kMaxWindows2 = 512;
constexpr uint32_t kMaxPartK = 16;   // pass 1 is instantiated for 1..16 hashes (more take the direct path)
constexpr uint32_t kMaxQueryK = 15;  // ... the partitioned query for 1..15

struct PartDev {
    uint32_t *stage;         // [n_windows][n_sub][sub_cap]
    uint32_t *counts;        // [n_windows][n_sub]
    uint32_t *words;         // the bitmap (overflow fallback of the single-GPU path; nullptr when routing)
    // partitioned QUERY (bloom_part4<.., IDS = true> + bloom_probe2): every staged bit index carries the number of its
    // key inside the chunk; a bit that is not set clears the key's answer byte
    uint32_t *ids;           // [n_windows][n_sub][sub_cap], parallel to stage (nullptr for inserts)
    uint8_t *out;            // answers of the chunk's keys, preset to 1
    uint64_t m;              // number of bits (modulus)
    uint64_t recip;          // floor(2^64 / m)
    uint32_t sub_cap;        // entries per sublist (multiple of 4)
    uint32_t n_sub;          // sublists per window = CTAs of the pass-1 launch
    uint32_t window_log2;    // <= 31
    uint32_t n_windows;
    uint32_t k;
    uint32_t recip_fits32;   // m > 2^32
    // m >= 2^33 (every BASELINE-size filter): quotient < 2^31, so h - q*m is one signed mad.wide + one mad.lo
    uint32_t fast33;
    int32_t m_lo_s;          // low word of m read as a signed number
    uint32_t m_hi_adj;       // high word of m, plus one when the low word's sign bit is set
    // multi-GPU routing: the bitmap of most windows lives on another GPU, so an index that does not fit its
    // sublist cannot fall back to a local RED; it goes to this list of global bit indices instead
    uint64_t *ovf_list;
    unsigned long long *ovf_count;
    uint64_t ovf_cap;
};

__device__ __forceinline__ void part_overflow(const PartDev &p, uint64_t idx) {
    if (p.ovf_list) {
        const unsigned long long pos = atomicAdd(p.ovf_count, 1ull);
        if (pos < p.ovf_cap) p.ovf_list[pos] = idx;
    } else {
        atomicOr(p.words + (idx >> 5), 1u << (uint32_t)(idx & 31));
    }
}

// exact h % m with R = floor(2^64/m) < 2^32 (m > 2^32): q = floor(h*R / 2^64) is floor(h/m) or one less
__device__ __forceinline__ uint64_t mod_big(uint64_t h, uint64_t m, uint32_t r32) {
    const uint64_t t = (uint64_t)(uint32_t)h * r32;
    const uint64_t u = (h >> 32) * (uint64_t)r32 + (t >> 32);
    const uint32_t q = (uint32_t)(u >> 32);
    uint64_t r = h - (uint64_t)q * m;
    return r >= m ? r - m : r;
}

// modulus-derived fields of PartDev (host)
inline void part_set_modulus(PartDev &pd, uint64_t m) {
    pd.m = m;
    pd.recip = make_fastmod(m).recip;
    pd.recip_fits32 = m > (1ull << 32) ? 1u : 0u;
    pd.fast33 = m >= (1ull << 33) ? 1u : 0u;
    pd.m_lo_s = (int32_t)(uint32_t)m;
    pd.m_hi_adj = (uint32_t)(m >> 32) + (uint32_t)((m >> 31) & 1u);
}

// exact h % m for m >= 2^33.  R = floor(2^64/m) < 2^31; q = floor(h*R / 2^64) is floor(h/m) or one less and < 2^31,
// so -q is a signed 32-bit number: r = h - q*m = h + (-q)*m_lo_s (signed 32x32+64) with the high word corrected by
// (-q)*m_hi_adj, then one conditional subtract.  Host twin (same steps in portable C): pbt_mod_fast33 in pb_ctx.cu.
__device__ __forceinline__ uint64_t mod_fast33(uint64_t h, const PartDev &p) {
    const uint32_t r32 = (uint32_t)p.recip;
    const uint32_t t_hi = __umulhi((uint32_t)h, r32);
    uint64_t u;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(u) : "r"((uint32_t)(h >> 32)), "r"(r32), "l"((uint64_t)t_hi));
    const int32_t nq = -(int32_t)(uint32_t)(u >> 32);
    int64_t r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(nq), "r"(p.m_lo_s), "l"((int64_t)h));
    uint32_t r_hi;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r_hi) : "r"((uint32_t)nq), "r"(p.m_hi_adj), "r"((uint32_t)((uint64_t)r >> 32)));
    const uint64_t rr = ((uint64_t)r_hi << 32) | (uint32_t)r;
    return rr >= p.m ? rr - p.m : rr;
}

__device__ __forceinline__ uint64_t mod_any(uint64_t h, const PartDev &p) {
    if (p.recip_fits32) return mod_big(h, p.m, (uint32_t)p.recip);
    const uint64_t q = __umul64hi(h, p.recip);
    const uint64_t r = h - q * p.m;
    return r >= p.m ? r - p.m : r;
}

// ---- layout of one pass-1 launch (host) -----------------------------------------------------------------------
struct PartLayout {
    int grid;          // CTAs of the launch = sublists per window
    int block;         // 256, or 512 (twice the entries per window and tile) when there are many windows
    uint32_t sub_cap;  // entries per sublist, multiple of 4
};

// 512-key tiles once a 256-key tile would give a window fewer than ~16 entries ("bloom_part_tile": 0 auto, 256, 512);
// they exist for 16-byte keys and k <= 8 only (shared memory of the sorted tile)
inline bool part_big_tile(const pb_ctx *ctx, uint32_t n_windows, uint32_t k, bool fixed16, bool query = false) {
    if (!fixed16 || k > (query ? 7u : 8u) || ctx->bloom_part_tile == 256) return false;
    if (ctx->bloom_part_tile == 512) return true;
    return n_windows > 112;
}

// Layout for launches of up to n_keys keys: k hashes, windows of 2^window_log2 bits of an m-bit filter.
// overlapped: pass 2 of the previous chunk runs beside this launch -- then three pass-1 CTAs per SM instead of four, which
// leaves registers and shared memory for three pass-2 CTAs instead of one (r2 sweep: 63.1 -> 59.6 ms per 1e9 keys).
inline PartLayout part_layout(const pb_ctx *ctx, uint64_t n_keys, uint32_t k, uint64_t m, uint32_t window_log2, uint32_t n_windows,
                              bool fixed16, bool overlapped, bool query = false) {
    PartLayout L;
    L.block = part_big_tile(ctx, n_windows, k, fixed16, query) ? 512 : 256;
    // K <= 8: 56 registers -> four 256-thread CTAs per SM (+ one pass-2 CTA); K > 8: 72 registers -> three
    const int most = k <= 8 ? 4 : 3;
    const int64_t want = ctx->bloom_part_ctas_per_sm > 0 ? ctx->bloom_part_ctas_per_sm : (overlapped ? 3 : most);
    const int base = (int)std::max<int64_t>(1, std::min<int64_t>(want, most));
    const int per_sm = L.block == 512 ? (base + 1) / 2 : base;  // 512-thread CTAs: two per SM
    const uint64_t tiles = (n_keys + L.block - 1) / L.block;
    L.grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)ctx->num_sms * per_sm));
    const double keys_per_cta = (double)((tiles + L.grid - 1) / L.grid) * L.block;
    const double frac = std::min(1.0, (double)(1ull << window_log2) / (double)m);  // share of the indices one window gets
    const double e = keys_per_cta * k * frac;
    const double cap = e + 7.0 * sqrt(e) + 64.0;
    L.sub_cap = ((uint32_t)cap + 3u) & ~3u;
    return L;
}

// fourth generation of pass 1 (pb_bloom_part4.cu): any key layout, K = 1..16 hashes.  Launches pd.n_sub CTAs of
// `block` threads (CTAs beyond the tiles of a short batch only write their zero counts).
cudaError_t launch_part4(int block, cudaStream_t stream, const DevKeys &dk, const PartDev &pd);
// the same pass for a partitioned query (16-byte keys only): also stores each index's key number into pd.ids
cudaError_t launch_part4_ids(int block, cudaStream_t stream, const DevKeys &dk, const PartDev &pd);

// ---- pass 2 ----------------------------------------------------------------------------------------------------
// OR lists into a window's slice of the bitmap.  The lists are streamed through shared memory with TMA bulk copies
// (cp.async.bulk, four 8 KB stages behind mbarriers, L2 evict-first so they do not push the window out of L2): the
// memory-level parallelism lives in the copy queue, not in registers or resident warps.  That matters because pass 2
// shares the SMs with pass 1 of the next chunk and gets ONE CTA per SM there; the register-staged version of round 1
// (one LDG.128 in flight per thread, long-scoreboard stall 180 per issue) ran at a third of its stand-alone rate
// when co-resident and, through the two staging halves, throttled pass 1 as well.
constexpr uint32_t kApplyTile = 2048;  // entries per stage (8 KB)
constexpr int kApplyStages = 4;

struct ApplySmem {
    alignas(128) uint32_t buf[kApplyStages][kApplyTile];
    alignas(8) uint64_t full[kApplyStages];
};

__device__ __forceinline__ void tma_bulk_g2s_evict_first(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

__device__ __forceinline__ void apply_smem_init(ApplySmem &sm) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kApplyStages; ++i) mbar_init(&sm.full[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
}

// `g` counts the tiles this CTA has pushed through the stages so far (stage = g % S, mbarrier phase = g / S & 1);
// it is CTA-uniform and carried from list to list.  All threads of the CTA call this together.
// ADD: the window is a slice of a COUNTER array (Counting Bloom) and every entry adds `amount` to its counter
// instead of setting its bit.
template <bool ADD = false>
__device__ __forceinline__ void apply_list_tma(uint32_t *__restrict__ words, const uint32_t *__restrict__ list, uint32_t cnt, ApplySmem &sm,
                                               uint32_t &g, uint32_t amount = 1u) {
    const uint32_t T = (cnt + kApplyTile - 1) / kApplyTile;
    auto issue = [&](uint32_t i) {  // thread 0 only
        const uint32_t st = (g + i) % kApplyStages;
        const uint32_t left = cnt - i * kApplyTile;
        const uint32_t bytes = (left >= kApplyTile ? kApplyTile : ((left + 3u) & ~3u)) * 4u;  // lists are padded to 4 entries
        fence_proxy_async_smem();
        mbar_expect_tx(&sm.full[st], bytes);
        tma_bulk_g2s_evict_first(sm.buf[st], list + (size_t)i * kApplyTile, bytes, &sm.full[st]);
    };
    if (threadIdx.x == 0)
        for (uint32_t i = 0; i < T && i < (uint32_t)kApplyStages; ++i) issue(i);
    for (uint32_t i = 0; i < T; ++i) {
        const uint32_t st = (g + i) % kApplyStages;
        mbar_wait(&sm.full[st], ((g + i) / kApplyStages) & 1u);
        const uint32_t n = min(kApplyTile, cnt - i * kApplyTile);
        const uint4 *b4 = reinterpret_cast<const uint4 *>(sm.buf[st]);
#pragma unroll
        for (uint32_t q = 0; q < kApplyTile / 4 / 256; ++q) {
            const uint32_t e = (q * 256 + threadIdx.x) * 4;
            if (e < n) {
                const uint4 v = b4[q * 256 + threadIdx.x];
                if (ADD) {
                    atomicAdd(words + v.x, amount);
                    if (e + 1 < n) atomicAdd(words + v.y, amount);
                    if (e + 2 < n) atomicAdd(words + v.z, amount);
                    if (e + 3 < n) atomicAdd(words + v.w, amount);
                } else {
                    atomicOr(words + (v.x >> 5), 1u << (v.x & 31));
                    if (e + 1 < n) atomicOr(words + (v.y >> 5), 1u << (v.y & 31));
                    if (e + 2 < n) atomicOr(words + (v.z >> 5), 1u << (v.z & 31));
                    if (e + 3 < n) atomicOr(words + (v.w >> 5), 1u << (v.w & 31));
                }
            }
        }
        __syncthreads();  // everybody is done reading the stage before it is refilled
        if (threadIdx.x == 0 && i + kApplyStages < T) issue(i + kApplyStages);
    }
    g += T;
}

// pass 2: one window at a time (launch order); its bitmap slice stays L2 resident while its sublists stream by
static __global__ void __launch_bounds__(256) bloom_apply2(PartDev p, uint32_t ctas_per_window) {
    __shared__ ApplySmem sm;
    apply_smem_init(sm);
    const uint32_t w = blockIdx.x / ctas_per_window;
    const uint32_t c = blockIdx.x % ctas_per_window;
    uint32_t *words = p.words + ((uint64_t)w << (p.window_log2 - 5));
    uint32_t g = 0;
    for (uint32_t s = c; s < p.n_sub; s += ctas_per_window) {
        const size_t li = (size_t)w * p.n_sub + s;
        uint32_t cnt = p.counts[li];
        if (cnt > p.sub_cap) cnt = p.sub_cap;
        apply_list_tma(words, p.stage + li * p.sub_cap, cnt, sm, g);
    }
}

// pass 2 for a Counting Bloom filter: windows of 2^window_log2 COUNTERS (p.words = the counter array); an index that
// two hashes of one key share is listed twice and counts twice, as countingbloom.py:143-153 does
static __global__ void __launch_bounds__(256) cbloom_apply2(PartDev p, uint32_t ctas_per_window, uint32_t amount) {
    __shared__ ApplySmem sm;
    apply_smem_init(sm);
    const uint32_t w = blockIdx.x / ctas_per_window;
    const uint32_t c = blockIdx.x % ctas_per_window;
    uint32_t *counters = p.words + ((uint64_t)w << p.window_log2);
    uint32_t g = 0;
    for (uint32_t s = c; s < p.n_sub; s += ctas_per_window) {
        const size_t li = (size_t)w * p.n_sub + s;
        uint32_t cnt = p.counts[li];
        if (cnt > p.sub_cap) cnt = p.sub_cap;
        apply_list_tma<true>(counters, p.stage + li * p.sub_cap, cnt, sm, g, amount);
    }
}

// pass 2 on a range shard (multi-GPU): window w of this shard receives the sublists of every source rank, laid
// out [source][wps][n_sub][sub_cap] with counts [source][wps][n_sub] -- exactly what the exchange delivers.
static __global__ void __launch_bounds__(256) bloom_apply_sources(uint32_t *__restrict__ shard_words, const uint32_t *__restrict__ stage,
                                                           const uint32_t *__restrict__ counts, uint32_t n_sources, uint32_t wps,
                                                           uint32_t n_sub, uint32_t sub_cap, uint32_t window_log2,
                                                           uint32_t ctas_per_window) {
    __shared__ ApplySmem sm;
    apply_smem_init(sm);
    const uint32_t w = blockIdx.x / ctas_per_window;
    const uint32_t c = blockIdx.x % ctas_per_window;
    uint32_t *words = shard_words + ((uint64_t)w << (window_log2 - 5));
    const uint32_t lists = n_sources * n_sub;  // lists of this window: (source, sublist) pairs
    uint32_t g = 0;
    for (uint32_t q = c; q < lists; q += ctas_per_window) {
        const uint32_t src = q / n_sub, s = q % n_sub;
        const size_t li = ((size_t)src * wps + w) * n_sub + s;
        uint32_t cnt = counts[li];
        if (cnt > sub_cap) cnt = sub_cap;
        apply_list_tma(words, stage + li * sub_cap, cnt, sm, g);
    }
}

// ---- pass 2 of the partitioned QUERY: test the staged bit indices while their window is L2 resident; a clear bit
// writes 0 into its key's answer byte (the bytes were preset to 1, every writer stores 0: no atomics needed).
// Entries and key numbers stream through shared memory with TMA bulk copies exactly like bloom_apply2's lists.
struct ProbeSmem {
    alignas(128) uint32_t loc[kApplyStages][kApplyTile];
    alignas(128) uint32_t id[kApplyStages][kApplyTile];
    alignas(8) uint64_t full[kApplyStages];
};

static __global__ void __launch_bounds__(256) bloom_probe2(PartDev p, uint32_t ctas_per_window) {
    extern __shared__ __align__(128) uint8_t probe_smem_raw[];
    ProbeSmem &sm = *reinterpret_cast<ProbeSmem *>(probe_smem_raw);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kApplyStages; ++i) mbar_init(&sm.full[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const uint32_t w = blockIdx.x / ctas_per_window;
    const uint32_t c = blockIdx.x % ctas_per_window;
    const uint32_t *__restrict__ words = p.words + ((uint64_t)w << (p.window_log2 - 5));
    uint32_t g = 0;
    for (uint32_t s = c; s < p.n_sub; s += ctas_per_window) {
        const size_t li = (size_t)w * p.n_sub + s;
        uint32_t cnt = p.counts[li];
        if (cnt > p.sub_cap) cnt = p.sub_cap;
        const uint32_t *list = p.stage + li * p.sub_cap;
        const uint32_t *ids = p.ids + li * p.sub_cap;
        const uint32_t T = (cnt + kApplyTile - 1) / kApplyTile;
        auto issue = [&](uint32_t i) {  // thread 0 only
            const uint32_t st = (g + i) % kApplyStages;
            const uint32_t left = cnt - i * kApplyTile;
            const uint32_t bytes = (left >= kApplyTile ? kApplyTile : ((left + 3u) & ~3u)) * 4u;
            fence_proxy_async_smem();
            mbar_expect_tx(&sm.full[st], 2 * bytes);
            tma_bulk_g2s_evict_first(sm.loc[st], list + (size_t)i * kApplyTile, bytes, &sm.full[st]);
            tma_bulk_g2s_evict_first(sm.id[st], ids + (size_t)i * kApplyTile, bytes, &sm.full[st]);
        };
        if (threadIdx.x == 0)
            for (uint32_t i = 0; i < T && i < (uint32_t)kApplyStages; ++i) issue(i);
        for (uint32_t i = 0; i < T; ++i) {
            const uint32_t st = (g + i) % kApplyStages;
            mbar_wait(&sm.full[st], ((g + i) / kApplyStages) & 1u);
            const uint32_t n = min(kApplyTile, cnt - i * kApplyTile);
            const uint4 *l4 = reinterpret_cast<const uint4 *>(sm.loc[st]);
            const uint4 *i4 = reinterpret_cast<const uint4 *>(sm.id[st]);
            // all eight probes of a thread are in flight before the first one is tested (the loads return through L2:
            // unlike pass 2's REDs they are not fire-and-forget, so memory-level parallelism per thread decides)
            constexpr uint32_t Q = kApplyTile / 4 / 256;
            uint32_t loc[Q][4], wd[Q][4];
#pragma unroll
            for (uint32_t q = 0; q < Q; ++q) {
                const uint32_t e = (q * 256 + threadIdx.x) * 4;
                const uint4 v = l4[q * 256 + threadIdx.x];
                loc[q][0] = v.x, loc[q][1] = v.y, loc[q][2] = v.z, loc[q][3] = v.w;
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) wd[q][j] = e + j < n ? __ldg(words + (loc[q][j] >> 5)) : 0xFFFFFFFFu;
            }
#pragma unroll
            for (uint32_t q = 0; q < Q; ++q) {
                const uint32_t e = (q * 256 + threadIdx.x) * 4;
                bool any = false;
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) any |= e + j < n && !((wd[q][j] >> (loc[q][j] & 31)) & 1u);
                if (any) {  // rare for members: only then are the key numbers read
                    const uint4 k = i4[q * 256 + threadIdx.x];
                    const uint32_t kk[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j)
                        if (e + j < n && !((wd[q][j] >> (loc[q][j] & 31)) & 1u)) p.out[kk[j]] = 0;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0 && i + kApplyStages < T) issue(i + kApplyStages);
        }
        g += T;
    }
}

static inline cudaError_t probe_configure() {
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(bloom_probe2, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bloom_probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ProbeSmem));
    done = e == cudaSuccess;
    return e;
}

// launchers of pass 2 (largest shared-memory carveout, see launch_part4_inst)
static inline cudaError_t apply_prefer_max_smem() {
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(bloom_apply2, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(bloom_apply_sources, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    done = e == cudaSuccess;
    return e;
}

}  // namespace pb
