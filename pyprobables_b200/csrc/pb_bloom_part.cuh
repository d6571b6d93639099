// pb_bloom_part.cuh -- second generation of the partitioned Bloom insert (pass 1 + pass 2).
//
// Same idea as bloom_part_fixed16 / bloom_apply_windows in pb_bloom.cu (bin the bit indices of a chunk of
// keys by L2-sized bitmap window, then apply window by window), rebuilt around what the round-1 ncu profile
// of the first version showed (issue slots 35 % busy; stalls: shared-memory atomics' results and constant
// reloads 27 %, the un-prefetched key load 19 %, three CTA barriers per tile 17 %, and a global cursor
// atomic round trip inside every tile):
//   * window cursors are reserved in quotas: a CTA takes kQuota list entries of a window at a time with ONE
//     global atomic and hands them out from shared memory, so most tiles touch no global atomic at all.
//     Entries of a quota a CTA does not use are filled with kSentinel, which pass 2 skips;
//   * histogram and tile bases ping-pong between two shared-memory copies: two barriers per tile instead of
//     three, and the zeroing rides along in the serial section;
//   * the next tile's key is loaded before the current one is hashed;
//   * list positions are 32-bit entry numbers relative to one base pointer (no 64-bit pointer math per bit);
//   * h % m uses a 32-bit reciprocal when m > 2^32 (every filter large enough to take this path): two
//     IMAD.WIDE instead of four for the high product.
// The result is exact for any input: a position beyond a window's capacity falls back to a direct RED.OR.
#pragma once
#include "pb_common.cuh"
#include "pb_hash.cuh"
#include "pb_keys.cuh"

namespace pb {

constexpr uint32_t kSentinel = 0xFFFFFFFFu;  // never a valid window-local bit index (windows are <= 2^31 bits)
constexpr uint32_t kQuota = 1024;            // most list entries a CTA reserves per window and refill (Part2Dev.quota)
constexpr int kMaxWindows2 = 512;

struct Part2Dev {
    uint32_t *stage;         // n_windows * cap entries
    unsigned int *cursors;   // n_windows, entries reserved so far per window (may run past cap)
    uint32_t *words;         // the bitmap (overflow fallback)
    uint64_t m;              // number of bits (modulus)
    uint64_t recip;          // floor(2^64 / m)
    uint32_t cap;            // entries per window list
    uint32_t window_log2;    // <= 31
    uint32_t n_windows;
    uint32_t k;
    uint32_t recip_fits32;   // m > 2^32
    // m >= 2^33 (every BASELINE-size filter): quotient < 2^31, so h - q*m is one signed mad.wide + one mad.lo
    uint32_t fast33;
    int32_t m_lo_s;          // low word of m read as a signed number
    uint32_t m_hi_adj;       // high word of m, plus one when the low word's sign bit is set
    uint32_t quota;          // list entries a CTA reserves per window and refill (<= kQuota)
    // multi-GPU routing: the bitmap of most windows lives on another GPU, so an index that does not fit its
    // window list cannot fall back to a local RED; it goes to this list of global bit indices instead
    uint64_t *ovf_list;
    unsigned long long *ovf_count;
    uint64_t ovf_cap;
};

__device__ __forceinline__ void part_overflow(const Part2Dev &p, uint64_t idx) {
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

// modulus-derived fields of Part2Dev (host)
inline void part_set_modulus(Part2Dev &pd, uint64_t m) {
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
__device__ __forceinline__ uint64_t mod_fast33(uint64_t h, const Part2Dev &p) {
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

__device__ __forceinline__ uint64_t mod_any(uint64_t h, const Part2Dev &p) {
    if (p.recip_fits32) return mod_big(h, p.m, (uint32_t)p.recip);
    const uint64_t q = __umul64hi(h, p.recip);
    const uint64_t r = h - q * p.m;
    return r >= p.m ? r - p.m : r;
}

// ---- third generation (round 1; kept as the cross-check of bloom_part4 and for the P2P direct-store variant):
// quota cursors as described above, and the tile's indices are first sorted by
// window in shared memory and then copied out by consecutive threads, so a warp's 32 stores fall into one to
// three contiguous runs instead of ~20 scattered 4-byte writes.  The round-1 profiles showed pass 1 pinned
// at ~45 G L2 write requests/s whatever the window count (time grew with the number of windows because the
// requests per store instruction did); coalesced runs cut the requests by ~8x and make the cost independent
// of the window count.
// P2P = true: the window lists live in the receive staging of the window's OWNER GPU (peer memory mapped over
// NVLink): pass 1 stores its entries straight into rank d's buffer, block `src_rank` of it, so the exchange step
// of the multi-GPU insert is the kernel's own coalesced stores -- no separate all-to-all, no local staging.
struct P2PDst {
    uint32_t *stage[16];      // per destination rank: its receive staging [n_src][wps][cap] for this chunk's half
    uint32_t wps;             // windows per rank
    uint32_t src_rank;        // this GPU's block inside every destination's staging
};

// BS = threads (= keys) per tile: 256, or 512 when there are many windows (longer runs per window, the scan and
// the barriers amortised over twice the entries)
// 56 registers at most: four 256-thread CTAs then leave 8192 registers per SM, exactly one pass-2 CTA, which is what
// lets pass 2 of the previous chunk run beside this kernel (at 64 registers the overlap disappears)
template <int KG, int NG, bool P2P, int BS = 256>
__global__ void __maxnreg__(56) bloom_part3_fixed16(const uint4 *__restrict__ keys, uint64_t n, Part2Dev p, P2PDst dst) {
    __shared__ uint32_t *wbase[P2P ? kMaxWindows2 : 1];  // P2P: start of window w's list in its owner's memory
    __shared__ uint32_t hist[2][kMaxWindows2];
    __shared__ uint32_t tbase[kMaxWindows2];       // window-relative list position of the tile's first entry
    __shared__ uint32_t wstart[kMaxWindows2 + 1];  // exclusive prefix sum of hist: start of the window's run in sorted[]
    __shared__ uint32_t cur[kMaxWindows2], lim[kMaxWindows2];
    __shared__ uint32_t sorted_loc[BS * NG * KG];
    __shared__ uint16_t sorted_win[BS * NG * KG];
    const uint32_t tid = threadIdx.x;
    const uint32_t W = p.n_windows;
    const uint32_t mask = (1u << p.window_log2) - 1u;
    for (uint32_t w = tid; w < W; w += blockDim.x) {
        hist[0][w] = 0;
        hist[1][w] = 0;
        cur[w] = 0;
        lim[w] = 0;
        if (P2P) wbase[w] = dst.stage[w / dst.wps] + (size_t)(dst.src_rank * dst.wps + w % dst.wps) * p.cap;
    }
    __syncthreads();
    auto list_of = [&](uint32_t w) -> uint32_t * { return P2P ? wbase[w] : p.stage + (size_t)w * p.cap; };
    const uint64_t tiles = (n + blockDim.x - 1) / blockDim.x;
    uint32_t pp = 0;
    uint64_t tile = blockIdx.x;
    uint4 nextk = make_uint4(0, 0, 0, 0);
    if (tile < tiles && tile * blockDim.x + tid < n) nextk = __ldcs(keys + tile * blockDim.x + tid);
    const uint32_t per_lane = (W + 31) / 32;
    for (; tile < tiles; tile += gridDim.x) {
        const uint64_t i = tile * blockDim.x + tid;
        const bool live = i < n;
        const uint4 kw = nextk;
        {
            const uint64_t ni = (tile + gridDim.x) * blockDim.x + tid;
            if (ni < n) nextk = __ldcs(keys + ni);
        }
        uint32_t loc[NG * KG];
        uint32_t wr[NG * KG];  // window << 16 | rank within the tile
        if (live) {
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                uint64_t h[KG];
                fnv_group_16<KG>(kw, g * KG, h);
#pragma unroll
                for (int j = 0; j < KG; ++j) {
                    if ((uint32_t)(g * KG + j) < p.k) {
                        const uint64_t idx = mod_any(h[j], p);
                        const uint32_t w = (uint32_t)(idx >> p.window_log2);
                        loc[g * KG + j] = (uint32_t)idx & mask;
                        wr[g * KG + j] = (w << 16) | atomicAdd(&hist[pp][w], 1u);
                    }
                }
            }
        }
        __syncthreads();
        if (tid < 32) {
            // warp 0: exclusive scan of the histogram (lane l owns windows [l*per_lane, (l+1)*per_lane))
            uint32_t sum = 0;
            const uint32_t w0 = tid * per_lane;
            for (uint32_t q = 0; q < per_lane; ++q)
                if (w0 + q < W) sum += hist[pp][w0 + q];
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)tid >= o) incl += v;
            }
            uint32_t run = incl - sum;
            for (uint32_t q = 0; q < per_lane; ++q) {
                if (w0 + q < W) {
                    wstart[w0 + q] = run;
                    run += hist[pp][w0 + q];
                }
            }
            if (tid == 31) wstart[W] = incl;
        } else {
            for (uint32_t w = tid - 32; w < W; w += blockDim.x - 32) {
                const uint32_t need = hist[pp][w];
                uint32_t c = cur[w];
                if (need) {
                    const uint32_t e = lim[w];
                    if (c + need > e) {
                        const uint32_t stop = e < p.cap ? e : p.cap;
                        for (uint32_t q = c; q < stop; ++q) list_of(w)[q] = kSentinel;
                        const uint32_t take = need > p.quota ? need : p.quota;
                        c = atomicAdd(p.cursors + w, take);
                        lim[w] = c + take;
                    }
                    cur[w] = c + need;
                }
                tbase[w] = c;
            }
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int s = 0; s < NG * KG; ++s) {
                if ((uint32_t)s < p.k) {
                    const uint32_t w = wr[s] >> 16;
                    const uint32_t e = wstart[w] + (wr[s] & 0xFFFFu);
                    sorted_loc[e] = loc[s];
                    sorted_win[e] = (uint16_t)w;
                }
            }
        }
        // the other histogram copy was last read in the serial section of the previous tile
        for (uint32_t w = tid; w < W; w += blockDim.x) hist[pp ^ 1][w] = 0;
        __syncthreads();
        const uint32_t total = wstart[W];
        for (uint32_t e = tid; e < total; e += blockDim.x) {
            const uint32_t w = sorted_win[e];
            const uint32_t pos = tbase[w] + (e - wstart[w]);
            const uint32_t v = sorted_loc[e];
            if (pos < p.cap) {
                __stcs(list_of(w) + pos, v);
            } else {
                part_overflow(p, ((uint64_t)w << p.window_log2) | v);
            }
        }
        pp ^= 1;
    }
    __syncthreads();
    for (uint32_t w = tid; w < W; w += blockDim.x) {
        const uint32_t e = lim[w] < p.cap ? lim[w] : p.cap;
        for (uint32_t q = cur[w]; q < e; ++q) list_of(w)[q] = kSentinel;
    }
    if (P2P) __threadfence_system();  // the entries must have reached the owners before the flags are raised
}

// host-side launcher: 512-key tiles only exist for k <= 8 (NG == 1): two index groups would not fit the 48 KB of
// static shared memory
template <int KG, int NG, bool P2P>
static void launch_part3(bool big_tile, int grid, cudaStream_t stream, const uint4 *keys, uint64_t n, const Part2Dev &pd,
                         const P2PDst &dst) {
    if constexpr (NG == 1) {
        if (big_tile && grid >= 2) {
            bloom_part3_fixed16<KG, NG, P2P, 512><<<grid / 2, 512, 0, stream>>>(keys, n, pd, dst);
            return;
        }
    }
    bloom_part3_fixed16<KG, NG, P2P, 256><<<grid, 256, 0, stream>>>(keys, n, pd, dst);
}

// fourth generation of pass 1 (pb_bloom_part4.cu): any key layout, K = 1..16 hashes
constexpr uint32_t kMaxPartK = 16;  // partitioned insert is instantiated for 1..16 hashes (more take the direct path)
cudaError_t launch_part4(uint32_t k, bool big_tile, int grid, cudaStream_t stream, const DevKeys &dk, const Part2Dev &pd);

// pass 2: one window at a time (launch order); its bitmap slice stays L2 resident while its list streams by
static __global__ void __launch_bounds__(256) bloom_apply2(Part2Dev p, uint32_t ctas_per_window) {
    const uint32_t w = blockIdx.x / ctas_per_window;
    const uint32_t c = blockIdx.x % ctas_per_window;
    uint32_t cnt = p.cursors[w];
    if (cnt > p.cap) cnt = p.cap;
    uint32_t *words = p.words + ((uint64_t)w << (p.window_log2 - 5));
    const uint32_t *list = p.stage + (uint64_t)w * p.cap;
    const uint32_t n4 = cnt >> 2;
    const uint4 *list4 = reinterpret_cast<const uint4 *>(list);
    for (uint32_t i = c * blockDim.x + threadIdx.x; i < n4; i += ctas_per_window * blockDim.x) {
        const uint4 v = __ldcs(list4 + i);
        if (v.x != kSentinel) atomicOr(words + (v.x >> 5), 1u << (v.x & 31));
        if (v.y != kSentinel) atomicOr(words + (v.y >> 5), 1u << (v.y & 31));
        if (v.z != kSentinel) atomicOr(words + (v.z >> 5), 1u << (v.z & 31));
        if (v.w != kSentinel) atomicOr(words + (v.w >> 5), 1u << (v.w & 31));
    }
    if (c == 0) {
        for (uint32_t i = (n4 << 2) + threadIdx.x; i < cnt; i += blockDim.x) {
            const uint32_t v = list[i];
            if (v != kSentinel) atomicOr(words + (v >> 5), 1u << (v & 31));
        }
    }
}

// pass 2 on a range shard (multi-GPU): window w of this shard receives one list per source rank,
// laid out [source][w][cap] with counts [source][w] -- exactly what the all-to-all of the per-rank stagings
// delivers.
static __global__ void __launch_bounds__(256) bloom_apply_sources(uint32_t *__restrict__ shard_words, const uint32_t *__restrict__ stage,
                                                           const unsigned int *__restrict__ cursors, uint32_t n_sources,
                                                           uint32_t wps, uint32_t cap, uint32_t window_log2,
                                                           uint32_t ctas_per_window) {
    const uint32_t w = blockIdx.x / ctas_per_window;
    const uint32_t c = blockIdx.x % ctas_per_window;
    uint32_t *words = shard_words + ((uint64_t)w << (window_log2 - 5));
    for (uint32_t s = 0; s < n_sources; ++s) {
        uint32_t cnt = cursors[s * wps + w];
        if (cnt > cap) cnt = cap;
        const uint32_t *list = stage + (uint64_t)(s * wps + w) * cap;
        const uint32_t n4 = cnt >> 2;
        const uint4 *list4 = reinterpret_cast<const uint4 *>(list);
        for (uint32_t i = c * blockDim.x + threadIdx.x; i < n4; i += ctas_per_window * blockDim.x) {
            const uint4 v = __ldcs(list4 + i);
            if (v.x != kSentinel) atomicOr(words + (v.x >> 5), 1u << (v.x & 31));
            if (v.y != kSentinel) atomicOr(words + (v.y >> 5), 1u << (v.y & 31));
            if (v.z != kSentinel) atomicOr(words + (v.z >> 5), 1u << (v.z & 31));
            if (v.w != kSentinel) atomicOr(words + (v.w >> 5), 1u << (v.w & 31));
        }
        if (c == 0) {
            for (uint32_t i = (n4 << 2) + threadIdx.x; i < cnt; i += blockDim.x) {
                const uint32_t v = list[i];
                if (v != kSentinel) atomicOr(words + (v >> 5), 1u << (v & 31));
            }
        }
    }
}

}  // namespace pb
