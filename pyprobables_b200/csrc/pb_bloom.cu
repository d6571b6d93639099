// pb_bloom.cu -- BloomFilter.add / check for whole batches (reference: probables/blooms/bloom.py:234-272).
//
// State: the reference's byte array (bit b -> byte b/8, mask 1<<(b%8), bloom.py:247-249) held as
// little-endian u32 words, so bit b is bit (b & 31) of word (b >> 5) and RED.OR.b32 reproduces the
// byte array exactly.
//
// Insert has two device paths:
//   direct      one thread per key: k FNV chains -> exact % num_bits -> k RED.OR.b32 into HBM.
//               Every RED is a random 32-byte sector read-modify-write (64 B of DRAM traffic).
//   partitioned pass 1 hashes the keys once and bins the bit indices by bitmap *window*
//               (2^W bits, sized to sit in B200's 126 MB L2) into a staging area with coalesced,
//               write-combined stores; pass 2 walks the windows in order and applies each
//               window's indices with RED.OR while that window is L2-resident.  DRAM traffic drops
//               from 16+64k to about 16+8k bytes per key plus one read+write of the bitmap per chunk.
// Bucket overflow (skewed / duplicate keys) falls back to direct REDs, so the result is exact for
// any input.
#include <algorithm>
#include <vector>
#include <new>

#include "pb_common.cuh"
#include "pb_hash.cuh"
#include "pb_keys.cuh"
#include "pb_bloom_part.cuh"

using namespace pb;

struct pb_bloom {
    pb_ctx *ctx = nullptr;
    uint64_t num_bits = 0;  // modulus of the (global) filter
    uint32_t k = 0;
    uint64_t lo_bit = 0, hi_bit = 0;  // bits owned by this handle ([0, num_bits) unless sharded)
    uint32_t *words = nullptr;
    uint64_t nbytes = 0;  // ceil((hi-lo)/8)
    uint64_t nwords = 0;  // allocation, multiple of 4 words
    FastMod fm;
    uint32_t *first_setter = nullptr;  // check-then-add batches: per owned bit, (0xFF - epoch) << 24 | lowest row of the call that touched it
    uint32_t first_epoch = 0;          // 1..254; a later call's tags are smaller than any earlier call's, so stale entries lose the min
};

namespace pb {

uint32_t *bloom_words(pb_bloom *b) { return b->words; }
pb_ctx *bloom_ctx(pb_bloom *b) { return b->ctx; }

struct BloomDev {
    uint32_t *words;
    FastMod fm;
    uint64_t lo, hi;  // owned bit range
    uint32_t k;
};

__device__ __forceinline__ void bloom_set(const BloomDev &b, uint64_t h) {
    const uint64_t idx = fastmod(h, b.fm);
    if (idx >= b.lo && idx < b.hi) {
        const uint64_t l = idx - b.lo;
        atomicOr(b.words + (l >> 5), 1u << (uint32_t)(l & 31));  // result unused -> RED.E.OR
    }
}

__device__ __forceinline__ uint32_t bloom_test(const BloomDev &b, uint64_t h) {
    const uint64_t idx = fastmod(h, b.fm);
    if (idx >= b.lo && idx < b.hi) {
        const uint64_t l = idx - b.lo;
        return (__ldg(b.words + (l >> 5)) >> (uint32_t)(l & 31)) & 1u;
    }
    return 1u;  // not this shard's bit: neutral for the AND
}

// ---------------------------------------------------------------- direct insert / check
template <int KG>
__global__ void __launch_bounds__(256) bloom_add_fixed16(const uint4 *__restrict__ keys, uint64_t n, BloomDev b) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 w = __ldcs(keys + i);
        for (uint32_t s0 = 0; s0 < b.k; s0 += KG) {
            uint64_t h[KG];
            fnv_group_16<KG>(w, s0, h);
#pragma unroll
            for (int j = 0; j < KG; ++j)
                if (s0 + j < b.k) bloom_set(b, h[j]);
        }
    }
}

template <int KG>
__global__ void __launch_bounds__(256)
    bloom_check_fixed16(const uint4 *__restrict__ keys, uint64_t n, BloomDev b, uint8_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 w = __ldcs(keys + i);
        uint32_t ok = 1u;
        for (uint32_t s0 = 0; s0 < b.k; s0 += KG) {
            uint64_t h[KG];
            fnv_group_16<KG>(w, s0, h);
            uint32_t bit[KG];
            // Probes of a phase are issued together (independent sector reads in flight).  The reference stops at
            // the first zero bit (bloom.py:268-271); two phases keep most of that saving for absent keys -- at 50 %
            // fill 7 of 8 absent keys end after the first three probes -- without serialising present keys.
            constexpr int kFirst = KG < 3 ? KG : 3;
#pragma unroll
            for (int j = 0; j < kFirst; ++j) bit[j] = (s0 + j < b.k) ? bloom_test(b, h[j]) : 1u;
#pragma unroll
            for (int j = 0; j < kFirst; ++j) ok &= bit[j];
            if (ok) {
#pragma unroll
                for (int j = kFirst; j < KG; ++j) bit[j] = (s0 + j < b.k) ? bloom_test(b, h[j]) : 1u;
#pragma unroll
                for (int j = kFirst; j < KG; ++j) ok &= bit[j];
            }
            if (!ok) break;
        }
        out[i] = (uint8_t)ok;
    }
}

template <int KG, int SYMW, bool CHECK>
__global__ void __launch_bounds__(kTileKeys) bloom_staged(DevKeys dk, BloomDev b, uint8_t *__restrict__ out) {
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
            uint32_t ok = 1u;
            for (uint32_t s0 = 0; s0 < b.k; s0 += KG) {
                uint64_t h[KG];
                fnv_group_ptr<KG, SYMW>(kr.p, kr.len, s0, h);
#pragma unroll
                for (int j = 0; j < KG; ++j) {
                    if (s0 + j < b.k) {
                        if (CHECK) ok &= bloom_test(b, h[j]);
                        else bloom_set(b, h[j]);
                    }
                }
            }
            if (CHECK) out[first + threadIdx.x] = (uint8_t)ok;
        }
    }
}

// add_alt / check_alt: hashes made by a host-side hash_function (bloom.py:241-250, :261-272)
__global__ void __launch_bounds__(256) bloom_add_hashes_kernel(const uint64_t *__restrict__ h, uint64_t total, BloomDev b) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
        bloom_set(b, h[i]);
}
__global__ void __launch_bounds__(256)
    bloom_check_hashes_kernel(const uint64_t *__restrict__ h, uint64_t n, BloomDev b, uint8_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t ok = 1u;
        for (uint32_t s = 0; s < b.k; ++s) ok &= bloom_test(b, h[i * b.k + s]);
        out[i] = (uint8_t)ok;
    }
}

// bit indices that already went through % num_bits (multi-GPU apply side)
__global__ void __launch_bounds__(256)
    bloom_add_idx_kernel(const uint64_t *__restrict__ idx, uint64_t n, BloomDev b, unsigned long long *stray) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g = __ldcs(idx + i);
        if (g >= b.lo && g < b.hi) {
            const uint64_t l = g - b.lo;
            atomicOr(b.words + (l >> 5), 1u << (uint32_t)(l & 31));
        } else {
            atomicAdd(stray, 1ull);
        }
    }
}
__global__ void __launch_bounds__(256)
    bloom_test_idx_kernel(const uint64_t *__restrict__ idx, uint64_t n, BloomDev b, uint8_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g = __ldcs(idx + i);
        uint32_t bit = 0;
        if (g >= b.lo && g < b.hi) {
            const uint64_t l = g - b.lo;
            bit = (__ldg(b.words + (l >> 5)) >> (uint32_t)(l & 31)) & 1u;
        }
        out[i] = (uint8_t)bit;
    }
}

// ---- ordered "check, then add" batches (ExpandingBloomFilter.add_alt, expandingbloom.py:159-169) ------------------
// One key at a time the reference adds key i to the newest filter iff key i is not found when its turn comes.  Keys
// that are found add nothing new (all their bits are set already), so the bits set before key i's turn are the
// filter's bits plus the bits of ALL earlier rows that are not skipped, and key i is added iff it owns a bit that is
// clear in the filter and that no earlier row touches: first_setter[bit] = min row over the batch, then one compare.
// The table is never reset between calls: entries carry the call's epoch in their top byte, counted DOWN, so whatever
// an earlier call left behind is larger than anything this call writes and loses the atomicMin.
__global__ void __launch_bounds__(256)
    bloom_first_setter_kernel(const uint64_t *__restrict__ idx, uint64_t n_entries, uint32_t k, BloomDev b,
                              const uint8_t *__restrict__ skip, uint32_t *__restrict__ first, uint32_t tag) {
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < n_entries; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t row = e / k;
        if (skip && skip[row]) continue;
        const uint64_t l = __ldcs(idx + e) - b.lo;
        if (!((__ldg(b.words + (l >> 5)) >> (uint32_t)(l & 31)) & 1u)) atomicMin(first + l, tag | (uint32_t)row);
    }
}
__global__ void __launch_bounds__(256)
    bloom_novel_rows_kernel(const uint64_t *__restrict__ idx, uint64_t n, uint32_t k, BloomDev b,
                            const uint8_t *__restrict__ skip, const uint32_t *__restrict__ first, uint32_t tag,
                            uint8_t *__restrict__ novel) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t nv = 0;
        if (!(skip && skip[i])) {
            for (uint32_t s = 0; s < k; ++s) {
                const uint64_t l = idx[i * k + s] - b.lo;
                const bool clear = !((__ldg(b.words + (l >> 5)) >> (uint32_t)(l & 31)) & 1u);
                nv |= (clear && __ldcg(first + l) == (tag | (uint32_t)i)) ? 1u : 0u;
            }
        }
        novel[i] = (uint8_t)nv;
    }
}
// rows [0, n) of k bit indices each; a row is applied iff mask[row] != 0 (mask NULL: every row)
__global__ void __launch_bounds__(256)
    bloom_add_rows_kernel(const uint64_t *__restrict__ idx, uint64_t n_entries, uint32_t k, BloomDev b,
                          const uint8_t *__restrict__ mask) {
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < n_entries; e += (uint64_t)gridDim.x * blockDim.x) {
        if (mask && !mask[e / k]) continue;
        const uint64_t l = __ldcs(idx + e) - b.lo;
        atomicOr(b.words + (l >> 5), 1u << (uint32_t)(l & 31));
    }
}
__global__ void __launch_bounds__(256)
    bloom_rows_in_range_kernel(const uint64_t *__restrict__ idx, uint64_t n_entries, BloomDev b, unsigned long long *stray) {
    unsigned long long bad = 0;
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < n_entries; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g = idx[e];
        bad += (g < b.lo || g >= b.hi) ? 1ull : 0ull;
    }
    if (bad) atomicAdd(stray, bad);
}

// found[i] = 1 iff one of the filters holds every bit of row i (ExpandingBloomFilter.check_alt, expandingbloom.py:140-147:
// `any(blm.check_alt(hashes) for blm in self._blooms)`).  One thread per row, first clear bit ends a filter, first
// full match ends the row: ~2 probes per filter for an absent key instead of k.
__global__ void __launch_bounds__(256)
    bloom_rows_in_any_kernel(const uint64_t *__restrict__ idx, uint64_t n, uint32_t k, const uint32_t *const *__restrict__ words,
                             uint32_t n_filters, uint8_t *__restrict__ found) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t *row = idx + i * k;
        uint32_t hit = 0;
        for (uint32_t f = 0; f < n_filters && !hit; ++f) {
            const uint32_t *w = words[f];
            uint32_t s = 0;
            for (; s < k; ++s) {
                const uint64_t l = row[s];
                if (!((__ldg(w + (l >> 5)) >> (uint32_t)(l & 31)) & 1u)) break;
            }
            hit = s == k;
        }
        found[i] = (uint8_t)hit;
    }
}

// bloom.py:371-428: res = a | b (op 0) or a & b (op 1), streaming 128-bit words
__global__ void __launch_bounds__(256) bloom_combine_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ a,
                                                            const uint4 *__restrict__ b, uint64_t n4, int op) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 x = __ldcs(a + i), y = __ldcs(b + i);
        uint4 r;
        if (op == 0) r = make_uint4(x.x | y.x, x.y | y.y, x.z | y.z, x.w | y.w);
        else r = make_uint4(x.x & y.x, x.y & y.y, x.z & y.z, x.w & y.w);
        __stcs(dst + i, r);
    }
}

// bloom.py:430-460: popcounts of a | b and a & b in one pass
__global__ void __launch_bounds__(256) bloom_pair_popcount_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b,
                                                                  uint64_t n4, unsigned long long *out) {
    unsigned long long cu = 0, ci = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 x = __ldcs(a + i), y = __ldcs(b + i);
        cu += __popc(x.x | y.x) + __popc(x.y | y.y) + __popc(x.z | y.z) + __popc(x.w | y.w);
        ci += __popc(x.x & y.x) + __popc(x.y & y.y) + __popc(x.z & y.z) + __popc(x.w & y.w);
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

// bloom.py:552-557
__global__ void __launch_bounds__(256) popcount_kernel(const uint4 *__restrict__ w, uint64_t n4, unsigned long long *out) {
    unsigned long long c = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcs(w + i);
        c += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// bit indices in key order: out[i*k + s] = fnv_1a(key_i, seed s) % num_bits (multi-GPU query: the indices travel to
// the owners of their bits, the answers come back and are ANDed per key at the source)
template <int KG>
__global__ void __launch_bounds__(256) bloom_index_fixed16(const uint4 *__restrict__ keys, uint64_t n, FastMod fm, uint32_t k,
                                                           uint64_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 w = __ldcs(keys + i);
        for (uint32_t s0 = 0; s0 < k; s0 += KG) {
            uint64_t h[KG];
            fnv_group_16<KG>(w, s0, h);
#pragma unroll
            for (int j = 0; j < KG; ++j)
                if (s0 + j < k) out[i * k + s0 + j] = fastmod(h[j], fm);
        }
    }
}
__global__ void __launch_bounds__(256) reduce_mod_kernel(uint64_t *__restrict__ v, uint64_t total, FastMod fm) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
        v[i] = fastmod(v[i], fm);
}
// out[i] = AND of bits[i*k .. i*k+k) (the per-index answers of one key)
__global__ void __launch_bounds__(256) and_rows_kernel(const uint8_t *__restrict__ bits, uint64_t n, uint32_t k, uint8_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t ok = 1u;
        for (uint32_t s = 0; s < k; ++s) ok &= bits[i * k + s];
        out[i] = (uint8_t)ok;
    }
}

// number of 1 bytes in a 0/1 byte vector (hit rate of a check sample)
__global__ void __launch_bounds__(256) count_ones_kernel(const uint4 *__restrict__ v, uint64_t n16, unsigned long long *out) {
    unsigned long long c = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 x = v[i];
        c += __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w);
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// ---------------------------------------------------------------- multi-GPU routing
// hash keys, find the owning shard of every bit index, append the (global) index to that shard's slot.
template <int KG>
__global__ void __launch_bounds__(256) bloom_route_fixed16(const uint4 *__restrict__ keys, uint64_t n, FastMod fm, uint32_t k,
                                                           FastMod shard_div, uint64_t shard_bits, uint32_t n_shards,
                                                           uint64_t *__restrict__ out, uint64_t slot_cap,
                                                           unsigned long long *counts) {
    __shared__ uint32_t hist[64];
    __shared__ unsigned long long base[64];
    (void)shard_bits;
    const uint64_t tiles = (n + blockDim.x - 1) / blockDim.x;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        if (threadIdx.x < n_shards) hist[threadIdx.x] = 0;
        __syncthreads();
        const uint64_t i = tile * blockDim.x + threadIdx.x;
        uint4 w = make_uint4(0, 0, 0, 0);
        if (i < n) w = __ldcs(keys + i);
        for (uint32_t s0 = 0; s0 < k; s0 += KG) {
            uint64_t h[KG];
            uint64_t idx[KG];
            uint32_t dst[KG], rank[KG];
            fnv_group_16<KG>(w, s0, h);
#pragma unroll
            for (int j = 0; j < KG; ++j) {
                if (i < n && s0 + j < k) {
                    idx[j] = fastmod(h[j], fm);
                    dst[j] = (uint32_t)fastdiv(idx[j], shard_div);
                    rank[j] = atomicAdd(&hist[dst[j]], 1u);
                }
            }
            __syncthreads();
            if (threadIdx.x < n_shards) {
                base[threadIdx.x] =
                    hist[threadIdx.x] ? atomicAdd(counts + threadIdx.x, (unsigned long long)hist[threadIdx.x]) : 0ull;
                hist[threadIdx.x] = 0;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < KG; ++j) {
                if (i < n && s0 + j < k) {
                    const unsigned long long pos = base[dst[j]] + rank[j];
                    if (pos < slot_cap) out[(uint64_t)dst[j] * slot_cap + pos] = idx[j];
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------- host side
static BloomDev dev_view(const pb_bloom *b) {
    BloomDev d;
    d.words = b->words;
    d.fm = b->fm;
    d.lo = b->lo_bit;
    d.hi = b->hi_bit;
    d.k = b->k;
    return d;
}

template <int KG>
static int launch_add_direct(pb_ctx *ctx, const DevKeys &dk, const BloomDev &bd) {
    launch_begin(ctx);
    if (is_fixed16(dk)) {
        bloom_add_fixed16<KG><<<grid_for(ctx, dk.n, 256, 8), 256, 0, ctx->stream>>>((const uint4 *)dk.data, dk.n, bd);
    } else {
        uint64_t tiles = (dk.n + kTileKeys - 1) / kTileKeys;
        int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->num_sms * 4);
        if (dk.sym_width == 4) bloom_staged<KG, 4, false><<<grid, kTileKeys, 0, ctx->stream>>>(dk, bd, nullptr);
        else bloom_staged<KG, 1, false><<<grid, kTileKeys, 0, ctx->stream>>>(dk, bd, nullptr);
    }
    return check_launch(ctx, "bloom_add");
}

template <int KG>
static int launch_check(pb_ctx *ctx, const DevKeys &dk, const BloomDev &bd, uint8_t *out) {
    launch_begin(ctx);
    if (is_fixed16(dk)) {
        bloom_check_fixed16<KG><<<grid_for(ctx, dk.n, 256, 8), 256, 0, ctx->stream>>>((const uint4 *)dk.data, dk.n, bd, out);
    } else {
        uint64_t tiles = (dk.n + kTileKeys - 1) / kTileKeys;
        int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->num_sms * 4);
        if (dk.sym_width == 4) bloom_staged<KG, 4, true><<<grid, kTileKeys, 0, ctx->stream>>>(dk, bd, out);
        else bloom_staged<KG, 1, true><<<grid, kTileKeys, 0, ctx->stream>>>(dk, bd, out);
    }
    return check_launch(ctx, "bloom_check");
}

#define PB_DISPATCH_KG(kg, CALL)            \
    switch (kg) {                           \
        case 1: st = CALL(1); break;        \
        case 2: st = CALL(2); break;        \
        case 3: st = CALL(3); break;        \
        case 4: st = CALL(4); break;        \
        case 5: st = CALL(5); break;        \
        case 6: st = CALL(6); break;        \
        case 7: st = CALL(7); break;        \
        default: st = CALL(8); break;       \
    }

struct PartPlan {
    bool use = false;
    uint32_t window_log2 = 0;
    uint32_t n_windows = 0;
    uint64_t chunk_keys = 0;
    int halves = 1;  // 2: the staging is split in two so pass 2 of one chunk overlaps pass 1 of the next
    PartLayout lay{};
};

// Decide whether (and how) a batch of n keys goes through the partitioned path.
// query: the partitioned CHECK (two staged words per index, 16-byte device keys, k <= kMaxQueryK); mode = the
// caller's 0 auto / 1 direct / 2 partitioned option
static PartPlan plan_partition(const pb_bloom *b, uint64_t n, bool fixed16, int64_t mode, bool query = false) {
    PartPlan pl;
    pb_ctx *ctx = b->ctx;
    if (mode == 1 || b->k > (query ? kMaxQueryK : kMaxPartK) || n == 0) return pl;
    if (b->lo_bit != 0 || b->hi_bit != b->num_bits) return pl;  // shards take routed indices instead
    const uint64_t l2 = ctx->l2_bytes ? ctx->l2_bytes : ((uint64_t)96 << 20);
    if (mode == 0) {
        // auto: only when the bitmap cannot live in L2 and the batch is large enough to amortise
        // one extra read+write of the bitmap
        if (b->nbytes <= l2) return pl;
        if ((double)n * b->k * 64.0 < 4.0 * (double)b->nbytes) return pl;
    }
    const uint32_t wl_max = 31;
    uint32_t wl = (uint32_t)std::max<int64_t>(5, std::min<int64_t>(ctx->bloom_window_log2_bits, wl_max));
    while (((b->num_bits + ((1ull << wl) - 1)) >> wl) > (uint64_t)kMaxWindows2 && wl < wl_max) ++wl;
    const uint64_t nw = (b->num_bits + ((1ull << wl) - 1)) >> wl;
    if (nw > (uint64_t)kMaxWindows2) return pl;
    // overlap only pays when the batch is big enough to split; then at least "bloom_min_chunks" chunks so that the
    // pass 2 left over after the last pass 1 is a small part of the whole
    const int halves = (ctx->bloom_overlap && n >= (1ull << 24)) ? 2 : 1;
    uint64_t chunk = n;
    if (halves == 2) chunk = std::max<uint64_t>((n + (uint64_t)ctx->bloom_min_chunks - 1) / (uint64_t)std::max<int64_t>(ctx->bloom_min_chunks, 2), 1ull << 22);
    const uint64_t budget_entries =
        std::min<uint64_t>((uint64_t)ctx->stage_bytes / (query ? 8 : 4) / (uint64_t)halves, 0xFFFFFFF0ull);  // u32 entry numbers
    if (query) chunk = std::min<uint64_t>(chunk, 0xFFFFFF00ull);  // key numbers inside a chunk are u32
    PartLayout lay;
    for (;;) {
        lay = part_layout(ctx, chunk, b->k, b->num_bits, wl, (uint32_t)nw, fixed16, halves == 2, query);
        if ((uint64_t)lay.sub_cap * (uint64_t)lay.grid * nw <= budget_entries) break;
        if (chunk <= 4096) return pl;  // the staging budget cannot even hold a tiny chunk: direct path
        chunk = chunk - chunk / 4;
    }
    pl.use = true;
    pl.halves = halves;
    pl.window_log2 = wl;
    pl.n_windows = (uint32_t)nw;
    pl.chunk_keys = chunk;
    pl.lay = lay;
    return pl;
}

struct AddArgs {
    pb_bloom *b;
    PartPlan plan;
    uint64_t chunk_no = 0;
    bool overlapped = false;  // pass 2 launches went to ctx->aux_stream: join before returning
};

static int add_chunk(pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) {
    (void)first;
    (void)slot;
    AddArgs *a = (AddArgs *)user;
    pb_bloom *b = a->b;
    const BloomDev bd = dev_view(b);
    int st = PB_OK;
    if (a->plan.use) {
        const PartPlan &pl = a->plan;
        const bool overlap = pl.halves == 2;
        const int half = overlap ? (int)(a->chunk_no++ & 1) : 0;
        const size_t n_lists = (size_t)pl.n_windows * (size_t)pl.lay.grid;
        const size_t half_entries = n_lists * pl.lay.sub_cap;
        PB_TRY(scratch_reserve(ctx, ctx->part_stage, half_entries * 4 * (size_t)pl.halves));
        PB_TRY(scratch_reserve(ctx, ctx->part_cursors, n_lists * 4 * (size_t)pl.halves));
        PartDev pd;
        pd.stage = (uint32_t *)ctx->part_stage.p + (size_t)half * half_entries;
        pd.counts = (uint32_t *)ctx->part_cursors.p + (size_t)half * n_lists;
        pd.words = b->words;
        part_set_modulus(pd, b->num_bits);
        pd.sub_cap = pl.lay.sub_cap;
        pd.n_sub = (uint32_t)pl.lay.grid;
        pd.window_log2 = pl.window_log2;
        pd.n_windows = pl.n_windows;
        pd.k = b->k;
        pd.ovf_list = nullptr;
        pd.ovf_count = nullptr;
        pd.ovf_cap = 0;
        // this half of the staging is free again once the pass 2 that read it last has finished
        if (overlap && ctx->apply_pending[half]) PB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_apply[half], 0));
        // the layout was planned for 16-byte keys or for staged keys: a batch is one or the other throughout
        launch_begin(ctx);
        cudaError_t e = launch_part4(pl.lay.block, ctx->stream, dk, pd);
        if (e != cudaSuccess) {
            set_error("launch of bloom_part4 (k=%u) failed: %s", pd.k, cudaGetErrorString(e));
            return PB_ERR_CUDA;
        }
        PB_TRY(check_launch(ctx, "bloom_part"));
        const uint32_t cpw = (uint32_t)ctx->num_sms * (uint32_t)std::max<int64_t>(1, std::min<int64_t>(ctx->bloom_apply_cpw_per_sm, 32));
        cudaStream_t s2 = ctx->stream;
        if (overlap) {
            s2 = ctx->aux_stream;
            PB_CUDA(cudaEventRecord(ctx->ev_part[half], ctx->stream));
            PB_CUDA(cudaStreamWaitEvent(s2, ctx->ev_part[half], 0));
        }
        PB_CUDA(apply_prefer_max_smem());
        launch_begin(ctx, s2);
        bloom_apply2<<<pl.n_windows * cpw, 256, 0, s2>>>(pd, cpw);
        PB_TRY(check_launch(ctx, "bloom_apply_windows", s2));
        if (overlap) {
            PB_CUDA(cudaEventRecord(ctx->ev_apply[half], s2));
            ctx->apply_pending[half] = true;
            a->overlapped = true;
        }
        return PB_OK;
    }
    const int kg = pick_group(b->k);
#define CALL(K) launch_add_direct<K>(ctx, dk, bd)
    PB_DISPATCH_KG(kg, CALL)
#undef CALL
    return st;
}

struct CheckArgs {
    pb_bloom *b;
    uint8_t *out_dev;
    uint8_t *out_host;
};

static int check_chunk(pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) {
    CheckArgs *a = (CheckArgs *)user;
    const BloomDev bd = dev_view(a->b);
    uint8_t *out = a->out_dev ? a->out_dev + first : nullptr;
    if (!out) {
        PB_TRY(scratch_reserve(ctx, ctx->out_stage[slot], dk.n));
        out = (uint8_t *)ctx->out_stage[slot].p;
    }
    int st = PB_OK;
    const int kg = pick_group(a->b->k);
#define CALL(K) launch_check<K>(ctx, dk, bd, out)
    PB_DISPATCH_KG(kg, CALL)
#undef CALL
    PB_TRY(st);
    if (a->out_host) PB_CUDA(cudaMemcpyAsync(a->out_host + first, out, dk.n, cudaMemcpyDeviceToHost, ctx->stream));
    return PB_OK;
}

static int bloom_alloc(pb_ctx *ctx, uint64_t num_bits, uint32_t k, uint64_t lo, uint64_t hi, pb_bloom **out) {
    PB_REQUIRE(ctx && out, "NULL argument");
    PB_REQUIRE(num_bits >= 1, "num_bits must be >= 1");
    PB_REQUIRE(k >= 1, "number of hashes must be >= 1");
    PB_REQUIRE(lo < hi && hi <= num_bits, "bad shard range");
    PB_REQUIRE((lo & 31u) == 0, "shard start must be a multiple of 32 bits");
    DeviceGuard g(ctx->device);
    pb_bloom *b = new (std::nothrow) pb_bloom();
    if (!b) return PB_ERR_OOM;
    b->ctx = ctx;
    b->num_bits = num_bits;
    b->k = k;
    b->lo_bit = lo;
    b->hi_bit = hi;
    b->nbytes = (hi - lo + 7) / 8;
    b->nwords = (((b->nbytes + 3) / 4) + 3) & ~(uint64_t)3;
    b->fm = make_fastmod(num_bits);
    cudaError_t e = cudaMalloc(&b->words, b->nwords * 4);
    if (e != cudaSuccess) {
        set_error("cudaMalloc of %llu bitmap bytes failed: %s", (unsigned long long)(b->nwords * 4), cudaGetErrorString(e));
        delete b;
        return PB_ERR_OOM;
    }
    e = cudaMemsetAsync(b->words, 0, b->nwords * 4, ctx->stream);
    if (e != cudaSuccess) {
        set_error("memset failed: %s", cudaGetErrorString(e));
        cudaFree(b->words);
        delete b;
        return PB_ERR_CUDA;
    }
    *out = b;
    return PB_OK;
}

}  // namespace pb

extern "C" {

int pb_bloom_create(pb_ctx *ctx, uint64_t num_bits, uint32_t k, pb_bloom **out) {
    return bloom_alloc(ctx, num_bits, k, 0, num_bits, out);
}

int pb_bloom_create_shard(pb_ctx *ctx, uint64_t num_bits, uint32_t k, uint64_t lo_bit, uint64_t hi_bit, pb_bloom **out) {
    return bloom_alloc(ctx, num_bits, k, lo_bit, hi_bit, out);
}

int pb_bloom_destroy(pb_bloom *b) {
    if (!b) return PB_OK;
    DeviceGuard g(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    cudaFree(b->words);
    cudaFree(b->first_setter);
    delete b;
    return PB_OK;
}

int pb_bloom_clear(pb_bloom *b) {
    PB_REQUIRE(b, "handle is NULL");
    DeviceGuard g(b->ctx->device);
    PB_CUDA(cudaMemsetAsync(b->words, 0, b->nwords * 4, b->ctx->stream));
    return PB_OK;
}

int pb_bloom_upload(pb_bloom *b, const uint8_t *bytes, uint64_t nbytes) {
    PB_REQUIRE(b && bytes, "NULL argument");
    PB_REQUIRE(nbytes == b->nbytes, "expected %llu bytes, got %llu", (unsigned long long)b->nbytes, (unsigned long long)nbytes);
    DeviceGuard g(b->ctx->device);
    PB_CUDA(cudaMemsetAsync(b->words, 0, b->nwords * 4, b->ctx->stream));
    PB_CUDA(cudaMemcpyAsync(b->words, bytes, nbytes, cudaMemcpyHostToDevice, b->ctx->stream));
    PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    return PB_OK;
}

int pb_bloom_download(pb_bloom *b, uint8_t *bytes, uint64_t nbytes) {
    PB_REQUIRE(b && bytes, "NULL argument");
    PB_REQUIRE(nbytes == b->nbytes, "expected %llu bytes, got %llu", (unsigned long long)b->nbytes, (unsigned long long)nbytes);
    DeviceGuard g(b->ctx->device);
    PB_CUDA(cudaMemcpyAsync(bytes, b->words, nbytes, cudaMemcpyDeviceToHost, b->ctx->stream));
    PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    return PB_OK;
}

int pb_bloom_device_ptr(pb_bloom *b, void **out_dev, uint64_t *out_nbytes) {
    PB_REQUIRE(b && out_dev, "NULL argument");
    *out_dev = b->words;
    if (out_nbytes) *out_nbytes = b->nbytes;
    return PB_OK;
}

int pb_bloom_add_keys(pb_bloom *b, const pb_keys *keys) {
    PB_REQUIRE(b && keys, "NULL argument");
    DeviceGuard g(b->ctx->device);
    PB_TRY(validate_keys(keys));
    AddArgs a;
    a.b = b;
    const bool fixed16 = keys->offsets == nullptr && keys->sym_width == 1 && keys->stride == 16 &&
                         (keys->on_device ? ((uintptr_t)keys->data & 15u) == 0 : true);
    a.plan = plan_partition(b, keys->n, fixed16, b->ctx->bloom_insert_mode);
    pb_ctx *ctx = b->ctx;
    const int st = for_each_chunk(ctx, keys, add_chunk, &a, a.plan.use ? a.plan.chunk_keys : 0);
    if (a.overlapped) {
        // later work on the context's stream (checks, popcount, downloads) must see every pass 2
        for (int h = 0; h < 2; ++h) {
            if (ctx->apply_pending[h]) {
                PB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_apply[h], 0));
                ctx->apply_pending[h] = false;
            }
        }
        if (!keys->on_device) PB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return st;
}

// Partitioned query of device-resident 16-byte keys (out on the device): per chunk, preset the answers to 1, bin
// the bit indices with their key numbers by window (bloom_part4<IDS>), then test every window's indices while the
// window is L2 resident (bloom_probe2) -- the same two overlapped passes as the insert, without a random DRAM access.
static int check_partitioned(pb_bloom *b, const uint4 *keys, uint64_t n, uint8_t *out, const PartPlan &pl) {
    pb_ctx *ctx = b->ctx;
    const bool overlap = pl.halves == 2;
    const size_t n_lists = (size_t)pl.n_windows * (size_t)pl.lay.grid;
    const size_t half_entries = n_lists * pl.lay.sub_cap;
    PB_TRY(scratch_reserve(ctx, ctx->part_stage, half_entries * 8 * (size_t)pl.halves));
    PB_TRY(scratch_reserve(ctx, ctx->part_cursors, n_lists * 4 * (size_t)pl.halves));
    PB_CUDA(probe_configure());
    const uint32_t cpw = (uint32_t)ctx->num_sms * (uint32_t)std::max<int64_t>(1, std::min<int64_t>(ctx->bloom_apply_cpw_per_sm, 32));
    uint64_t chunk_no = 0;
    for (uint64_t c0 = 0; c0 < n; c0 += pl.chunk_keys, ++chunk_no) {
        const uint64_t cn = std::min<uint64_t>(pl.chunk_keys, n - c0);
        const int half = overlap ? (int)(chunk_no & 1) : 0;
        PartDev pd;
        pd.stage = (uint32_t *)ctx->part_stage.p + (size_t)half * 2 * half_entries;
        pd.ids = pd.stage + half_entries;
        pd.counts = (uint32_t *)ctx->part_cursors.p + (size_t)half * n_lists;
        pd.words = b->words;
        pd.out = out + c0;
        part_set_modulus(pd, b->num_bits);
        pd.sub_cap = pl.lay.sub_cap;
        pd.n_sub = (uint32_t)pl.lay.grid;
        pd.window_log2 = pl.window_log2;
        pd.n_windows = pl.n_windows;
        pd.k = b->k;
        pd.ovf_list = nullptr;
        pd.ovf_count = nullptr;
        pd.ovf_cap = 0;
        if (overlap && ctx->apply_pending[half]) PB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_apply[half], 0));
        PB_CUDA(cudaMemsetAsync(pd.out, 1, cn, ctx->stream));
        DevKeys dk;
        dk.data = (const uint8_t *)(keys + c0);
        dk.offsets = nullptr;
        dk.n = cn;
        dk.stride = 16;
        dk.sym_width = 1;
        dk.base_symbol = 0;
        dk.total_bytes = cn * 16;
        launch_begin(ctx);
        cudaError_t e = launch_part4_ids(pl.lay.block, ctx->stream, dk, pd);
        if (e != cudaSuccess) {
            set_error("launch of bloom_part4 (query, k=%u) failed: %s", pd.k, cudaGetErrorString(e));
            return PB_ERR_CUDA;
        }
        PB_TRY(check_launch(ctx, "bloom_check_part"));
        cudaStream_t s2 = ctx->stream;
        if (overlap) {
            s2 = ctx->aux_stream;
            PB_CUDA(cudaEventRecord(ctx->ev_part[half], ctx->stream));
            PB_CUDA(cudaStreamWaitEvent(s2, ctx->ev_part[half], 0));
        }
        launch_begin(ctx, s2);
        bloom_probe2<<<pl.n_windows * cpw, 256, sizeof(ProbeSmem), s2>>>(pd, cpw);
        PB_TRY(check_launch(ctx, "bloom_check_probe", s2));
        if (overlap) {
            PB_CUDA(cudaEventRecord(ctx->ev_apply[half], s2));
            ctx->apply_pending[half] = true;
        }
    }
    for (int h = 0; h < 2; ++h) {
        if (ctx->apply_pending[h]) {
            PB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_apply[h], 0));
            ctx->apply_pending[h] = false;
        }
    }
    return PB_OK;
}

int pb_bloom_check_keys(pb_bloom *b, const pb_keys *keys, uint8_t *out, int out_on_device) {
    PB_REQUIRE(b && keys, "NULL argument");
    PB_REQUIRE(out || keys->n == 0, "out is NULL");
    pb_ctx *ctx = b->ctx;
    DeviceGuard g(ctx->device);
    PB_TRY(validate_keys(keys));
    pb_keys rest = *keys;
    uint8_t *out_rest = out;
    const bool fixed16 = keys->offsets == nullptr && keys->sym_width == 1 && keys->stride == 16 && ((uintptr_t)keys->data & 15u) == 0;
    const int64_t mode = ctx->bloom_check_mode;
    if (mode != 1 && keys->on_device && out_on_device && fixed16 && keys->n >= (mode == 2 ? 1u : (1u << 24))) {
        // auto: the partitioned query pays when most probes run to the last bit (keys that ARE members); absent keys
        // leave the direct kernel after two probes on average.  Decide on the hit rate of the first 2^20 keys.
        bool go = true;
        if (mode == 0) {
            const uint64_t ns = 1u << 20;
            pb_keys head = *keys;
            head.n = ns;
            CheckArgs a{b, out, nullptr};
            PB_TRY(for_each_chunk(ctx, &head, check_chunk, &a));
            PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
            unsigned long long *acc = (unsigned long long *)ctx->small.p + 100;
            PB_CUDA(cudaMemsetAsync(acc, 0, 8, ctx->stream));
            count_ones_kernel<<<grid_for(ctx, ns / 16, 256, 8), 256, 0, ctx->stream>>>((const uint4 *)out, ns / 16, acc);
            PB_TRY(check_launch(ctx, "bloom_check_sample"));
            uint64_t *h = (uint64_t *)((uint8_t *)ctx->pinned_small + 3584);
            PB_CUDA(cudaMemcpyAsync(h, acc, 8, cudaMemcpyDeviceToHost, ctx->stream));
            PB_CUDA(cudaStreamSynchronize(ctx->stream));
            go = *h * 2 >= ns;
            rest.data = (const uint8_t *)keys->data + ns * 16;
            rest.n = keys->n - ns;
            out_rest = out + ns;
        }
        if (go && rest.n) {
            const PartPlan pl = plan_partition(b, rest.n, true, mode == 2 ? 2 : 0, true);
            if (pl.use) return check_partitioned(b, (const uint4 *)rest.data, rest.n, out_rest, pl);
        }
        if (rest.n == 0) return PB_OK;
    }
    CheckArgs a;
    a.b = b;
    a.out_dev = out_on_device ? out_rest : nullptr;
    a.out_host = out_on_device ? nullptr : out_rest;
    PB_TRY(for_each_chunk(ctx, &rest, check_chunk, &a));
    if (!out_on_device) PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

static int stage_hashes(pb_ctx *ctx, const uint64_t *hashes, uint64_t count, int on_device, const uint64_t **dev) {
    if (on_device) {
        *dev = hashes;
        return PB_OK;
    }
    PB_TRY(scratch_reserve(ctx, ctx->aux_stage[0], count * 8));
    PB_CUDA(cudaMemcpyAsync(ctx->aux_stage[0].p, hashes, count * 8, cudaMemcpyHostToDevice, ctx->stream));
    *dev = (const uint64_t *)ctx->aux_stage[0].p;
    return PB_OK;
}

int pb_bloom_add_hashes(pb_bloom *b, const uint64_t *hashes, uint64_t n, int on_device) {
    PB_REQUIRE(b && (hashes || n == 0), "NULL argument");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = b->ctx;
    DeviceGuard g(ctx->device);
    const uint64_t *d;
    PB_TRY(stage_hashes(ctx, hashes, n * b->k, on_device, &d));
    bloom_add_hashes_kernel<<<grid_for(ctx, n * b->k, 256, 8), 256, 0, ctx->stream>>>(d, n * b->k, dev_view(b));
    PB_TRY(check_launch(ctx, "bloom_add_hashes"));
    if (!on_device) PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

int pb_bloom_check_hashes(pb_bloom *b, const uint64_t *hashes, uint64_t n, int on_device, uint8_t *out, int out_on_device) {
    PB_REQUIRE(b && ((hashes && out) || n == 0), "NULL argument");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = b->ctx;
    DeviceGuard g(ctx->device);
    const uint64_t *d;
    PB_TRY(stage_hashes(ctx, hashes, n * b->k, on_device, &d));
    uint8_t *o = out;
    if (!out_on_device) {
        PB_TRY(scratch_reserve(ctx, ctx->out_stage[0], n));
        o = (uint8_t *)ctx->out_stage[0].p;
    }
    bloom_check_hashes_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(d, n, dev_view(b), o);
    PB_TRY(check_launch(ctx, "bloom_check_hashes"));
    if (!out_on_device) {
        PB_CUDA(cudaMemcpyAsync(out, o, n, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}

int pb_bloom_popcount(pb_bloom *b, uint64_t *out) {
    PB_REQUIRE(b && out, "NULL argument");
    pb_ctx *ctx = b->ctx;
    DeviceGuard g(ctx->device);
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    unsigned long long *acc = (unsigned long long *)ctx->small.p;
    PB_CUDA(cudaMemsetAsync(acc, 0, 8, ctx->stream));
    popcount_kernel<<<grid_for(ctx, b->nwords / 4, 256, 8), 256, 0, ctx->stream>>>((const uint4 *)b->words, b->nwords / 4, acc);
    PB_TRY(check_launch(ctx, "popcount"));
    PB_CUDA(cudaMemcpyAsync(ctx->pinned_small, acc, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = *(uint64_t *)ctx->pinned_small;
    return PB_OK;
}

static int same_shape(const pb_bloom *a, const pb_bloom *b) {
    return a->num_bits == b->num_bits && a->lo_bit == b->lo_bit && a->hi_bit == b->hi_bit && a->ctx->device == b->ctx->device;
}

// BloomFilter.union / intersection (bloom.py:371-428): dst = a | b (op 0) or a & b (op 1); dst may be a or b
int pb_bloom_combine(pb_bloom *dst, pb_bloom *a, pb_bloom *b, int op) {
    PB_REQUIRE(dst && a && b, "NULL argument");
    PB_REQUIRE(op == 0 || op == 1, "op must be 0 (union) or 1 (intersection)");
    PB_REQUIRE(same_shape(dst, a) && same_shape(dst, b), "Bloom Filters are not similar");
    pb_ctx *ctx = dst->ctx;
    DeviceGuard g(ctx->device);
    if (a->ctx != ctx) PB_CUDA(cudaStreamSynchronize(a->ctx->stream));
    if (b->ctx != ctx) PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    launch_begin(ctx);
    bloom_combine_kernel<<<grid_for(ctx, dst->nwords / 4, 256, 8), 256, 0, ctx->stream>>>((uint4 *)dst->words, (const uint4 *)a->words,
                                                                                          (const uint4 *)b->words, dst->nwords / 4, op);
    return check_launch(ctx, "bloom_combine");
}

// counts[0] = bits set in a | b, counts[1] = bits set in a & b (jaccard_index, bloom.py:430-460)
int pb_bloom_pair_popcounts(pb_bloom *a, pb_bloom *b, uint64_t *counts) {
    PB_REQUIRE(a && b && counts, "NULL argument");
    PB_REQUIRE(same_shape(a, b), "Bloom Filters are not similar");
    pb_ctx *ctx = a->ctx;
    DeviceGuard g(ctx->device);
    if (b->ctx != ctx) PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    unsigned long long *acc = (unsigned long long *)ctx->small.p + 40;
    PB_CUDA(cudaMemsetAsync(acc, 0, 16, ctx->stream));
    launch_begin(ctx);
    bloom_pair_popcount_kernel<<<grid_for(ctx, a->nwords / 4, 256, 8), 256, 0, ctx->stream>>>((const uint4 *)a->words,
                                                                                              (const uint4 *)b->words, a->nwords / 4, acc);
    PB_TRY(check_launch(ctx, "bloom_pair_popcount"));
    uint64_t *h = (uint64_t *)((uint8_t *)ctx->pinned_small + 512);
    PB_CUDA(cudaMemcpyAsync(h, acc, 16, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    counts[0] = h[0];
    counts[1] = h[1];
    return PB_OK;
}

int pb_bloom_route_keys(pb_ctx *ctx, const pb_keys *keys, uint64_t num_bits, uint32_t k, uint64_t shard_bits,
                        uint32_t n_shards, uint64_t *out_idx_dev, uint64_t slot_cap, uint64_t *counts_dev) {
    PB_REQUIRE(ctx && keys && out_idx_dev && counts_dev, "NULL argument");
    PB_REQUIRE(keys->on_device, "pb_bloom_route_keys takes device keys");
    PB_REQUIRE(n_shards >= 1 && n_shards <= 64, "n_shards must be in 1..64");
    PB_REQUIRE(shard_bits >= 1 && shard_bits * n_shards >= num_bits, "shards do not cover the filter");
    PB_REQUIRE(keys->offsets == nullptr && keys->sym_width == 1 && keys->stride == 16 && ((uintptr_t)keys->data & 15u) == 0,
               "routing currently takes fixed 16-byte keys");
    PB_REQUIRE(k >= 1, "k must be >= 1");
    DeviceGuard g(ctx->device);
    if (keys->n == 0) return PB_OK;
    FastMod fm = make_fastmod(num_bits);
    FastMod sd = make_fastmod(shard_bits);
    int st = PB_OK;
    const int kg = pick_group(k);
    const int grid = grid_for(ctx, keys->n, 256, 6);
    launch_begin(ctx);
#define CALL(K)                                                                                                       \
    (bloom_route_fixed16<K><<<grid, 256, 0, ctx->stream>>>((const uint4 *)keys->data, keys->n, fm, k, sd, shard_bits, \
                                                           n_shards, out_idx_dev, slot_cap,                           \
                                                           (unsigned long long *)counts_dev),                         \
     check_launch(ctx, "bloom_route"))
    PB_DISPATCH_KG(kg, CALL)
#undef CALL
    return st;
}

// Global bit indices of device-resident keys in key order (any key layout): out_idx_dev[i*k + s].
int pb_bloom_index_keys(pb_ctx *ctx, const pb_keys *keys, uint64_t num_bits, uint32_t k, uint64_t *out_idx_dev) {
    PB_REQUIRE(ctx && keys && (out_idx_dev || keys->n == 0), "NULL argument");
    PB_REQUIRE(keys->on_device, "pb_bloom_index_keys takes device keys");
    PB_REQUIRE(k >= 1 && num_bits >= 1, "k and num_bits must be >= 1");
    PB_TRY(validate_keys(keys));
    if (keys->n == 0) return PB_OK;
    DeviceGuard g(ctx->device);
    const FastMod fm = make_fastmod(num_bits);
    DevKeys dk;
    dk.data = (const uint8_t *)keys->data;
    dk.offsets = keys->offsets;
    dk.n = keys->n;
    dk.stride = keys->stride;
    dk.sym_width = keys->sym_width;
    dk.base_symbol = 0;
    dk.total_bytes = keys->offsets ? 0 : keys->n * (uint64_t)keys->stride * keys->sym_width;
    if (is_fixed16(dk)) {
        int st = PB_OK;
        const int grid = grid_for(ctx, keys->n, 256, 8);
        launch_begin(ctx);
#define CALL(K) (bloom_index_fixed16<K><<<grid, 256, 0, ctx->stream>>>((const uint4 *)dk.data, dk.n, fm, k, out_idx_dev), check_launch(ctx, "bloom_index"))
        PB_DISPATCH_KG(pick_group(k), CALL)
#undef CALL
        return st;
    }
    PB_TRY(hash_dev_keys(ctx, dk, k, out_idx_dev));
    reduce_mod_kernel<<<grid_for(ctx, keys->n * k, 256, 8), 256, 0, ctx->stream>>>(out_idx_dev, keys->n * k, fm);
    return check_launch(ctx, "bloom_index");
}

// out_dev[i] = AND over the k per-index answers of key i (bits_dev[i*k + s], 0/1 bytes)
int pb_bloom_and_rows(pb_ctx *ctx, const uint8_t *bits_dev, uint64_t n, uint32_t k, uint8_t *out_dev) {
    PB_REQUIRE(ctx && ((bits_dev && out_dev) || n == 0) && k >= 1, "bad argument");
    if (n == 0) return PB_OK;
    DeviceGuard g(ctx->device);
    and_rows_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(bits_dev, n, k, out_dev);
    return check_launch(ctx, "bloom_and_rows");
}

// Layout of the multi-GPU exchange for chunks of up to n_keys keys per rank (pb_p2p_create takes the two numbers):
// sublists per window = CTAs of the pass-1 launch, entries per sublist = a CTA's expected share + 7 sigma.
int pb_bloom_partition_layout(pb_ctx *ctx, uint64_t n_keys, uint32_t k, uint64_t num_bits, uint32_t window_log2,
                              uint32_t n_windows, uint32_t *n_sub, uint32_t *sub_cap) {
    PB_REQUIRE(ctx && n_sub && sub_cap && n_keys >= 1 && k >= 1 && k <= kMaxPartK, "bad argument (k must be 1..%u)", kMaxPartK);
    PB_REQUIRE(window_log2 >= 5 && window_log2 <= 31 && n_windows >= 1 && n_windows <= (uint32_t)kMaxWindows2, "bad window geometry");
    const PartLayout L = part_layout(ctx, n_keys, k, num_bits, window_log2, n_windows, true, true);
    *n_sub = (uint32_t)L.grid;
    *sub_cap = L.sub_cap;
    return PB_OK;
}

int pb_bloom_add_bit_indices(pb_bloom *b, const uint64_t *idx_dev, uint64_t n) {
    PB_REQUIRE(b && (idx_dev || n == 0), "NULL argument");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = b->ctx;
    DeviceGuard g(ctx->device);
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    unsigned long long *stray = (unsigned long long *)ctx->small.p + 8;
    PB_CUDA(cudaMemsetAsync(stray, 0, 8, ctx->stream));
    launch_begin(ctx);
    bloom_add_idx_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(idx_dev, n, dev_view(b), stray);
    PB_TRY(check_launch(ctx, "bloom_add_bit_indices"));
    PB_CUDA(cudaMemcpyAsync(ctx->pinned_small, stray, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint64_t bad = *(uint64_t *)ctx->pinned_small;
    PB_REQUIRE(bad == 0, "%llu bit indices fell outside this shard", (unsigned long long)bad);
    return PB_OK;
}

int pb_bloom_test_bit_indices(pb_bloom *b, const uint64_t *idx_dev, uint64_t n, uint8_t *out_dev) {
    PB_REQUIRE(b && ((idx_dev && out_dev) || n == 0), "NULL argument");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = b->ctx;
    DeviceGuard g(ctx->device);
    bloom_test_idx_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(idx_dev, n, dev_view(b), out_dev);
    return check_launch(ctx, "bloom_test_bit_indices");
}

// every index must fall into the bits this handle owns (the kernels below subtract lo_bit unchecked)
static int rows_in_range(pb_bloom *b, const uint64_t *idx_dev, uint64_t n_entries) {
    pb_ctx *ctx = b->ctx;
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    unsigned long long *stray = (unsigned long long *)ctx->small.p + 8;
    PB_CUDA(cudaMemsetAsync(stray, 0, 8, ctx->stream));
    bloom_rows_in_range_kernel<<<grid_for(ctx, n_entries, 256, 8), 256, 0, ctx->stream>>>(idx_dev, n_entries, dev_view(b), stray);
    PB_TRY(check_launch(ctx, "bloom_rows_in_range"));
    PB_CUDA(cudaMemcpyAsync(ctx->pinned_small, stray, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint64_t bad = *(uint64_t *)ctx->pinned_small;
    PB_REQUIRE(bad == 0, "%llu bit indices fall outside the bits of this filter", (unsigned long long)bad);
    return PB_OK;
}

// novel_dev[i] = 1 iff the reference, adding the rows one at a time in order with "add only if not found"
// (expandingbloom.py:166-169), would add row i to THIS filter.  Rows with skip_dev[i] != 0 take no part (found in an
// older filter of the stack).  The filter itself is not modified.
int pb_bloom_novel_rows(pb_bloom *b, const uint64_t *idx_dev, uint64_t n, const uint8_t *skip_dev, uint8_t *novel_dev) {
    PB_REQUIRE(b && ((idx_dev && novel_dev) || n == 0), "NULL argument");
    PB_REQUIRE(n <= (1ull << 24), "at most 2^24 rows per call (24-bit row numbers in the first-setter table)");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = b->ctx;
    DeviceGuard g(ctx->device);
    const uint64_t n_entries = n * b->k;
    PB_TRY(rows_in_range(b, idx_dev, n_entries));
    const size_t table_bytes = b->nwords * 32 * sizeof(uint32_t);
    if (!b->first_setter) {
        PB_CUDA(cudaMalloc(&b->first_setter, table_bytes));
        b->first_epoch = 0;
    }
    if (b->first_epoch == 0 || b->first_epoch >= 254) {  // fresh table, or the epochs have run out: start over
        PB_CUDA(cudaMemsetAsync(b->first_setter, 0xFF, table_bytes, ctx->stream));
        b->first_epoch = 0;
    }
    const uint32_t tag = (0xFFu - ++b->first_epoch) << 24;
    const BloomDev bd = dev_view(b);
    launch_begin(ctx);
    bloom_first_setter_kernel<<<grid_for(ctx, n_entries, 256, 8), 256, 0, ctx->stream>>>(idx_dev, n_entries, b->k, bd, skip_dev, b->first_setter, tag);
    PB_TRY(check_launch(ctx, "bloom_first_setter"));
    launch_begin(ctx);
    bloom_novel_rows_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(idx_dev, n, b->k, bd, skip_dev, b->first_setter, tag, novel_dev);
    return check_launch(ctx, "bloom_novel_rows");
}

// BloomFilter.add_alt (bloom.py:241-250) for the rows whose mask byte is non-zero (NULL: all rows); the caller keeps
// elements_added
int pb_bloom_add_rows(pb_bloom *b, const uint64_t *idx_dev, uint64_t n, const uint8_t *mask_dev) {
    PB_REQUIRE(b && (idx_dev || n == 0), "NULL argument");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = b->ctx;
    DeviceGuard g(ctx->device);
    const uint64_t n_entries = n * b->k;
    PB_TRY(rows_in_range(b, idx_dev, n_entries));
    launch_begin(ctx);
    bloom_add_rows_kernel<<<grid_for(ctx, n_entries, 256, 8), 256, 0, ctx->stream>>>(idx_dev, n_entries, b->k, dev_view(b), mask_dev);
    return check_launch(ctx, "bloom_add_rows");
}

// hands the first-setter table of pb_bloom_novel_rows from a filter that stopped being the newest of its stack to
// its successor (same geometry).  No reset is needed: the epoch counter travels with the table, so everything the
// old filter's calls left behind is "an earlier call" to the new one.  Saves a 4-byte-per-bit cudaFree + cudaMalloc
// per growth step (allocation cost, not GPU time, is what a stack that grows often is bound by).
int pb_bloom_move_scratch(pb_bloom *from, pb_bloom *to) {
    PB_REQUIRE(from && to && from != to, "bad argument");
    PB_REQUIRE(from->ctx == to->ctx && from->nwords == to->nwords && from->lo_bit == to->lo_bit && from->hi_bit == to->hi_bit,
               "pb_bloom_move_scratch takes two filters of one context with the same geometry");
    if (!from->first_setter) return PB_OK;
    DeviceGuard g(to->ctx->device);
    if (to->first_setter) {
        PB_CUDA(cudaStreamSynchronize(to->ctx->stream));
        PB_CUDA(cudaFree(to->first_setter));
    }
    to->first_setter = from->first_setter;
    to->first_epoch = from->first_epoch;
    from->first_setter = nullptr;
    from->first_epoch = 0;
    return PB_OK;
}

// drops the first-setter table of pb_bloom_novel_rows (4 bytes per bit; a filter that stopped being the newest of
// its stack never needs it again)
int pb_bloom_release_scratch(pb_bloom *b) {
    PB_REQUIRE(b, "handle is NULL");
    if (!b->first_setter) return PB_OK;
    DeviceGuard g(b->ctx->device);
    PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    PB_CUDA(cudaFree(b->first_setter));
    b->first_setter = nullptr;
    return PB_OK;
}

// found_dev[i] = 1 iff one of the n_filters filters (same context, num_bits and k, unsharded) holds every bit of row i
int pb_bloom_rows_in_any(pb_bloom *const *filters, uint32_t n_filters, const uint64_t *idx_dev, uint64_t n, uint8_t *found_dev) {
    PB_REQUIRE(filters && n_filters >= 1 && ((idx_dev && found_dev) || n == 0), "bad argument");
    if (n == 0) return PB_OK;
    pb_bloom *b0 = filters[0];
    PB_REQUIRE(b0, "filter handle is NULL");
    pb_ctx *ctx = b0->ctx;
    DeviceGuard g(ctx->device);
    std::vector<const uint32_t *> ptrs(n_filters);
    for (uint32_t f = 0; f < n_filters; ++f) {
        pb_bloom *b = filters[f];
        PB_REQUIRE(b && b->ctx == ctx && b->num_bits == b0->num_bits && b->k == b0->k && b->lo_bit == 0 && b->hi_bit == b->num_bits,
                   "pb_bloom_rows_in_any takes whole filters of one context with equal num_bits and k");
        ptrs[f] = b->words;
    }
    PB_TRY(rows_in_range(b0, idx_dev, n * b0->k));
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096 + (size_t)n_filters * 8));
    const uint32_t **d_ptrs = (const uint32_t **)((uint8_t *)ctx->small.p + 4096);
    PB_CUDA(cudaMemcpyAsync(d_ptrs, ptrs.data(), (size_t)n_filters * 8, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));  // `ptrs` is pageable host memory of this frame
    launch_begin(ctx);
    bloom_rows_in_any_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(idx_dev, n, b0->k, d_ptrs, n_filters, found_dev);
    return check_launch(ctx, "bloom_rows_in_any");
}

}  // extern "C"
