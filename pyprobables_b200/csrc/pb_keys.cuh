// pb_keys.cuh -- device-side key access and the k-seed FNV-1a evaluation shared by every kernel.
//
// Two ways a thread gets at its key:
//   * fixed 16-byte keys (the BASELINE workload): one coalesced LDG.128 per thread, bytes stay in
//     registers, KG independent FNV chains run interleaved for ILP;
//   * everything else (other strides, variable-length packed keys, u32 code-point symbols): the CTA's
//     contiguous byte range is staged into shared memory with one TMA bulk copy
//     (cp.async.bulk ... mbarrier::complete_tx) and every thread walks its key out of smem;
//     oversize tiles fall back to reading global memory directly.
// Reference semantics: probables/hashes.py:71-103 (seed s starts at basis + 31*s; per symbol xor then
// multiply by the 64-bit FNV prime).
#pragma once
#include "pb_common.cuh"
#include "pb_hash.cuh"

namespace pb {

constexpr int kMaxGroup = 8;  // FNV chains evaluated together by one thread

// group size for k hashes: minimise wasted chains, prefer wider groups (more ILP)
inline int pick_group(uint32_t k) {
    if (k <= (uint32_t)kMaxGroup) return (int)k;
    int best = kMaxGroup;
    uint32_t best_total = ((k + kMaxGroup - 1) / kMaxGroup) * kMaxGroup;
    for (int g = kMaxGroup - 1; g >= 4; --g) {
        uint32_t total = ((k + g - 1) / g) * g;
        if (total < best_total) {
            best_total = total;
            best = g;
        }
    }
    return best;
}

// ---- fixed 16-byte keys held in registers ---------------------------------------------------------
template <int KG>
__device__ __forceinline__ void fnv_group_16(const uint4 &w, uint32_t seed0, uint64_t (&h)[KG]) {
    uint32_t lo[KG], hi[KG];
#pragma unroll
    for (int j = 0; j < KG; ++j) {
        const uint64_t h0 = fnv_init(seed0 + j);
        lo[j] = (uint32_t)h0;
        hi[j] = (uint32_t)(h0 >> 32);
    }
    const uint32_t words[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t sym = __byte_perm(words[wi], 0u, 0x4440 + b);  // byte b, zero extended (one PRMT)
#pragma unroll
            for (int j = 0; j < KG; ++j) fnv_step_halves(lo[j], hi[j], sym);
        }
    }
#pragma unroll
    for (int j = 0; j < KG; ++j) h[j] = ((uint64_t)hi[j] << 32) | lo[j];
}

// ---- generic keys: symbols behind a pointer (shared or global) -------------------------------------
template <int KG, int SYMW>
__device__ __forceinline__ void fnv_group_ptr(const uint8_t *p, uint32_t len, uint32_t seed0, uint64_t (&h)[KG]) {
#pragma unroll
    for (int j = 0; j < KG; ++j) h[j] = fnv_init(seed0 + j);
    if (SYMW == 4) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
        for (uint32_t i = 0; i < len; ++i) {
            const uint32_t sym = q[i];
#pragma unroll
            for (int j = 0; j < KG; ++j) h[j] = fnv_step(h[j], sym);
        }
    } else {
        uint32_t i = 0;
        // head bytes up to 4-byte alignment, then whole words, then the tail
        while (i < len && ((uintptr_t)(p + i) & 3u) != 0) {
            const uint32_t sym = p[i++];
#pragma unroll
            for (int j = 0; j < KG; ++j) h[j] = fnv_step(h[j], sym);
        }
        for (; i + 4 <= len; i += 4) {
            const uint32_t w = *reinterpret_cast<const uint32_t *>(p + i);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t sym = (w >> (8 * b)) & 0xFFu;
#pragma unroll
                for (int j = 0; j < KG; ++j) h[j] = fnv_step(h[j], sym);
            }
        }
        for (; i < len; ++i) {
            const uint32_t sym = p[i];
#pragma unroll
            for (int j = 0; j < KG; ++j) h[j] = fnv_step(h[j], sym);
        }
    }
}

// ---- TMA bulk staging of a tile's bytes ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

constexpr uint32_t kTileKeys = 256;         // keys per CTA tile on the staged path (= blockDim.x)
constexpr uint32_t kStageBytes = 40 * 1024;  // smem staging buffer per CTA

struct TileSmem {
    alignas(128) uint8_t buf[kStageBytes];
    alignas(8) uint64_t bar;
};

// Per-thread view of one key inside a tile.
struct KeyRef {
    const uint8_t *p;
    uint32_t len;  // symbols
};

// Stage keys [first, first+count) of dk (count <= blockDim.x) into `buf` (cap_bytes, 128-byte aligned shared
// memory) and return this thread's key (thread t owns key first+t; threads beyond count get len 0 and must not use
// the key).  All threads of the CTA must call this; `parity` is the CTA-uniform mbarrier phase, flipped on use.
template <int SYMW>
__device__ __forceinline__ KeyRef stage_tile_buf(const DevKeys &dk, uint64_t first, uint32_t count, uint8_t *buf,
                                                 uint32_t cap_bytes, uint64_t *bar, uint32_t &parity) {
    const uint32_t t = threadIdx.x;
    uint64_t sym_beg, sym_end, my_beg = 0, my_end = 0;
    if (dk.offsets) {
        sym_beg = dk.offsets[first] - dk.base_symbol;
        sym_end = dk.offsets[first + count] - dk.base_symbol;
        if (t < count) {
            my_beg = dk.offsets[first + t] - dk.base_symbol;
            my_end = dk.offsets[first + t + 1] - dk.base_symbol;
        }
    } else {
        sym_beg = first * (uint64_t)dk.stride;
        sym_end = (first + count) * (uint64_t)dk.stride;
        if (t < count) {
            my_beg = sym_beg + (uint64_t)t * dk.stride;
            my_end = my_beg + dk.stride;
        }
    }
    const uint8_t *g_beg = dk.data + sym_beg * SYMW;
    const uint8_t *g_end = dk.data + sym_end * SYMW;
    const uint8_t *a_beg = (const uint8_t *)((uintptr_t)g_beg & ~(uintptr_t)15);
    const uint8_t *a_end = (const uint8_t *)((uintptr_t)g_end & ~(uintptr_t)15);
    const uint64_t span = (uint64_t)(g_end - a_beg);
    KeyRef r;
    r.len = (uint32_t)(my_end - my_beg);
    if (span > cap_bytes || g_end <= g_beg) {
        // oversize (or empty) tile: read straight from global memory
        r.p = dk.data + my_beg * SYMW;
        return r;
    }
    __syncthreads();  // every thread is done reading the previous tile out of buf
    const uint32_t bulk = a_end > a_beg ? (uint32_t)(a_end - a_beg) : 0u;
    if (bulk) {
        if (t == 0) {
            fence_proxy_async_smem();  // order prior generic-proxy reads of buf before the async-proxy write
            mbar_expect_tx(bar, bulk);
            tma_bulk_g2s(buf, a_beg, bulk, bar);
        }
    }
    // the < 16 tail bytes (and the whole span when it never reaches a 16-byte boundary)
    const uint8_t *tail_src = bulk ? a_end : a_beg;
    const uint32_t tail_off = bulk;
    const uint32_t tail_n = (uint32_t)(g_end - tail_src);
    for (uint32_t i = t; i < tail_n; i += blockDim.x) buf[tail_off + i] = tail_src[i];
    if (bulk) {
        mbar_wait(bar, parity);
        parity ^= 1u;
    }
    __syncthreads();
    r.p = buf + (uint32_t)(g_beg - a_beg) + (uint32_t)((my_beg - sym_beg) * SYMW);
    return r;
}

template <int SYMW>
__device__ __forceinline__ KeyRef stage_tile(const DevKeys &dk, uint64_t first, uint32_t count, TileSmem &sm,
                                             uint32_t &parity) {
    return stage_tile_buf<SYMW>(dk, first, count, sm.buf, kStageBytes, &sm.bar, parity);
}

// True when the fast register path applies.
inline bool is_fixed16(const DevKeys &dk) {
    return dk.offsets == nullptr && dk.sym_width == 1 && dk.stride == 16 && ((uintptr_t)dk.data & 15u) == 0;
}

}  // namespace pb
