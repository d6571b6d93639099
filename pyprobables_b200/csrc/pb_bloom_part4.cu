// pb_bloom_part4.cu -- pass 1 of the partitioned Bloom insert, fourth generation (see pb_bloom_part.cuh for the
// scheme: hash -> % m -> bin the bit indices by L2-sized bitmap window; reference semantics bloom.py:234-250).
#include "pb_bloom_part.cuh"

namespace pb {

// ---- fourth generation (round 2).  The round-1 ncu capture of bloom_part3 showed an ALU-bound kernel: 1219
// thread instructions per key of which only 448 are the irreducible FNV steps, top stall "math pipe throttle".
// What is different here, each item removing integer-pipe instructions per bit index:
//   * the number of hashes K is a template parameter (no runtime `s < p.k` predicates);
//   * phase B builds ONE 8-byte table entry per window {start of its run in the sorted tile, list position minus
//     run start}, so the scatter into shared memory is one LDS.64 + one STS.64 per index and the copy-out is one
//     LDS.64, one add and one STG per index -- no window id is stored or looked up again, no bounds test per
//     index (a tile whose lists would overflow takes a slow path that is chosen once per tile);
//   * phase B gives every window its own thread; a warp gets the offset of its 32 windows from a strided read of
//     the histogram in front of them + one warp reduction, so there is neither a single-warp serial section nor a
//     second barrier, however many windows there are;
//   * the keys come through a KeySource: 16-byte keys in registers (LDG.128, prefetched one tile ahead) or any
//     other shape staged per tile into shared memory with one TMA bulk copy (stage_tile), which gives
//     variable-length / str batches the partitioned path too.
// Phases per tile: A hash + histogram (smem atomics return the rank) | B scan + list reservation | C scatter into
// window order | D coalesced copy-out.  Three barriers per tile; the histogram ping-pongs so zeroing is free.
struct Part4Tab {
    uint32_t run_bytes;  // byte offset in sorted[] of the window's run (8 bytes per entry)
    uint32_t gdelta;     // (w * cap + position of the run in the window list) - first entry of the run
};

template <int K>
struct PartGroups {
    static constexpr int NG = K <= 8 ? 1 : 2;
    static constexpr int KG = (K + NG - 1) / NG;
};

// fixed 16-byte keys: one LDG.128 per thread, next tile prefetched
struct KeySrcFixed16 {
    const uint4 *keys;
    static constexpr bool kStaged = false;
};
// any other layout: TMA-staged tile (pb_keys.cuh)
template <int SYMW>
struct KeySrcStaged {
    DevKeys dk;
    static constexpr bool kStaged = true;
    static constexpr int kSymW = SYMW;
};

constexpr uint32_t kPart4StageBytes = 24 * 1024;  // staged variant: bytes of key data per 256-key tile kept in smem
constexpr uint32_t kRankStep = 8;                 // the histogram counts in units of one sorted[] entry (8 bytes)

// One bit index of the tile: reduce, split into window / window-local index, take a rank in the window's histogram.
// The histogram word of window w starts at w << 16, so the atomic's return value already is
// (window << 16 | rank * 8): nothing to pack.
template <bool FAST33>
__device__ __forceinline__ void part4_bin(uint64_t h, const PartDev &p, uint32_t mask, uint32_t *hist, uint32_t &loc, uint32_t &wr) {
    const uint64_t idx = FAST33 ? mod_fast33(h, p) : mod_any(h, p);
    const uint32_t w = __funnelshift_r((uint32_t)idx, (uint32_t)(idx >> 32), p.window_log2);  // window_log2 <= 31
    loc = (uint32_t)idx & mask;
    wr = atomicAdd(hist + w, kRankStep);
}

// IDS: partitioned query -- every staged index also carries the number of its key inside the chunk (p.ids) and an
// index that overflows its sublist is tested directly (p.out) instead of being ORed in
template <int K, int BS, class KS, bool IDS = false>
__global__ void __maxnreg__(K <= 8 ? 56 : 72) bloom_part4(KS src, uint64_t n, PartDev p) {
    constexpr int NG = PartGroups<K>::NG, KG = PartGroups<K>::KG;
    constexpr uint32_t NW = BS / 32;
    static_assert(BS * K * kRankStep <= 65536, "rank * 8 must fit the low 16 bits of a histogram word");
    __shared__ uint32_t hist[2][kMaxWindows2];  // window << 16 | entries of the tile so far * 8
    __shared__ Part4Tab tab[kMaxWindows2];
    __shared__ uint32_t cur[kMaxWindows2];  // entries written so far to this CTA's sublist of each window
    __shared__ uint2 sorted[BS * K];  // {window-local bit index, gdelta}; slow path: {index, window}
    __shared__ uint32_t tile_flags[2];  // [0]: a list of this tile overflows -> slow path; [1]: entries in the tile
    __shared__ uint16_t sorted_id[IDS ? BS * K : 1];  // IDS: the thread (= key of the tile) every sorted entry came from
    extern __shared__ __align__(128) uint8_t dyn_smem[];  // staged key bytes (KeySrcStaged only)
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t W = p.n_windows;
    const uint32_t mask = (1u << p.window_log2) - 1u;
    for (uint32_t w = tid; w < W; w += BS) {
        hist[0][w] = w << 16;
        hist[1][w] = w << 16;
        cur[w] = 0;
    }
    if (tid == 0) tile_flags[0] = 0;
    uint64_t *bar = reinterpret_cast<uint64_t *>(dyn_smem + kPart4StageBytes);
    if constexpr (KS::kStaged) {
        if (tid == 0) mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t tiles = (n + BS - 1) / BS;
    const bool fast33 = p.fast33 != 0;
    uint32_t pp = 0, parity = 0;
    uint64_t tile = blockIdx.x;
    uint4 nextk = make_uint4(0, 0, 0, 0);
    if constexpr (!KS::kStaged) {
        if (tile < tiles && tile * BS + tid < n) nextk = __ldcs(src.keys + tile * BS + tid);
    }
    for (; tile < tiles; tile += gridDim.x) {
        const uint64_t i = tile * BS + tid;
        const bool live = i < n;
        uint32_t loc[K], wr[K];  // window-local index; window << 16 | rank * 8 within the tile
        uint32_t *const hcur = hist[pp];
        // ---- phase A
        KeyRef kr{nullptr, 0};
        uint4 kw = make_uint4(0, 0, 0, 0);
        if constexpr (!KS::kStaged) {
            kw = nextk;
            const uint64_t ni = (tile + gridDim.x) * BS + tid;
            if (ni < n) nextk = __ldcs(src.keys + ni);
        } else {
            const uint32_t count = (uint32_t)min((uint64_t)BS, n - tile * BS);
            kr = stage_tile_buf<KS::kSymW>(src.dk, tile * BS, count, dyn_smem, kPart4StageBytes, bar, parity);
        }
        if (live) {
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                uint64_t h[KG];
                if constexpr (!KS::kStaged) fnv_group_16<KG>(kw, g * KG, h);
                else fnv_group_ptr<KG, KS::kSymW>(kr.p, kr.len, g * KG, h);
                if (fast33) {  // CTA-uniform
#pragma unroll
                    for (int j = 0; j < KG; ++j)
                        if (g * KG + j < K) part4_bin<true>(h[j], p, mask, hcur, loc[g * KG + j], wr[g * KG + j]);
                } else {
#pragma unroll
                    for (int j = 0; j < KG; ++j)
                        if (g * KG + j < K) part4_bin<false>(h[j], p, mask, hcur, loc[g * KG + j], wr[g * KG + j]);
                }
            }
        }
        __syncthreads();
        // ---- phase B: thread t owns window t (round r: window r*BS + t).  A warp's offset is the sum of the
        // histogram in front of it, read redundantly (W/32 loads per lane + one warp reduction) instead of a second
        // barrier; warps whose 32 windows lie beyond W skip the phase.  (The round-2 ncu capture of the first
        // version, where every warp scanned a slice with a few active lanes, spent 32 % of the kernel here.)
        for (uint32_t base = 0; base < W; base += BS) {
            const uint32_t wfirst = base + warp * 32;  // first window of this warp in this round
            if (wfirst >= W) break;
            const uint32_t w = wfirst + lane;
            const uint32_t mine = w < W ? (hcur[w] & 0xFFFFu) : 0u;  // bytes of sorted[] the window's run takes
            uint32_t before = 0;
            for (uint32_t x = lane; x < wfirst; x += 32) before += hcur[x] & 0xFFFFu;
#pragma unroll
            for (int o = 16; o; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)lane >= o) incl += v;
            }
            const uint32_t run8 = before + incl - mine;  // byte offset of the run in sorted[]
            if (w < W) {
                const uint32_t need = mine / kRankStep;
                const uint32_t c = cur[w];  // entries this CTA has written to its sublist of window w so far
                cur[w] = c + need;
                if (c + need > p.sub_cap) tile_flags[0] = 1;  // benign race: every writer stores 1
                tab[w] = Part4Tab{run8, (w * p.n_sub + blockIdx.x) * p.sub_cap + c - run8 / kRankStep};
                hist[pp ^ 1][w] = w << 16;  // last read in phase B of the previous tile
                if (w == W - 1) tile_flags[1] = (run8 + mine) / kRankStep;  // entries in the tile
            }
        }
        __syncthreads();
        const bool slow = tile_flags[0] != 0;
        const uint32_t total = tile_flags[1];
        uint8_t *const sorted_b = reinterpret_cast<uint8_t *>(sorted);
        // ---- phase C: scatter into window order
        if (live) {
            if (!slow) {
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    const Part4Tab t = tab[wr[s] >> 16];
                    const uint32_t off = t.run_bytes + (wr[s] & 0xFFFFu);
                    *reinterpret_cast<uint2 *>(sorted_b + off) = make_uint2(loc[s], t.gdelta);
                    if (IDS) sorted_id[off / kRankStep] = (uint16_t)tid;
                }
            } else {
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    const uint32_t w = wr[s] >> 16;
                    const uint32_t off = tab[w].run_bytes + (wr[s] & 0xFFFFu);
                    *reinterpret_cast<uint2 *>(sorted_b + off) = make_uint2(loc[s], w);
                    if (IDS) sorted_id[off / kRankStep] = (uint16_t)tid;
                }
            }
        }
        __syncthreads();
        // ---- phase D: coalesced copy-out (consecutive threads, consecutive list positions inside a run)
        if (!slow) {
            if (total == (uint32_t)(BS * K)) {
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    const uint32_t e = s * BS + tid;
                    const uint2 v = sorted[e];
                    __stcs(p.stage + (v.y + e), v.x);
                    if (IDS) __stcs(p.ids + (v.y + e), (uint32_t)(tile * BS) + sorted_id[e]);
                }
            } else {
                for (uint32_t e = tid; e < total; e += BS) {
                    const uint2 v = sorted[e];
                    __stcs(p.stage + (v.y + e), v.x);
                    if (IDS) __stcs(p.ids + (v.y + e), (uint32_t)(tile * BS) + sorted_id[e]);
                }
            }
        } else {
            for (uint32_t e = tid; e < total; e += BS) {
                const uint2 v = sorted[e];
                const Part4Tab t = tab[v.y];
                const uint32_t first = (v.y * p.n_sub + blockIdx.x) * p.sub_cap;  // entry number of the sublist's start
                const uint32_t pos = t.gdelta + e - first;                        // position inside the sublist
                if (pos < p.sub_cap) {
                    __stcs(p.stage + (size_t)first + pos, v.x);
                    if (IDS) __stcs(p.ids + (size_t)first + pos, (uint32_t)(tile * BS) + sorted_id[e]);
                } else if (IDS) {  // query: test the bit right here
                    const uint64_t idx = ((uint64_t)v.y << p.window_log2) | v.x;
                    if (!((p.words[idx >> 5] >> (uint32_t)(idx & 31)) & 1u)) p.out[(uint32_t)(tile * BS) + sorted_id[e]] = 0;
                } else {
                    part_overflow(p, ((uint64_t)v.y << p.window_log2) | v.x);
                }
            }
            __syncthreads();  // everyone has read tile_flags[0] before it is cleared
            if (tid == 0) tile_flags[0] = 0;
        }
        pp ^= 1;
    }
    __syncthreads();
    for (uint32_t w = tid; w < W; w += BS) p.counts[(size_t)w * p.n_sub + blockIdx.x] = min(cur[w], p.sub_cap);
}

// host-side launcher of bloom_part4 for a key batch of any layout: pd.n_sub CTAs of `block` threads.  512-key
// tiles only exist for K <= 8 (the sorted tile of a larger K would not fit the 48 KB of static shared memory) and
// for 16-byte keys (part_big_tile).
// Both passes ask for the largest shared-memory carveout: pass 2 (33 KB per CTA) has to fit NEXT TO the four resident
// pass-1 CTAs of the following chunk, and an SM cannot change its L1/shared split while CTAs are resident -- with
// the default heuristic split pass 1 left no room and the two passes silently ran one after the other.
template <class Kern>
static cudaError_t prefer_max_smem(Kern kern, size_t dyn) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess && dyn) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    return e;
}

template <int K, int BS, class KS, bool IDS = false>
static cudaError_t launch_part4_inst(const KS &src, size_t dyn, cudaStream_t stream, uint64_t n, const PartDev &pd) {
    static bool configured = false;  // per template instance
    if (!configured) {
        cudaError_t e = prefer_max_smem(bloom_part4<K, BS, KS, IDS>, dyn);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    bloom_part4<K, BS, KS, IDS><<<(int)pd.n_sub, BS, dyn, stream>>>(src, n, pd);
    return cudaSuccess;
}

// host-side launcher of bloom_part4 for a key batch of any layout: pd.n_sub CTAs of `block` threads.  512-key
// tiles only exist for K <= 8 (the sorted tile of a larger K would not fit the 48 KB of static shared memory) and
// for 16-byte keys (part_big_tile).
template <int K>
static cudaError_t launch_part4_k(int block, cudaStream_t stream, const DevKeys &dk, const PartDev &pd) {
    if (is_fixed16(dk)) {
        KeySrcFixed16 src{(const uint4 *)dk.data};
        if constexpr (K <= 8) {
            if (block == 512) return launch_part4_inst<K, 512>(src, 0, stream, dk.n, pd);
        }
        if (block != 256) return cudaErrorInvalidValue;
        return launch_part4_inst<K, 256>(src, 0, stream, dk.n, pd);
    }
    if (block != 256) return cudaErrorInvalidValue;
    const size_t dyn = kPart4StageBytes + 16;
    if (dk.sym_width == 4) return launch_part4_inst<K, 256>(KeySrcStaged<4>{dk}, dyn, stream, dk.n, pd);
    return launch_part4_inst<K, 256>(KeySrcStaged<1>{dk}, dyn, stream, dk.n, pd);
}

template <int K>
static cudaError_t launch_part4_ids_k(int block, cudaStream_t stream, const DevKeys &dk, const PartDev &pd) {
    // the key-number array of the sorted tile costs 2 bytes per entry of shared memory: 512-key tiles up to K = 7 and
    // 256-key tiles up to K = 15 stay inside the 48 KB of static shared memory (kMaxQueryK)
    if (!is_fixed16(dk) || dk.n > 0xFFFFFFFFull) return cudaErrorInvalidValue;
    KeySrcFixed16 src{(const uint4 *)dk.data};
    if constexpr (K <= 7) {
        if (block == 512) return launch_part4_inst<K, 512, KeySrcFixed16, true>(src, 0, stream, dk.n, pd);
    }
    if constexpr (K <= (int)kMaxQueryK) {
        if (block == 256) return launch_part4_inst<K, 256, KeySrcFixed16, true>(src, 0, stream, dk.n, pd);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_part4_ids(int block, cudaStream_t stream, const DevKeys &dk, const PartDev &pd) {
    switch (pd.k) {
#define PB_P4(KK) case KK: return launch_part4_ids_k<KK>(block, stream, dk, pd);
        PB_P4(1) PB_P4(2) PB_P4(3) PB_P4(4) PB_P4(5) PB_P4(6) PB_P4(7) PB_P4(8)
        PB_P4(9) PB_P4(10) PB_P4(11) PB_P4(12) PB_P4(13) PB_P4(14) PB_P4(15) PB_P4(16)
#undef PB_P4
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_part4(int block, cudaStream_t stream, const DevKeys &dk, const PartDev &pd) {
    switch (pd.k) {
#define PB_P4(KK) case KK: return launch_part4_k<KK>(block, stream, dk, pd);
        PB_P4(1) PB_P4(2) PB_P4(3) PB_P4(4) PB_P4(5) PB_P4(6) PB_P4(7) PB_P4(8)
        PB_P4(9) PB_P4(10) PB_P4(11) PB_P4(12) PB_P4(13) PB_P4(14) PB_P4(15) PB_P4(16)
#undef PB_P4
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace pb
