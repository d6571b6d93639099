// pb_cuckoo.cu -- CuckooFilter.add / check for whole batches
// (reference: probables/cuckoo/cuckoo.py:291-315, :361-392, :440-453, :483-506; utilities.py:32-35).
//
// State: u32 slots[capacity][bucket_size], 0 = empty.  The reference keeps a Python list per bucket;
// what `check` observes is only *which fingerprints are stored* -- both candidate buckets are a
// function of the fingerprint alone (:483-490), so membership does not depend on slot placement or on
// the random eviction choices (:373, :377).  That is the parity contract: the set of stored
// fingerprints, elements_added (= number of distinct fingerprints placed) and every check() result.
// Fingerprint 0 is legal (utilities.py:35) but equals "empty" in the flat layout (and in the
// reference's own export format, :346/:429); it is kept as a one-word device flag.
//
// add() for a batch runs as three kernels:
//   1. claim+place   hash key -> fp (:499-500); one thread per *distinct* fp of the batch wins a claim
//                    (an exact set in scratch memory: one atomicCAS / atomicOr per key); the winner
//                    looks the fp up in its two buckets (:300-302, :440-446) and, when absent, tries the
//                    first empty slot of idx_1, then idx_2 (:363-368) from the same bucket snapshots; only
//                    fingerprints whose two buckets are full go to a compact list for kernel 2.
//   2. evict         one thread per listed fp: (re-try both buckets, then) the eviction walk of <= max_swaps steps (:371-389) with
//                    atomicExch -- every fingerprint is always either in a slot or in exactly one
//                    thread's hand, so nothing is duplicated or lost.  Walks that run out of swaps hand
//                    their homeless fingerprint back to the caller (:392, :508-516).
//   3. the claim scratch is cleared again.  Small batches claim in an open-addressing set sized to the batch
//      (context scratch); only batches whose set would be larger use the 2^fp_bits-bit bitmap (allocated lazily).
// Because kernel 2 only ever sees fingerprints that are distinct and absent, no presence check has to
// race with an eviction in flight.
// "cuckoo_serial" (context option) replaces 1-3 by a one-thread kernel that inserts in key order, which
// reproduces the reference's append order (export bytes match at loads without evictions).
#include <algorithm>
#include <new>

#include "pb_common.cuh"
#include "pb_hash.cuh"
#include "pb_keys.cuh"

using namespace pb;

struct pb_cuckoo {
    pb_ctx *ctx = nullptr;
    uint64_t capacity = 0;
    uint32_t bucket_size = 0, max_swaps = 0, fp_bits = 0;
    uint64_t rng_seed = 0;
    uint64_t epoch = 0;  // bumps per insert launch so eviction choices differ between batches
    uint32_t *slots = nullptr;
    uint64_t nslots = 0;
    uint32_t *zero_flag = nullptr;  // device word: 1 when fingerprint 0 is stored
    uint64_t *alt = nullptr;        // pre-indexed mode: idx_2 per slot (allocated by the first pre-indexed call)
    FastMod fm;
    // CountingCuckooFilter (pb_cuckoo_counts_*): fingerprint -> count, open addressing, entry `cap` = fingerprint 0
    uint32_t *cnt_keys = nullptr, *cnt_vals = nullptr, *cnt_tickets = nullptr;
    uint64_t cnt_cap = 0;   // power of two (entries 0..cap-1 hashed, entry cap for fingerprint 0)
    uint64_t cnt_used = 0;  // keys ever claimed (entries whose count fell to 0 stay claimed until the next rebuild)
};

namespace pb {

struct CuckooDev {
    uint32_t *slots;
    uint32_t *zero_flag;
    uint32_t *claim;      // in-batch dedupe scratch: a 2^fp_bits-bit bitmap, or (claim_mask != 0) an open-addressing set
    uint32_t claim_mask;  // set mode: number of u32 entries - 1 (power of two, >= 4 x batch)
    uint64_t *alt;        // pre-indexed filters (custom hash_function): idx_2 of the fingerprint in every slot
    FastMod fm;
    uint32_t bucket_size, max_swaps, fp_bits;
};

// exactly one thread per distinct fingerprint of the batch gets `true` (fp != 0)
__device__ __forceinline__ bool claim_fp(const CuckooDev &c, uint32_t fp) {
    if (c.claim_mask == 0u) {
        const uint32_t bit = 1u << (fp & 31u);
        return (atomicOr(c.claim + (fp >> 5), bit) & bit) == 0u;
    }
    uint32_t slot = (fp * 0x9E3779B1u) & c.claim_mask;
    for (;;) {
        const uint32_t old = atomicCAS(c.claim + slot, 0u, fp);
        if (old == 0u) return true;
        if (old == fp) return false;
        slot = (slot + 1u) & c.claim_mask;
    }
}

struct CuckooCounters {  // device-resident, zeroed per call
    unsigned long long n_new;     // fingerprints handed to the insert kernel
    unsigned long long n_placed;  // fingerprints newly stored
    unsigned long long n_failed;  // homeless fingerprints
    unsigned long long cursor;    // next list entry the insert kernel hands out
};

__device__ __forceinline__ void cuckoo_buckets(const CuckooDev &c, uint32_t fp, uint64_t &i1, uint64_t &i2) {
    i1 = fastmod((uint64_t)fp, c.fm);       // cuckoo.py:488
    i2 = fastmod(fnv_of_decimal(fp), c.fm);  // cuckoo.py:489
}

// fp in bucket?  (cuckoo.py:442-445)
template <int BS>
__device__ __forceinline__ bool bucket_has(const CuckooDev &c, uint64_t b, uint32_t fp) {
    if (BS == 4) {
        const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(c.slots) + b);
        return v.x == fp || v.y == fp || v.z == fp || v.w == fp;
    }
    const uint32_t *s = c.slots + b * c.bucket_size;
    bool hit = false;
    for (uint32_t j = 0; j < c.bucket_size; ++j) hit |= (__ldcg(s + j) == fp);
    return hit;
}

// cuckoo.py:448-453: put fp into the first empty slot of bucket b
template <int BS>
__device__ __forceinline__ bool bucket_place(const CuckooDev &c, uint64_t b, uint32_t fp) {
    uint32_t *s = c.slots + b * c.bucket_size;
    if (BS == 4) {
        const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(c.slots) + b);
        if (v.x == 0 && atomicCAS(s + 0, 0u, fp) == 0u) return true;
        if (v.y == 0 && atomicCAS(s + 1, 0u, fp) == 0u) return true;
        if (v.z == 0 && atomicCAS(s + 2, 0u, fp) == 0u) return true;
        if (v.w == 0 && atomicCAS(s + 3, 0u, fp) == 0u) return true;
        return false;
    }
    for (uint32_t j = 0; j < c.bucket_size; ++j)
        if (__ldcg(s + j) == 0u && atomicCAS(s + j, 0u, fp) == 0u) return true;
    return false;
}

// bucket_place for 4-slot buckets from an already loaded snapshot of the bucket
__device__ __forceinline__ bool bucket_place_from(const CuckooDev &c, uint64_t b, uint32_t fp, const uint4 &v) {
    uint32_t *s = c.slots + b * 4;
    if (v.x == 0 && atomicCAS(s + 0, 0u, fp) == 0u) return true;
    if (v.y == 0 && atomicCAS(s + 1, 0u, fp) == 0u) return true;
    if (v.z == 0 && atomicCAS(s + 2, 0u, fp) == 0u) return true;
    if (v.w == 0 && atomicCAS(s + 3, 0u, fp) == 0u) return true;
    return false;
}

// Placement that doubles as the in-batch dedupe of the first kernel.  While that kernel runs, slots only ever go from
// empty to occupied, and every thread walks the empty slots of idx_1, then idx_2, in the same order with a CAS each:
// two threads that carry the same fingerprint meet on the first slot either of them wins -- the loser's CAS hands it
// its own fingerprint back.  kPlaced / kTwin (the same fingerprint got there first: nothing to do) / kNoRoom.
enum : int { kNoRoom = 0, kPlaced = 1, kTwin = 2 };
__device__ __forceinline__ int slot_try(uint32_t *s, uint32_t fp) {
    const uint32_t old = atomicCAS(s, 0u, fp);
    return old == 0u ? kPlaced : (old == fp ? kTwin : kNoRoom);
}
__device__ __forceinline__ int bucket_place_dedupe(const CuckooDev &c, uint64_t b, uint32_t fp, const uint4 &v) {
    uint32_t *s = c.slots + b * 4;
    int r;
    if (v.x == 0 && (r = slot_try(s + 0, fp)) != kNoRoom) return r;
    if (v.y == 0 && (r = slot_try(s + 1, fp)) != kNoRoom) return r;
    if (v.z == 0 && (r = slot_try(s + 2, fp)) != kNoRoom) return r;
    if (v.w == 0 && (r = slot_try(s + 3, fp)) != kNoRoom) return r;
    return kNoRoom;
}
template <int BS>
__device__ __forceinline__ int bucket_place_dedupe_any(const CuckooDev &c, uint64_t b, uint32_t fp) {
    uint32_t *s = c.slots + b * c.bucket_size;
    for (uint32_t j = 0; j < c.bucket_size; ++j) {
        const uint32_t cur = __ldcg(s + j);
        if (cur == fp) return kTwin;
        if (cur == 0u) {
            const int r = slot_try(s + j, fp);
            if (r != kNoRoom) return r;
        }
    }
    return kNoRoom;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) { return sm64(x); }

// One eviction step of cuckoo.py:375-388 for the fingerprint in hand: pick the victim in bucket idx, swap, try the
// victim's other bucket.  true = everything has a home now.
template <int BS>
__device__ __forceinline__ bool cuckoo_walk_step(const CuckooDev &c, uint32_t &fp, uint64_t &idx, uint64_t &rng) {
    rng = rng * 6364136223846793005ULL + 1442695040888963407ULL;
    uint32_t slot = (uint32_t)(((rng >> 33) * (uint64_t)c.bucket_size) >> 31);  // :377 (uniform in [0,bs))
    if (BS == 4) {
        // Informed choice of the victim (the reference draws it at random, :377; any choice leaves the same set of
        // stored fingerprints): look at the other bucket of the residents and evict one that has room there, so
        // that the walk ends with this step.
        const uint4 cur = __ldcg(reinterpret_cast<const uint4 *>(c.slots) + idx);
        const uint32_t res[4] = {cur.x, cur.y, cur.z, cur.w};
        const uint32_t start = slot;
        // one resident after the other, from a random start: the kernel is bound by random sectors, not by latency
        // (other lanes and warps fill the gaps), and stopping at the first resident with room costs ~2.7 look-ahead
        // sectors per step at 93 % load instead of 4
        for (int j = 0; j < 4; ++j) {
            const uint32_t q = (start + j) & 3u;
            const uint32_t r = q == 0 ? res[0] : q == 1 ? res[1] : q == 2 ? res[2] : res[3];
            if (r == 0u) {
                slot = q;
                break;
            }
            uint64_t a, b;
            cuckoo_buckets(c, r, a, b);
            const uint4 alt = __ldcg(reinterpret_cast<const uint4 *>(c.slots) + ((idx == a) ? b : a));
            if (alt.x == 0u || alt.y == 0u || alt.z == 0u || alt.w == 0u) {
                slot = q;
                break;
            }
        }
    }
    const uint32_t victim = atomicExch(c.slots + idx * c.bucket_size + slot, fp);  // :379-380
    if (victim == 0u) return true;  // the slot was (still) empty: nobody was evicted
    fp = victim;
    uint64_t a, b;
    cuckoo_buckets(c, fp, a, b);  // :383
    idx = (idx == a) ? b : a;     // :385
    return bucket_place<BS>(c, idx, fp);  // :387-388
}

// cuckoo.py:361-392 for one fingerprint that is known to be absent.  Returns true when everything found
// a home; false with `fp` = the homeless fingerprint.
template <int BS>
__device__ __forceinline__ bool cuckoo_insert_one(const CuckooDev &c, uint32_t &fp, uint64_t rng) {
    uint64_t i1, i2;
    cuckoo_buckets(c, fp, i1, i2);
    if (BS == 4) {
        // both candidate buckets are fetched before either is examined: one DRAM round trip instead of two when
        // idx_1 is full (the insert kernel is bound by exactly this latency chain at high load)
        const uint4 v1 = __ldcg(reinterpret_cast<const uint4 *>(c.slots) + i1);
        const uint4 v2 = __ldcg(reinterpret_cast<const uint4 *>(c.slots) + i2);
        if (bucket_place_from(c, i1, fp, v1)) return true;  // :363-365
        if (bucket_place_from(c, i2, fp, v2)) return true;  // :366-368
    } else {
        if (bucket_place<BS>(c, i1, fp)) return true;  // :363-365
        if (bucket_place<BS>(c, i2, fp)) return true;  // :366-368
    }
    rng = mix64(rng ^ fp);
    uint64_t idx = (rng & 1ull) ? i2 : i1;  // :373
    for (uint32_t s = 0; s < c.max_swaps; ++s)
        if (cuckoo_walk_step<BS>(c, fp, idx, rng)) return true;
    return false;  // :392
}

// ---- kernel 1: key -> fp, in-batch claim, presence filter, compaction ------------------------------
template <int BS>
__device__ __forceinline__ void claim_filter(const CuckooDev &c, uint32_t fp, bool active, uint32_t *__restrict__ newlist,
                                             CuckooCounters *cnt) {
    bool is_new = false;   // still homeless after the first try: goes to the eviction kernel
    uint32_t placed = 0;
    if (active) {
        if (fp == 0u) {
            // the zero fingerprint lives in the flag word; the flag itself is the claim
            if (atomicExch(c.zero_flag, 1u) == 0u) placed = 1;
        } else {
            uint64_t i1, i2;
            cuckoo_buckets(c, fp, i1, i2);
            int r;
            bool present;
            if (BS == 4) {
                // both buckets in flight together; the snapshots serve the presence test (:300-302) AND the first
                // placement attempt (:363-368): the CAS hits the sector the load has just brought into L2, and the
                // second kernel only ever sees fingerprints whose two buckets were full
                const uint4 v1 = __ldcg(reinterpret_cast<const uint4 *>(c.slots) + i1);
                const uint4 v2 = __ldcg(reinterpret_cast<const uint4 *>(c.slots) + i2);
                const bool in1 = v1.x == fp || v1.y == fp || v1.z == fp || v1.w == fp;
                const bool in2 = v2.x == fp || v2.y == fp || v2.z == fp || v2.w == fp;
                present = in1 | in2;
                r = kNoRoom;
                if (!present) {
                    r = bucket_place_dedupe(c, i1, fp, v1);
                    if (r == kNoRoom) r = bucket_place_dedupe(c, i2, fp, v2);
                }
            } else {
                present = bucket_has<BS>(c, i1, fp) || bucket_has<BS>(c, i2, fp);
                r = kNoRoom;
                if (!present) {
                    r = bucket_place_dedupe_any<BS>(c, i1, fp);
                    if (r == kNoRoom) r = bucket_place_dedupe_any<BS>(c, i2, fp);
                }
            }
            // Repeats of one fingerprint inside the batch sort themselves out on the slots (see bucket_place_dedupe); only
            // the fingerprints that found both buckets full -- none at low load -- need the claim set, so that exactly
            // one copy of each goes on to the eviction walk.  (r2: one claim atomic per KEY in a 512 MiB bitmap was a
            // quarter of this kernel's random DRAM accesses.)
            if (r == kPlaced) placed = 1;
            else if (!present && r == kNoRoom) is_new = claim_fp(c, fp);
        }
    }
    // warp-aggregated counters and append to the list of fingerprints that need the eviction walk
    const uint32_t pm = __ballot_sync(0xffffffffu, placed != 0);
    const uint32_t m = __ballot_sync(0xffffffffu, is_new);
    const uint32_t lane = threadIdx.x & 31u;
    if (pm && lane == (uint32_t)(__ffs(pm) - 1)) atomicAdd(&cnt->n_placed, (unsigned long long)__popc(pm));
    if (m) {
        unsigned long long base = 0;
        if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(&cnt->n_new, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (is_new) newlist[base + __popc(m & ((1u << lane) - 1u))] = fp;
    }
}

template <int BS>
__global__ void __launch_bounds__(256) cuckoo_claim_fixed16(const uint4 *__restrict__ keys, uint64_t n, CuckooDev c,
                                                            uint32_t *__restrict__ fps, uint32_t *__restrict__ newlist,
                                                            CuckooCounters *cnt) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + stride - 1) / stride;
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (uint64_t r = 0; r < rounds; ++r, i += stride) {
        const bool active = i < n;
        uint32_t fp = 0;
        if (active) {
            uint64_t h[1];
            fnv_group_16<1>(__ldcs(keys + i), 0, h);
            fp = cuckoo_fingerprint(h[0], c.fp_bits);  // :499-500
            fps[i] = fp;
        }
        claim_filter<BS>(c, fp, active, newlist, cnt);
    }
}

template <int BS>
__global__ void __launch_bounds__(256) cuckoo_claim_fps(const uint32_t *__restrict__ fps, uint64_t n, CuckooDev c,
                                                        uint32_t *__restrict__ newlist, CuckooCounters *cnt) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (n + stride - 1) / stride;
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (uint64_t r = 0; r < rounds; ++r, i += stride) {
        const bool active = i < n;
        const uint32_t fp = active ? cuckoo_fingerprint((uint64_t)fps[i], c.fp_bits) : 0u;
        claim_filter<BS>(c, fp, active, newlist, cnt);
    }
}

// generic keys: fingerprints only (the claim then runs from the fp array)
template <int SYMW>
__global__ void __launch_bounds__(kTileKeys) cuckoo_fp_staged(DevKeys dk, uint32_t fp_bits, uint32_t *__restrict__ fps) {
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
            uint64_t h[1];
            fnv_group_ptr<1, SYMW>(kr.p, kr.len, 0, h);
            fps[first + threadIdx.x] = cuckoo_fingerprint(h[0], fp_bits);
        }
    }
}

__global__ void __launch_bounds__(256) cuckoo_fp_fixed16(const uint4 *__restrict__ keys, uint64_t n, uint32_t fp_bits,
                                                         uint32_t *__restrict__ fps) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h[1];
        fnv_group_16<1>(__ldcs(keys + i), 0, h);
        fps[i] = cuckoo_fingerprint(h[0], fp_bits);
    }
}

// ---- kernel 2: insert distinct, absent fingerprints ---------------------------------------------------
// A work queue instead of one list entry per thread: walks differ in length (most end after one step, some take tens),
// and with a fixed assignment a warp runs for as long as its longest walk with most lanes idle (r2 ncu at 93 % load:
// 4.3 G warp instructions for 37 M walks, issue slots 47 % busy -- this kernel was instruction bound).  Every lane that
// finishes takes the next entry, so a warp's instruction count follows the AVERAGE walk.
// skip_zero: the list is an old slot array (expand, :467-481) whose zeros are empty slots.
template <int BS>
__global__ void __launch_bounds__(256) cuckoo_insert_kernel(const uint32_t *__restrict__ list, const unsigned long long *n_ptr,
                                                            uint64_t n_fixed, int skip_zero, CuckooDev c, uint64_t rng_seed,
                                                            uint32_t *__restrict__ failed, uint64_t failed_cap, CuckooCounters *cnt) {
    const uint64_t n = n_ptr ? (uint64_t)*n_ptr : n_fixed;
    const uint32_t warp = __activemask();  // (the in-order variant launches a single thread)
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long placed = 0;
    uint32_t fp = 0, steps = 0;
    uint64_t idx = 0, rng = 0;
    bool busy = false, fresh = false, drained = false;
    uint64_t q_next = 0, q_end = 0;  // this warp's current block of list entries (warp-uniform)
    constexpr uint64_t kBlock = 128;  // entries per trip to the global cursor
    for (;;) {
        const uint32_t idle = __ballot_sync(warp, !busy);
        if (idle && !drained) {
            if (q_next == q_end) {
                unsigned long long base = 0;
                const int leader = __ffs(warp) - 1;
                if ((int)lane == leader) base = atomicAdd(&cnt->cursor, (unsigned long long)kBlock);
                base = __shfl_sync(warp, base, leader);
                q_next = base < n ? base : n;
                q_end = base + kBlock < n ? base + kBlock : n;
            }
            const uint64_t avail = q_end - q_next;
            const uint32_t rank = __popc(idle & ((1u << lane) - 1u));
            if (!busy && rank < avail) {
                const uint64_t i = q_next + rank;
                fp = list[i];
                if (!(skip_zero && fp == 0u)) {
                    busy = fresh = true;
                    rng = rng_seed + i * 0x9E3779B97F4A7C15ULL;
                }
            }
            const uint64_t want = (uint64_t)__popc(idle);
            q_next += want < avail ? want : avail;
            drained = q_next >= n;
        }
        if (!__any_sync(warp, busy)) {
            if (drained) break;
            continue;  // (only zeros of an old slot array were drawn: draw again)
        }
        if (busy) {
            bool done = false;
            if (fresh) {
                fresh = false;
                uint64_t i1, i2;
                cuckoo_buckets(c, fp, i1, i2);
                // entries of an add batch found both buckets full in the first kernel, and slots never empty again
                // while a batch is being added: only an old slot array (expand) has to try the plain placement here
                if (skip_zero) {
                    if (BS == 4) {
                        const uint4 v1 = __ldcg(reinterpret_cast<const uint4 *>(c.slots) + i1);
                        const uint4 v2 = __ldcg(reinterpret_cast<const uint4 *>(c.slots) + i2);
                        done = bucket_place_from(c, i1, fp, v1) || bucket_place_from(c, i2, fp, v2);  // :363-368
                    } else {
                        done = bucket_place<BS>(c, i1, fp) || bucket_place<BS>(c, i2, fp);
                    }
                }
                rng = mix64(rng ^ fp);
                idx = (rng & 1ull) ? i2 : i1;  // :373
                steps = 0;
            }
            if (done) {
                // placed without a walk
            } else if (steps < c.max_swaps) {
                ++steps;
                done = cuckoo_walk_step<BS>(c, fp, idx, rng);
            } else {  // :392 -- out of swaps: the fingerprint in hand is homeless
                const unsigned long long pos = atomicAdd(&cnt->n_failed, 1ull);
                if (pos < failed_cap) failed[pos] = fp;
                busy = false;
            }
            if (done) {
                ++placed;
                busy = false;
            }
        }
    }
    if (placed) atomicAdd(&cnt->n_placed, placed);  // one per thread: noise next to the table traffic
}

// ---- serial restatement on the device (one thread, key order) -----------------------------------------
template <int BS>
__global__ void cuckoo_serial_kernel(const uint32_t *__restrict__ fps, uint64_t n, CuckooDev c, uint64_t rng_seed,
                                     uint32_t *__restrict__ failed, uint64_t failed_cap, CuckooCounters *cnt) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t fp = cuckoo_fingerprint((uint64_t)fps[i], c.fp_bits);
        if (fp == 0u) {
            if (atomicExch(c.zero_flag, 1u) == 0u) cnt->n_placed++;
            continue;
        }
        uint64_t i1, i2;
        cuckoo_buckets(c, fp, i1, i2);
        if (bucket_has<BS>(c, i1, fp) || bucket_has<BS>(c, i2, fp)) continue;
        cnt->n_new++;
        if (cuckoo_insert_one<BS>(c, fp, rng_seed + i * 0x9E3779B97F4A7C15ULL)) {
            cnt->n_placed++;
        } else {
            if (cnt->n_failed < failed_cap) failed[cnt->n_failed] = fp;
            cnt->n_failed++;
        }
        __threadfence();
    }
}

// ---- check (cuckoo.py:306-315) ---------------------------------------------------------------------------
template <int BS>
__device__ __forceinline__ uint8_t cuckoo_lookup(const CuckooDev &c, uint32_t fp) {
    if (fp == 0u) return (uint8_t)(__ldcg(c.zero_flag) != 0u);
    uint64_t i1, i2;
    cuckoo_buckets(c, fp, i1, i2);
    // both bucket reads are issued before either is tested
    const bool a = bucket_has<BS>(c, i1, fp);
    const bool b = bucket_has<BS>(c, i2, fp);
    return (uint8_t)(a | b);
}

template <int BS>
__global__ void __launch_bounds__(256)
    cuckoo_check_fixed16(const uint4 *__restrict__ keys, uint64_t n, CuckooDev c, uint8_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h[1];
        fnv_group_16<1>(__ldcs(keys + i), 0, h);
        out[i] = cuckoo_lookup<BS>(c, cuckoo_fingerprint(h[0], c.fp_bits));
    }
}

template <int BS>
__global__ void __launch_bounds__(256)
    cuckoo_check_fps(const uint32_t *__restrict__ fps, uint64_t n, CuckooDev c, uint8_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = cuckoo_lookup<BS>(c, cuckoo_fingerprint((uint64_t)fps[i], c.fp_bits));
}

// ---- remove (cuckoo.py:317-330): clear the slot holding fp in idx_1, else in idx_2 -----------------------------
template <int BS>
__device__ __forceinline__ bool bucket_take(const CuckooDev &c, uint64_t b, uint32_t fp) {
    uint32_t *s = c.slots + b * c.bucket_size;
    for (uint32_t j = 0; j < c.bucket_size; ++j)
        if (__ldcg(s + j) == fp && atomicCAS(s + j, fp, 0u) == fp) return true;  // one winner per stored copy
    return false;
}

// i2_in == nullptr: idx_2 from the built-in FNV of the decimal digits; else the caller's (custom hash_function)
template <int BS>
__global__ void __launch_bounds__(256) cuckoo_remove_fps(const uint32_t *__restrict__ fps, const uint64_t *__restrict__ i2_in, uint64_t n,
                                                         CuckooDev c, uint8_t *__restrict__ out, CuckooCounters *cnt) {
    unsigned long long removed = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t fp = cuckoo_fingerprint((uint64_t)fps[i], c.fp_bits);
        bool hit;
        if (fp == 0u) {
            hit = atomicExch(c.zero_flag, 0u) != 0u;
        } else {
            uint64_t i1, i2;
            cuckoo_buckets(c, fp, i1, i2);
            if (i2_in) i2 = i2_in[i];
            hit = bucket_take<BS>(c, i1, fp) || bucket_take<BS>(c, i2, fp);  // :325-328 (idx_1 first)
        }
        out[i] = (uint8_t)hit;
        removed += hit;
    }
    if (removed) atomicAdd(&cnt->n_placed, removed);
}

// ---- pre-indexed filters (a custom hash_function: idx_2 = hash_function(str(fp)) % capacity, cuckoo.py:489, cannot
// be evaluated on the device).  The caller supplies idx_2 with every fingerprint and the table keeps the idx_2 of
// every stored fingerprint next to it (`alt`), which is what the eviction walk needs for its victims (:383-385).
// One thread, key order: these filters are bound by the user's Python hash anyway, and the order makes the
// result the reference's exactly (append order, :448-453).
template <int BS>
__device__ __forceinline__ bool indexed_place(const CuckooDev &c, uint64_t b, uint32_t fp, uint64_t i2) {
    uint32_t *s = c.slots + b * c.bucket_size;
    for (uint32_t j = 0; j < c.bucket_size; ++j) {
        if (s[j] == 0u) {
            s[j] = fp;
            c.alt[b * c.bucket_size + j] = i2;
            return true;
        }
    }
    return false;
}

template <int BS>
__global__ void cuckoo_serial_indexed(const uint32_t *__restrict__ fps, const uint64_t *__restrict__ i2_in, uint64_t n, CuckooDev c,
                                      uint64_t rng, uint32_t *__restrict__ failed_fp, uint64_t *__restrict__ failed_i2,
                                      uint64_t failed_cap, CuckooCounters *cnt) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t fp = cuckoo_fingerprint((uint64_t)fps[i], c.fp_bits);
        if (fp == 0u) {
            if (*c.zero_flag == 0u) {
                *c.zero_flag = 1u;
                cnt->n_placed++;
            }
            continue;
        }
        uint64_t b2 = i2_in[i];
        const uint64_t b1 = fastmod((uint64_t)fp, c.fm);
        if (bucket_has<BS>(c, b1, fp) || bucket_has<BS>(c, b2, fp)) continue;  // :300-302
        cnt->n_new++;
        bool done = indexed_place<BS>(c, b1, fp, b2) || indexed_place<BS>(c, b2, fp, b2);  // :363-368
        if (!done) {
            rng = sm64(rng ^ fp);
            uint64_t idx = (rng & 1ull) ? b2 : b1;  // :373
            for (uint32_t s = 0; s < c.max_swaps && !done; ++s) {
                rng = rng * 6364136223846793005ULL + 1442695040888963407ULL;
                const uint64_t pos = idx * c.bucket_size + (uint32_t)(((rng >> 33) * (uint64_t)c.bucket_size) >> 31);  // :377
                const uint32_t victim = c.slots[pos];
                const uint64_t victim_i2 = c.alt[pos];
                c.slots[pos] = fp;  // :379-380
                c.alt[pos] = b2;
                fp = victim;
                b2 = victim_i2;
                const uint64_t a = fastmod((uint64_t)fp, c.fm);  // :383
                idx = (idx == a) ? b2 : a;                       // :385
                done = indexed_place<BS>(c, idx, fp, b2);        // :387-388
            }
        }
        if (done) {
            cnt->n_placed++;
        } else {
            if (cnt->n_failed < failed_cap) {
                failed_fp[cnt->n_failed] = fp;
                failed_i2[cnt->n_failed] = b2;
            }
            cnt->n_failed++;
        }
        __threadfence();
    }
}

template <int BS>
__global__ void __launch_bounds__(256) cuckoo_check_indexed(const uint32_t *__restrict__ fps, const uint64_t *__restrict__ i2_in, uint64_t n,
                                                            CuckooDev c, uint8_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t fp = cuckoo_fingerprint((uint64_t)fps[i], c.fp_bits);
        if (fp == 0u) {
            out[i] = (uint8_t)(__ldcg(c.zero_flag) != 0u);
        } else {
            const bool a = bucket_has<BS>(c, fastmod((uint64_t)fp, c.fm), fp);
            const bool b = bucket_has<BS>(c, i2_in[i], fp);
            out[i] = (uint8_t)(a | b);
        }
    }
}

// _generate_fingerprint_info (:492-506) from fingerprints
__global__ void __launch_bounds__(256) cuckoo_info_kernel(const uint32_t *__restrict__ fps, uint64_t n, CuckooDev c,
                                                          uint64_t *__restrict__ i1, uint64_t *__restrict__ i2) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        cuckoo_buckets(c, fps[i], i1[i], i2[i]);
}

__global__ void __launch_bounds__(256) count_nonzero_kernel(const uint32_t *__restrict__ s, uint64_t n, unsigned long long *out) {
    unsigned long long c = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        c += (s[i] != 0u);
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// ---- CountingCuckooFilter (cuckoo/countingcuckoo.py:156-210) -------------------------------------------------------
// The reference keeps (fingerprint, count) bins; add() of a fingerprint that is already stored bumps its count (:165-171),
// otherwise the fingerprint goes in with count 1 (:172, :230-265) and the count travels with it through evictions.  A
// fingerprint is therefore stored at most once, and the counts are a function fingerprint -> count that no eviction
// touches: they live in their own open-addressing map next to the (unchanged) fingerprint table.  add = the set insert
// above + one atomicAdd per key here; check = one map lookup; remove = decrement, and take the fingerprint out of the
// table when its count reaches 0 (:193-210).
struct FpCounts {
    uint32_t *keys, *vals, *tickets;
    uint64_t mask;  // cap - 1; index cap = fingerprint 0
};

// `claimed` counts the entries this thread took for a fingerprint the map had not seen (callers add the per-thread
// totals up once per warp: one global atomic per new fingerprint would serialise a batch of new keys on one address)
__device__ __forceinline__ uint64_t counts_claim(const FpCounts &m, uint32_t fp, unsigned long long &claimed) {
    if (fp == 0u) return m.mask + 1;
    uint64_t s = sm64((uint64_t)fp) & m.mask;
    for (;;) {
        const uint32_t old = atomicCAS(m.keys + s, 0u, fp);
        if (old == 0u) {
            ++claimed;
            return s;
        }
        if (old == fp) return s;
        s = (s + 1) & m.mask;
    }
}
__device__ __forceinline__ void counts_flush_claimed(unsigned long long claimed, unsigned long long *used) {
    for (int o = 16; o; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
    if ((threadIdx.x & 31) == 0 && claimed) atomicAdd(used, claimed);
}
// index of fp's entry, or ~0 when the map never saw it
__device__ __forceinline__ uint64_t counts_find(const FpCounts &m, uint32_t fp) {
    if (fp == 0u) return m.mask + 1;
    uint64_t s = sm64((uint64_t)fp) & m.mask;
    for (;;) {
        const uint32_t k = __ldcg(m.keys + s);
        if (k == fp) return s;
        if (k == 0u) return ~0ull;
        s = (s + 1) & m.mask;
    }
}

// (fps may hold 32-bit hashes still to be cut to fp_bits, as in the set kernels above: the cut is idempotent)
__global__ void __launch_bounds__(256) counts_add_kernel(const uint32_t *__restrict__ fps, const uint32_t *__restrict__ amounts, uint64_t n,
                                                         uint32_t fp_bits, FpCounts m, unsigned long long *used) {
    unsigned long long claimed = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t fp = cuckoo_fingerprint((uint64_t)fps[i], fp_bits);
        atomicAdd(m.vals + counts_claim(m, fp, claimed), amounts ? amounts[i] : 1u);  // :165-171 / :230-241
    }
    counts_flush_claimed(claimed, used);
}
__global__ void __launch_bounds__(256) counts_get_kernel(const uint32_t *__restrict__ fps, uint64_t n, uint32_t fp_bits, FpCounts m,
                                                         uint32_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = counts_find(m, cuckoo_fingerprint((uint64_t)fps[i], fp_bits));
        out[i] = s == ~0ull ? 0u : __ldcg(m.vals + s);  // :175-191
    }
}
__global__ void __launch_bounds__(256) counts_set_kernel(const uint32_t *__restrict__ fps, const uint32_t *__restrict__ vals, uint64_t n,
                                                         FpCounts m, unsigned long long *used) {
    unsigned long long claimed = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        m.vals[counts_claim(m, fps[i], claimed)] = vals ? vals[i] : 0u;
    counts_flush_claimed(claimed, used);
}
// remove, pass 1: every occurrence of a stored fingerprint draws a ticket; tickets below the count are the removals
// that succeed (:199-208: a key removed more often than it was added runs dry).  The counts do not move in this pass.
__global__ void __launch_bounds__(256) counts_remove_tickets(const uint32_t *__restrict__ fps, uint64_t n, uint32_t fp_bits, FpCounts m,
                                                             uint32_t *__restrict__ ticket_of, uint8_t *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = counts_find(m, cuckoo_fingerprint((uint64_t)fps[i], fp_bits));
        uint32_t t = 0xFFFFFFFFu, have = 0;
        if (s != ~0ull && (have = __ldcg(m.vals + s)) != 0u) t = atomicAdd(m.tickets + s, 1u);
        ticket_of[i] = t;
        out[i] = (uint8_t)(t < have);
    }
}
// pass 2: the holder of ticket 0 settles its fingerprint: count -= min(tickets, count); a count of 0 takes the
// fingerprint out of the table (:205-207)
template <int BS>
__global__ void __launch_bounds__(256) counts_remove_settle(const uint32_t *__restrict__ fps, const uint64_t *__restrict__ i2_in, uint64_t n,
                                                            CuckooDev c, FpCounts m, const uint32_t *__restrict__ ticket_of,
                                                            unsigned long long *totals /* [0] removed, [1] bins removed */) {
    unsigned long long removed = 0, bins = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (ticket_of[i] != 0u) continue;
        const uint32_t fp = cuckoo_fingerprint((uint64_t)fps[i], c.fp_bits);
        const uint64_t s = counts_find(m, fp);
        const uint32_t drawn = m.tickets[s], have = m.vals[s];
        const uint32_t dec = drawn < have ? drawn : have;
        m.vals[s] = have - dec;
        m.tickets[s] = 0u;
        removed += dec;
        if (have == dec) {
            bool hit;
            if (fp == 0u) {
                hit = atomicExch(c.zero_flag, 0u) != 0u;
            } else {
                uint64_t i1, i2;
                cuckoo_buckets(c, fp, i1, i2);
                if (i2_in) i2 = i2_in[i];
                hit = bucket_take<BS>(c, i1, fp) || bucket_take<BS>(c, i2, fp);
            }
            bins += hit;
        }
    }
    counts_flush_claimed(removed, totals);
    counts_flush_claimed(bins, totals + 1);
}
__global__ void __launch_bounds__(256) counts_rehash_kernel(FpCounts from, FpCounts to, unsigned long long *used) {
    const uint64_t n = from.mask + 2;  // incl. the entry of fingerprint 0
    unsigned long long claimed = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t v = from.vals[i];
        if (v == 0u) continue;
        if (i == from.mask + 1) to.vals[to.mask + 1] = v;
        else to.vals[counts_claim(to, from.keys[i], claimed)] = v;
    }
    counts_flush_claimed(claimed, used);
}

// ---------------------------------------------------------------- host side
static CuckooDev dev_view(const pb_cuckoo *c) {
    CuckooDev d;
    d.slots = c->slots;
    d.zero_flag = c->zero_flag;
    d.claim = nullptr;  // chosen per call (add_fps_device)
    d.claim_mask = 0;
    d.alt = c->alt;
    d.fm = c->fm;
    d.bucket_size = c->bucket_size;
    d.max_swaps = c->max_swaps;
    d.fp_bits = c->fp_bits;
    return d;
}

#define PB_BS_DISPATCH(c, KERNEL, GRID, BLOCK, ...)                                               \
    do {                                                                                          \
        launch_begin((c)->ctx);                                                                   \
        if ((c)->bucket_size == 4) KERNEL<4><<<GRID, BLOCK, 0, (c)->ctx->stream>>>(__VA_ARGS__);    \
        else KERNEL<0><<<GRID, BLOCK, 0, (c)->ctx->stream>>>(__VA_ARGS__);                          \
    } while (0)

// The 2^fp_bits-bit claim bitmap lives in the CONTEXT's scratch (one per context, not one per filter) and is all
// zero between calls: a (re)allocation zeroes it, every user clears what it touched.
static int ensure_claim(pb_cuckoo *c, uint32_t **bitmap, uint64_t *bytes) {
    pb_ctx *ctx = c->ctx;
    *bytes = c->fp_bits >= 5 ? (1ull << (c->fp_bits - 3)) : 4ull;
    const void *before = ctx->claim_bitmap.p;
    PB_TRY(scratch_reserve(ctx, ctx->claim_bitmap, *bytes));
    if (ctx->claim_bitmap.p != before) PB_CUDA(cudaMemsetAsync(ctx->claim_bitmap.p, 0, ctx->claim_bitmap.cap, ctx->stream));
    *bitmap = (uint32_t *)ctx->claim_bitmap.p;
    return PB_OK;
}

// scratch layout inside ctx->small: [0] popcount, [8] stray, [16..17] cms sums, [32..35] CuckooCounters, [44..46] count map
static CuckooCounters *counters_dev(pb_ctx *ctx) { return (CuckooCounters *)((unsigned long long *)ctx->small.p + 32); }

struct CuckooResult {
    uint64_t n_added = 0, n_failed = 0;
    uint32_t *failed_out = nullptr;  // host
    uint64_t failed_cap = 0, failed_written = 0;
};

// Inserts the fingerprints fps_dev[0..n) (device) with in-batch dedupe and presence filtering.
// `fused_keys` != nullptr: kernel 1 hashes the fixed-16 keys itself and fills fps_dev.
static int add_fps_device(pb_cuckoo *c, const uint4 *fused_keys, uint32_t *fps_dev, uint64_t n, int slot, CuckooResult *res) {
    pb_ctx *ctx = c->ctx;
    if (n == 0) return PB_OK;
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    CuckooCounters *cnt = counters_dev(ctx);
    PB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(CuckooCounters), ctx->stream));
    // newlist and failed share one allocation: [newlist n][failed n]
    PB_TRY(scratch_reserve(ctx, ctx->aux_stage[slot], n * 8));
    uint32_t *newlist = (uint32_t *)ctx->aux_stage[slot].p;
    uint32_t *failed = newlist + n;
    const CuckooDev cd = dev_view(c);
    const uint64_t seed = c->rng_seed + (++c->epoch) * 0xD1B54A32D192ED03ULL;
    if (ctx->cuckoo_serial) {
        if (fused_keys) {
            cuckoo_fp_fixed16<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(fused_keys, n, c->fp_bits, fps_dev);
            PB_TRY(check_launch(ctx, "cuckoo_fp"));
        }
        PB_BS_DISPATCH(c, cuckoo_serial_kernel, 1, 32, fps_dev, n, cd, seed, failed, n, cnt);
        PB_TRY(check_launch(ctx, "cuckoo_serial"));
    } else {
        // in-batch dedupe scratch: an exact open-addressing set of >= 4n entries in the context's scratch for batches
        // whose set is smaller than the 2^fp_bits-bit bitmap (512 MiB for 32-bit fingerprints), the bitmap otherwise
        CuckooDev cd2 = dev_view(c);
        const uint64_t bitmap_bytes = c->fp_bits >= 5 ? (1ull << (c->fp_bits - 3)) : 4ull;
        uint64_t bitmap_bytes_used = 0;
        uint64_t set_entries = 64;
        while (set_entries < 4 * n) set_entries <<= 1;
        const bool use_set = set_entries * 4 < bitmap_bytes;
        if (use_set) {
            PB_TRY(scratch_reserve(ctx, ctx->claim_set, set_entries * 4));
            PB_CUDA(cudaMemsetAsync(ctx->claim_set.p, 0, set_entries * 4, ctx->stream));
            cd2.claim = (uint32_t *)ctx->claim_set.p;
            cd2.claim_mask = (uint32_t)(set_entries - 1);
        } else {
            PB_TRY(ensure_claim(c, &cd2.claim, &bitmap_bytes_used));
        }
        const int grid = grid_for(ctx, n, 256, 8);
        if (fused_keys) PB_BS_DISPATCH(c, cuckoo_claim_fixed16, grid, 256, fused_keys, n, cd2, fps_dev, newlist, cnt);
        else PB_BS_DISPATCH(c, cuckoo_claim_fps, grid, 256, fps_dev, n, cd2, newlist, cnt);
        PB_TRY(check_launch(ctx, "cuckoo_claim"));
        PB_BS_DISPATCH(c, cuckoo_insert_kernel, grid, 256, newlist, &cnt->n_new, 0, 0, cd2, seed, failed, n, cnt);
        PB_TRY(check_launch(ctx, "cuckoo_insert"));
        if (!use_set) PB_CUDA(cudaMemsetAsync(cd2.claim, 0, bitmap_bytes_used, ctx->stream));
    }
    // results (small D2H; this call is synchronous by contract because failures must be reported)
    CuckooCounters *h = (CuckooCounters *)((uint8_t *)ctx->pinned_small + 256);
    PB_CUDA(cudaMemcpyAsync(h, cnt, sizeof(CuckooCounters), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    res->n_added += h->n_placed;
    if (h->n_failed) {
        const uint64_t have = std::min<uint64_t>(h->n_failed, n);
        const uint64_t room = res->failed_cap > res->failed_written ? res->failed_cap - res->failed_written : 0;
        const uint64_t take = std::min(have, room);
        if (take && res->failed_out) {
            PB_CUDA(cudaMemcpyAsync(res->failed_out + res->failed_written, failed, take * 4, cudaMemcpyDeviceToHost, ctx->stream));
            PB_CUDA(cudaStreamSynchronize(ctx->stream));
            res->failed_written += take;
        }
        res->n_failed += h->n_failed;
    }
    return PB_OK;
}

struct CuckooAddArgs {
    pb_cuckoo *c;
    CuckooResult *res;
};

static int cuckoo_add_chunk(pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) {
    (void)first;
    CuckooAddArgs *a = (CuckooAddArgs *)user;
    pb_cuckoo *c = a->c;
    PB_TRY(scratch_reserve(ctx, ctx->out_stage[slot], dk.n * 4));
    uint32_t *fps = (uint32_t *)ctx->out_stage[slot].p;
    if (is_fixed16(dk)) return add_fps_device(c, (const uint4 *)dk.data, fps, dk.n, slot, a->res);
    const uint64_t tiles = (dk.n + kTileKeys - 1) / kTileKeys;
    const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->num_sms * 4);
    if (dk.sym_width == 4) cuckoo_fp_staged<4><<<grid, kTileKeys, 0, ctx->stream>>>(dk, c->fp_bits, fps);
    else cuckoo_fp_staged<1><<<grid, kTileKeys, 0, ctx->stream>>>(dk, c->fp_bits, fps);
    PB_TRY(check_launch(ctx, "cuckoo_fp_staged"));
    return add_fps_device(c, nullptr, fps, dk.n, slot, a->res);
}

struct CuckooCheckArgs {
    pb_cuckoo *c;
    uint8_t *out_dev;
    uint8_t *out_host;
};

static int cuckoo_check_chunk(pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) {
    CuckooCheckArgs *a = (CuckooCheckArgs *)user;
    pb_cuckoo *c = a->c;
    const CuckooDev cd = dev_view(c);
    uint8_t *out = a->out_dev ? a->out_dev + first : nullptr;
    if (!out) {
        PB_TRY(scratch_reserve(ctx, ctx->out_stage[slot], dk.n));
        out = (uint8_t *)ctx->out_stage[slot].p;
    }
    if (is_fixed16(dk)) {
        PB_BS_DISPATCH(c, cuckoo_check_fixed16, grid_for(ctx, dk.n, 256, 8), 256, (const uint4 *)dk.data, dk.n, cd, out);
        PB_TRY(check_launch(ctx, "cuckoo_check"));
    } else {
        PB_TRY(scratch_reserve(ctx, ctx->aux_stage[slot], dk.n * 4));
        uint32_t *fps = (uint32_t *)ctx->aux_stage[slot].p;
        const uint64_t tiles = (dk.n + kTileKeys - 1) / kTileKeys;
        const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->num_sms * 4);
        if (dk.sym_width == 4) cuckoo_fp_staged<4><<<grid, kTileKeys, 0, ctx->stream>>>(dk, c->fp_bits, fps);
        else cuckoo_fp_staged<1><<<grid, kTileKeys, 0, ctx->stream>>>(dk, c->fp_bits, fps);
        PB_TRY(check_launch(ctx, "cuckoo_fp_staged"));
        PB_BS_DISPATCH(c, cuckoo_check_fps, grid_for(ctx, dk.n, 256, 8), 256, fps, dk.n, cd, out);
        PB_TRY(check_launch(ctx, "cuckoo_check_fps"));
    }
    if (a->out_host) PB_CUDA(cudaMemcpyAsync(a->out_host + first, out, dk.n, cudaMemcpyDeviceToHost, ctx->stream));
    return PB_OK;
}

struct CuckooInfoArgs {
    pb_cuckoo *c;
    uint32_t *fp;
    uint64_t *i1, *i2;
    int out_on_device;
};

static int cuckoo_info_chunk(pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) {
    CuckooInfoArgs *a = (CuckooInfoArgs *)user;
    pb_cuckoo *c = a->c;
    uint32_t *fps;
    uint64_t *i1, *i2;
    if (a->out_on_device) {
        fps = a->fp + first;
        i1 = a->i1 + first;
        i2 = a->i2 + first;
    } else {
        PB_TRY(scratch_reserve(ctx, ctx->out_stage[slot], dk.n * 20 + 64));
        i1 = (uint64_t *)ctx->out_stage[slot].p;
        i2 = i1 + dk.n;
        fps = (uint32_t *)(i2 + dk.n);
    }
    if (is_fixed16(dk)) {
        cuckoo_fp_fixed16<<<grid_for(ctx, dk.n, 256, 8), 256, 0, ctx->stream>>>((const uint4 *)dk.data, dk.n, c->fp_bits, fps);
    } else {
        const uint64_t tiles = (dk.n + kTileKeys - 1) / kTileKeys;
        const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->num_sms * 4);
        if (dk.sym_width == 4) cuckoo_fp_staged<4><<<grid, kTileKeys, 0, ctx->stream>>>(dk, c->fp_bits, fps);
        else cuckoo_fp_staged<1><<<grid, kTileKeys, 0, ctx->stream>>>(dk, c->fp_bits, fps);
    }
    PB_TRY(check_launch(ctx, "cuckoo_fp"));
    cuckoo_info_kernel<<<grid_for(ctx, dk.n, 256, 8), 256, 0, ctx->stream>>>(fps, dk.n, dev_view(c), i1, i2);
    PB_TRY(check_launch(ctx, "cuckoo_info"));
    if (!a->out_on_device) {
        PB_CUDA(cudaMemcpyAsync(a->fp + first, fps, dk.n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(cudaMemcpyAsync(a->i1 + first, i1, dk.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(cudaMemcpyAsync(a->i2 + first, i2, dk.n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    return PB_OK;
}

static int finish_add(int st, const CuckooResult &res, uint64_t *n_added, uint64_t *n_failed) {
    if (n_added) *n_added = res.n_added;
    if (n_failed) *n_failed = res.n_failed;
    PB_TRY(st);
    if (res.n_failed) {
        set_error("The CuckooFilter is currently full (%llu fingerprints left homeless)", (unsigned long long)res.n_failed);
        return PB_ERR_CUCKOO_FULL;
    }
    return PB_OK;
}

static int cuckoo_alloc_table(pb_ctx *ctx, uint64_t capacity, uint32_t bucket_size, uint32_t **slots, uint64_t *nslots) {
    const uint64_t n = capacity * (uint64_t)bucket_size;
    PB_REQUIRE(n / bucket_size == capacity && n < (1ull << 60), "capacity * bucket_size overflows");
    const uint64_t alloc = (n + 3) & ~(uint64_t)3;
    cudaError_t e = cudaMalloc(slots, alloc * 4);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc of %llu cuckoo table bytes failed: %s", (unsigned long long)(alloc * 4), cudaGetErrorString(e));
        return PB_ERR_OOM;
    }
    PB_CUDA(cudaMemsetAsync(*slots, 0, alloc * 4, ctx->stream));
    *nslots = n;
    return PB_OK;
}

}  // namespace pb

extern "C" {

int pb_cuckoo_create(pb_ctx *ctx, uint64_t capacity, uint32_t bucket_size, uint32_t max_swaps, uint32_t fp_bits, uint64_t rng_seed,
                     pb_cuckoo **out) {
    PB_REQUIRE(ctx && out, "NULL argument");
    PB_REQUIRE(capacity >= 1 && bucket_size >= 1 && max_swaps >= 1,
               "CuckooFilter: capacity, bucket_size, and max_swaps must be an integer greater than 0");
    PB_REQUIRE(fp_bits >= 1 && fp_bits <= 32, "fingerprint size must be 1..32 bits (got %u)", fp_bits);
    DeviceGuard g(ctx->device);
    pb_cuckoo *c = new (std::nothrow) pb_cuckoo();
    if (!c) return PB_ERR_OOM;
    c->ctx = ctx;
    c->capacity = capacity;
    c->bucket_size = bucket_size;
    c->max_swaps = max_swaps;
    c->fp_bits = fp_bits;
    c->rng_seed = rng_seed ? rng_seed : 0x9E3779B97F4A7C15ULL;
    c->fm = make_fastmod(capacity);
    int st = cuckoo_alloc_table(ctx, capacity, bucket_size, &c->slots, &c->nslots);
    if (st != PB_OK) {
        delete c;
        return st;
    }
    if (cudaMalloc(&c->zero_flag, 256) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(c->slots);
        delete c;
        set_error("cudaMalloc failed");
        return PB_ERR_OOM;
    }
    cudaMemsetAsync(c->zero_flag, 0, 256, ctx->stream);
    *out = c;
    return PB_OK;
}

int pb_cuckoo_destroy(pb_cuckoo *c) {
    if (!c) return PB_OK;
    DeviceGuard g(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    cudaFree(c->slots);
    cudaFree(c->zero_flag);
    if (c->alt) cudaFree(c->alt);
    cudaFree(c->cnt_keys);
    cudaFree(c->cnt_vals);
    cudaFree(c->cnt_tickets);
    delete c;
    return PB_OK;
}

int pb_cuckoo_clear(pb_cuckoo *c) {
    PB_REQUIRE(c, "handle is NULL");
    DeviceGuard g(c->ctx->device);
    PB_CUDA(cudaMemsetAsync(c->slots, 0, ((c->nslots + 3) & ~(uint64_t)3) * 4, c->ctx->stream));
    PB_CUDA(cudaMemsetAsync(c->zero_flag, 0, 4, c->ctx->stream));
    if (c->cnt_cap) {
        PB_CUDA(cudaMemsetAsync(c->cnt_keys, 0, (c->cnt_cap + 1) * 4, c->ctx->stream));
        PB_CUDA(cudaMemsetAsync(c->cnt_vals, 0, (c->cnt_cap + 1) * 4, c->ctx->stream));
        c->cnt_used = 0;
    }
    return PB_OK;
}

int pb_cuckoo_add_keys(pb_cuckoo *c, const pb_keys *keys, uint64_t *n_added, uint64_t *n_failed, uint32_t *failed_fps,
                       uint64_t failed_cap) {
    PB_REQUIRE(c && keys, "NULL argument");
    DeviceGuard g(c->ctx->device);
    CuckooResult res;
    res.failed_out = failed_fps;
    res.failed_cap = failed_fps ? failed_cap : 0;
    CuckooAddArgs a{c, &res};
    // chunk so the per-chunk scratch (12 B/key) stays bounded for device-resident batches
    const int st = for_each_chunk(c->ctx, keys, cuckoo_add_chunk, &a, 1ull << 28);
    return finish_add(st, res, n_added, n_failed);
}

int pb_cuckoo_add_fingerprints(pb_cuckoo *c, const uint32_t *fps, uint64_t n, int on_device, uint64_t *n_added, uint64_t *n_failed,
                               uint32_t *failed_fps, uint64_t failed_cap) {
    PB_REQUIRE(c && (fps || n == 0), "NULL argument");
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    CuckooResult res;
    res.failed_out = failed_fps;
    res.failed_cap = failed_fps ? failed_cap : 0;
    int st = PB_OK;
    const uint64_t step = 1ull << 28;
    for (uint64_t c0 = 0; c0 < n && st == PB_OK; c0 += step) {
        const uint64_t cn = std::min(step, n - c0);
        uint32_t *d = const_cast<uint32_t *>(fps) + c0;
        if (!on_device) {
            st = scratch_reserve(ctx, ctx->out_stage[0], cn * 4);
            if (st != PB_OK) break;
            d = (uint32_t *)ctx->out_stage[0].p;
            cudaError_t e = cudaMemcpyAsync(d, fps + c0, cn * 4, cudaMemcpyHostToDevice, ctx->stream);
            if (e != cudaSuccess) {
                set_error("H2D copy failed: %s", cudaGetErrorString(e));
                st = PB_ERR_CUDA;
                break;
            }
        }
        st = add_fps_device(c, nullptr, d, cn, 0, &res);
    }
    return finish_add(st, res, n_added, n_failed);
}

int pb_cuckoo_check_keys(pb_cuckoo *c, const pb_keys *keys, uint8_t *out, int out_on_device) {
    PB_REQUIRE(c && keys, "NULL argument");
    PB_REQUIRE(out || keys->n == 0, "out is NULL");
    DeviceGuard g(c->ctx->device);
    CuckooCheckArgs a{c, out_on_device ? out : nullptr, out_on_device ? nullptr : out};
    PB_TRY(for_each_chunk(c->ctx, keys, cuckoo_check_chunk, &a));
    if (!out_on_device) PB_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return PB_OK;
}

int pb_cuckoo_check_fingerprints(pb_cuckoo *c, const uint32_t *fps, uint64_t n, int on_device, uint8_t *out, int out_on_device) {
    PB_REQUIRE(c && ((fps && out) || n == 0), "NULL argument");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    const uint32_t *d = fps;
    if (!on_device) {
        PB_TRY(scratch_reserve(ctx, ctx->aux_stage[0], n * 4));
        PB_CUDA(cudaMemcpyAsync(ctx->aux_stage[0].p, fps, n * 4, cudaMemcpyHostToDevice, ctx->stream));
        d = (const uint32_t *)ctx->aux_stage[0].p;
    }
    uint8_t *o = out;
    if (!out_on_device) {
        PB_TRY(scratch_reserve(ctx, ctx->out_stage[0], n));
        o = (uint8_t *)ctx->out_stage[0].p;
    }
    const CuckooDev cd = dev_view(c);
    PB_BS_DISPATCH(c, cuckoo_check_fps, grid_for(ctx, n, 256, 8), 256, d, n, cd, o);
    PB_TRY(check_launch(ctx, "cuckoo_check_fps"));
    if (!out_on_device) {
        PB_CUDA(cudaMemcpyAsync(out, o, n, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return PB_OK;
}

int pb_cuckoo_fingerprint_info(pb_cuckoo *c, const pb_keys *keys, uint32_t *fp, uint64_t *idx1, uint64_t *idx2, int out_on_device) {
    PB_REQUIRE(c && keys, "NULL argument");
    PB_REQUIRE((fp && idx1 && idx2) || keys->n == 0, "NULL output");
    DeviceGuard g(c->ctx->device);
    CuckooInfoArgs a{c, fp, idx1, idx2, out_on_device};
    // host outputs are copied back per chunk from per-slot scratch: keep chunks in step with the slots
    PB_TRY(for_each_chunk(c->ctx, keys, cuckoo_info_chunk, &a));
    if (!out_on_device) PB_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return PB_OK;
}

// stage a host array on the device (aux/out scratch slot `which`)
static int stage_host(pb_ctx *ctx, pb_scratch &sc, const void *host, size_t bytes, void **dev) {
    PB_TRY(scratch_reserve(ctx, sc, bytes ? bytes : 16));
    if (bytes) PB_CUDA(cudaMemcpyAsync(sc.p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *dev = sc.p;
    return PB_OK;
}

static int ensure_alt(pb_cuckoo *c) {
    if (c->alt) return PB_OK;
    PB_CUDA(cudaMalloc(&c->alt, ((c->nslots + 3) & ~(uint64_t)3) * 8));
    PB_CUDA(cudaMemsetAsync(c->alt, 0, c->nslots * 8, c->ctx->stream));
    return PB_OK;
}

// CuckooFilter.remove (cuckoo.py:317-330) for every key: out[i] = 1 when a stored copy of the key's fingerprint was
// cleared.  Equal keys inside one batch: exactly one of them wins (the reference: the first).
int pb_cuckoo_remove_keys(pb_cuckoo *c, const pb_keys *keys, uint8_t *out, int out_on_device, uint64_t *n_removed) {
    PB_REQUIRE(c && keys && n_removed, "NULL argument");
    PB_REQUIRE(out || keys->n == 0, "out is NULL");
    PB_REQUIRE(keys->on_device || !out_on_device, "device output needs device keys");
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    *n_removed = 0;
    if (keys->n == 0) return validate_keys(keys);
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    CuckooCounters *cnt = counters_dev(ctx);
    PB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(CuckooCounters), ctx->stream));
    struct Args {
        pb_cuckoo *c;
        uint8_t *out_dev, *out_host;
        CuckooCounters *cnt;
    } a{c, out_on_device ? out : nullptr, out_on_device ? nullptr : out, cnt};
    auto fn = [](pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) -> int {
        Args *a = (Args *)user;
        pb_cuckoo *c = a->c;
        PB_TRY(scratch_reserve(ctx, ctx->aux_stage[slot], dk.n * 4));
        uint32_t *fps = (uint32_t *)ctx->aux_stage[slot].p;
        if (is_fixed16(dk)) {
            cuckoo_fp_fixed16<<<grid_for(ctx, dk.n, 256, 8), 256, 0, ctx->stream>>>((const uint4 *)dk.data, dk.n, c->fp_bits, fps);
        } else {
            const uint64_t tiles = (dk.n + kTileKeys - 1) / kTileKeys;
            const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->num_sms * 4);
            if (dk.sym_width == 4) cuckoo_fp_staged<4><<<grid, kTileKeys, 0, ctx->stream>>>(dk, c->fp_bits, fps);
            else cuckoo_fp_staged<1><<<grid, kTileKeys, 0, ctx->stream>>>(dk, c->fp_bits, fps);
        }
        PB_TRY(check_launch(ctx, "cuckoo_fp"));
        uint8_t *o = a->out_dev ? a->out_dev + first : nullptr;
        if (!o) {
            PB_TRY(scratch_reserve(ctx, ctx->out_stage[slot], dk.n));
            o = (uint8_t *)ctx->out_stage[slot].p;
        }
        const CuckooDev cd = dev_view(c);
        PB_BS_DISPATCH(c, cuckoo_remove_fps, grid_for(ctx, dk.n, 256, 8), 256, fps, (const uint64_t *)nullptr, dk.n, cd, o, a->cnt);
        PB_TRY(check_launch(ctx, "cuckoo_remove"));
        if (a->out_host) PB_CUDA(cudaMemcpyAsync(a->out_host + first, o, dk.n, cudaMemcpyDeviceToHost, ctx->stream));
        return PB_OK;
    };
    PB_TRY(for_each_chunk(ctx, keys, fn, &a));
    CuckooCounters *h = (CuckooCounters *)((uint8_t *)ctx->pinned_small + 256);
    PB_CUDA(cudaMemcpyAsync(h, cnt, sizeof(CuckooCounters), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_removed = h->n_placed;
    return PB_OK;
}

// ---- pre-indexed entry points (custom hash_function: the caller hashes, see cuckoo_serial_indexed); host arrays
int pb_cuckoo_add_indexed(pb_cuckoo *c, const uint32_t *fps, const uint64_t *i2, uint64_t n, uint64_t *n_added, uint64_t *n_failed,
                          uint32_t *failed_fps, uint64_t *failed_i2, uint64_t failed_cap) {
    PB_REQUIRE(c && n_added && n_failed && ((fps && i2) || n == 0), "NULL argument");
    *n_added = *n_failed = 0;
    if (n == 0) return PB_OK;
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    PB_TRY(ensure_alt(c));
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    CuckooCounters *cnt = counters_dev(ctx);
    PB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(CuckooCounters), ctx->stream));
    void *dfp, *di2;
    PB_TRY(stage_host(ctx, ctx->out_stage[0], fps, n * 4, &dfp));
    PB_TRY(stage_host(ctx, ctx->aux_stage[0], i2, n * 8, &di2));
    const uint64_t fcap = std::max<uint64_t>(failed_cap, 1);
    PB_TRY(scratch_reserve(ctx, ctx->out_stage[1], fcap * 4));
    PB_TRY(scratch_reserve(ctx, ctx->aux_stage[1], fcap * 8));
    const uint64_t seed = c->rng_seed + (++c->epoch) * 0xD1B54A32D192ED03ULL;
    const CuckooDev cd = dev_view(c);
    PB_BS_DISPATCH(c, cuckoo_serial_indexed, 1, 32, (const uint32_t *)dfp, (const uint64_t *)di2, n, cd, seed,
                   (uint32_t *)ctx->out_stage[1].p, (uint64_t *)ctx->aux_stage[1].p, fcap, cnt);
    PB_TRY(check_launch(ctx, "cuckoo_serial_indexed"));
    CuckooCounters *h = (CuckooCounters *)((uint8_t *)ctx->pinned_small + 256);
    PB_CUDA(cudaMemcpyAsync(h, cnt, sizeof(CuckooCounters), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_added = h->n_placed;
    *n_failed = h->n_failed;
    if (h->n_failed) {
        const uint64_t take = std::min<uint64_t>(h->n_failed, (failed_fps && failed_i2) ? failed_cap : 0);
        if (take) {
            PB_CUDA(cudaMemcpyAsync(failed_fps, ctx->out_stage[1].p, take * 4, cudaMemcpyDeviceToHost, ctx->stream));
            PB_CUDA(cudaMemcpyAsync(failed_i2, ctx->aux_stage[1].p, take * 8, cudaMemcpyDeviceToHost, ctx->stream));
            PB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        set_error("The CuckooFilter is currently full (%llu fingerprints left homeless)", (unsigned long long)h->n_failed);
        return PB_ERR_CUCKOO_FULL;
    }
    return PB_OK;
}

int pb_cuckoo_check_indexed(pb_cuckoo *c, const uint32_t *fps, const uint64_t *i2, uint64_t n, uint8_t *out) {
    PB_REQUIRE(c && ((fps && i2 && out) || n == 0), "NULL argument");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    void *dfp, *di2;
    PB_TRY(stage_host(ctx, ctx->out_stage[0], fps, n * 4, &dfp));
    PB_TRY(stage_host(ctx, ctx->aux_stage[0], i2, n * 8, &di2));
    PB_TRY(scratch_reserve(ctx, ctx->out_stage[1], n));
    const CuckooDev cd = dev_view(c);
    PB_BS_DISPATCH(c, cuckoo_check_indexed, grid_for(ctx, n, 256, 8), 256, (const uint32_t *)dfp, (const uint64_t *)di2, n, cd,
                   (uint8_t *)ctx->out_stage[1].p);
    PB_TRY(check_launch(ctx, "cuckoo_check_indexed"));
    PB_CUDA(cudaMemcpyAsync(out, ctx->out_stage[1].p, n, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

int pb_cuckoo_remove_indexed(pb_cuckoo *c, const uint32_t *fps, const uint64_t *i2, uint64_t n, uint8_t *out, uint64_t *n_removed) {
    PB_REQUIRE(c && n_removed && ((fps && i2 && out) || n == 0), "NULL argument");
    *n_removed = 0;
    if (n == 0) return PB_OK;
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    CuckooCounters *cnt = counters_dev(ctx);
    PB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(CuckooCounters), ctx->stream));
    void *dfp, *di2;
    PB_TRY(stage_host(ctx, ctx->out_stage[0], fps, n * 4, &dfp));
    PB_TRY(stage_host(ctx, ctx->aux_stage[0], i2, n * 8, &di2));
    PB_TRY(scratch_reserve(ctx, ctx->out_stage[1], n));
    const CuckooDev cd = dev_view(c);
    PB_BS_DISPATCH(c, cuckoo_remove_fps, grid_for(ctx, n, 256, 8), 256, (const uint32_t *)dfp, (const uint64_t *)di2, n, cd,
                   (uint8_t *)ctx->out_stage[1].p, cnt);
    PB_TRY(check_launch(ctx, "cuckoo_remove"));
    CuckooCounters *h = (CuckooCounters *)((uint8_t *)ctx->pinned_small + 256);
    PB_CUDA(cudaMemcpyAsync(out, ctx->out_stage[1].p, n, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaMemcpyAsync(h, cnt, sizeof(CuckooCounters), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_removed = h->n_placed;
    return PB_OK;
}

// idx_2 of the fingerprint in every slot, for a table that was uploaded (a file written by the reference)
int pb_cuckoo_set_alt(pb_cuckoo *c, const uint64_t *alt, uint64_t count) {
    PB_REQUIRE(c && alt, "NULL argument");
    PB_REQUIRE(count == c->nslots, "expected %llu entries, got %llu", (unsigned long long)c->nslots, (unsigned long long)count);
    DeviceGuard g(c->ctx->device);
    PB_TRY(ensure_alt(c));
    PB_CUDA(cudaMemcpyAsync(c->alt, alt, count * 8, cudaMemcpyHostToDevice, c->ctx->stream));
    PB_CUDA(cudaStreamSynchronize(c->ctx->stream));
    return PB_OK;
}

// an empty table of new_capacity buckets (pre-indexed filters expand by re-inserting from the host, :455-481)
int pb_cuckoo_resize(pb_cuckoo *c, uint64_t new_capacity) {
    PB_REQUIRE(c && new_capacity >= 1, "bad argument");
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    uint32_t *new_slots = nullptr;
    uint64_t new_n = 0;
    PB_TRY(cuckoo_alloc_table(ctx, new_capacity, c->bucket_size, &new_slots, &new_n));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    PB_CUDA(cudaFree(c->slots));
    if (c->alt) {
        PB_CUDA(cudaFree(c->alt));
        c->alt = nullptr;
    }
    c->slots = new_slots;
    c->nslots = new_n;
    c->capacity = new_capacity;
    c->fm = make_fastmod(new_capacity);
    PB_CUDA(cudaMemsetAsync(c->zero_flag, 0, 4, ctx->stream));
    return PB_OK;
}

int pb_cuckoo_count(pb_cuckoo *c, uint64_t *out) {
    PB_REQUIRE(c && out, "NULL argument");
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    unsigned long long *acc = (unsigned long long *)ctx->small.p;
    PB_CUDA(cudaMemsetAsync(acc, 0, 8, ctx->stream));
    count_nonzero_kernel<<<grid_for(ctx, c->nslots, 256, 8), 256, 0, ctx->stream>>>(c->slots, c->nslots, acc);
    PB_TRY(check_launch(ctx, "cuckoo_count"));
    uint64_t *h = (uint64_t *)ctx->pinned_small;
    PB_CUDA(cudaMemcpyAsync(h, acc, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaMemcpyAsync(h + 1, c->zero_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = h[0] + (((uint32_t *)(h + 1))[0] ? 1u : 0u);
    return PB_OK;
}

int pb_cuckoo_download(pb_cuckoo *c, uint32_t *slots, uint64_t count, int *has_zero_fp) {
    PB_REQUIRE(c && (slots || count == 0), "NULL argument");
    PB_REQUIRE(count == c->nslots || (count == 0 && !slots), "expected %llu slots, got %llu", (unsigned long long)c->nslots,
               (unsigned long long)count);
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    if (count) PB_CUDA(cudaMemcpyAsync(slots, c->slots, count * 4, cudaMemcpyDeviceToHost, ctx->stream));
    uint32_t *h = (uint32_t *)ctx->pinned_small;
    PB_CUDA(cudaMemcpyAsync(h, c->zero_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (has_zero_fp) *has_zero_fp = h[0] ? 1 : 0;
    return PB_OK;
}

int pb_cuckoo_upload(pb_cuckoo *c, const uint32_t *slots, uint64_t count, int has_zero_fp) {
    PB_REQUIRE(c && slots, "NULL argument");
    PB_REQUIRE(count == c->nslots, "expected %llu slots, got %llu", (unsigned long long)c->nslots, (unsigned long long)count);
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    PB_CUDA(cudaMemcpyAsync(c->slots, slots, count * 4, cudaMemcpyHostToDevice, ctx->stream));
    uint32_t *h = (uint32_t *)ctx->pinned_small;
    h[0] = has_zero_fp ? 1u : 0u;
    PB_CUDA(cudaMemcpyAsync(c->zero_flag, h, 4, cudaMemcpyHostToDevice, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

int pb_cuckoo_device_ptr(pb_cuckoo *c, void **out_dev, uint64_t *out_count) {
    PB_REQUIRE(c && out_dev, "NULL argument");
    *out_dev = c->slots;
    if (out_count) *out_count = c->nslots;
    return PB_OK;
}

int pb_cuckoo_capacity(pb_cuckoo *c, uint64_t *out) {
    PB_REQUIRE(c && out, "NULL argument");
    *out = c->capacity;
    return PB_OK;
}

// cuckoo.py:455-481: capacity becomes new_capacity, every stored fingerprint is re-inserted from scratch.
int pb_cuckoo_expand(pb_cuckoo *c, uint64_t new_capacity, uint64_t *n_failed, uint32_t *failed_fps, uint64_t failed_cap) {
    PB_REQUIRE(c, "handle is NULL");
    PB_REQUIRE(new_capacity >= 1, "new capacity must be >= 1");
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    uint32_t *old_slots = c->slots;
    const uint64_t old_n = c->nslots;
    uint32_t *new_slots = nullptr;
    uint64_t new_n = 0;
    PB_TRY(cuckoo_alloc_table(ctx, new_capacity, c->bucket_size, &new_slots, &new_n));
    c->slots = new_slots;
    c->nslots = new_n;
    c->capacity = new_capacity;
    c->fm = make_fastmod(new_capacity);
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    CuckooCounters *cnt = counters_dev(ctx);
    PB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(CuckooCounters), ctx->stream));
    const uint64_t fcap = std::max<uint64_t>(failed_cap, 1);
    PB_TRY(scratch_reserve(ctx, ctx->aux_stage[0], fcap * 4));
    uint32_t *failed = (uint32_t *)ctx->aux_stage[0].p;
    const uint64_t seed = c->rng_seed + (++c->epoch) * 0xD1B54A32D192ED03ULL;
    const CuckooDev cd = dev_view(c);
    if (ctx->cuckoo_serial) {
        // one thread walks the old slot array in bucket order, like :467-481
        PB_BS_DISPATCH(c, cuckoo_insert_kernel, 1, 1, old_slots, nullptr, old_n, 1, cd, seed, failed, fcap, cnt);
    } else {
        PB_BS_DISPATCH(c, cuckoo_insert_kernel, grid_for(ctx, old_n, 256, 8), 256, old_slots, nullptr, old_n, 1, cd, seed, failed,
                       fcap, cnt);
    }
    PB_TRY(check_launch(ctx, "cuckoo_expand"));
    CuckooCounters *h = (CuckooCounters *)((uint8_t *)ctx->pinned_small + 256);
    PB_CUDA(cudaMemcpyAsync(h, cnt, sizeof(CuckooCounters), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    PB_CUDA(cudaFree(old_slots));
    if (n_failed) *n_failed = h->n_failed;
    if (h->n_failed) {
        const uint64_t take = std::min<uint64_t>(h->n_failed, failed_fps ? failed_cap : 0);
        if (take) {
            PB_CUDA(cudaMemcpyAsync(failed_fps, failed, take * 4, cudaMemcpyDeviceToHost, ctx->stream));
            PB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        set_error("The CuckooFilter failed to expand (%llu fingerprints left homeless)", (unsigned long long)h->n_failed);
        return PB_ERR_CUCKOO_FULL;
    }
    return PB_OK;
}

// ---------------------------------------------------------------- CountingCuckooFilter: the fingerprint -> count map
static FpCounts counts_view(const pb_cuckoo *c) {
    FpCounts m;
    m.keys = c->cnt_keys;
    m.vals = c->cnt_vals;
    m.tickets = c->cnt_tickets;
    m.mask = c->cnt_cap - 1;
    return m;
}

static unsigned long long *counts_scratch(pb_ctx *ctx) { return (unsigned long long *)ctx->small.p + 44; }  // [44..46]

// (re)builds the map with `cap` hashed entries, carrying over every non-zero count
static int counts_rebuild(pb_cuckoo *c, uint64_t cap) {
    pb_ctx *ctx = c->ctx;
    uint32_t *k = nullptr, *v = nullptr, *t = nullptr;
    const size_t bytes = (size_t)(cap + 1) * 4;
    if (cudaMalloc(&k, bytes) != cudaSuccess || cudaMalloc(&v, bytes) != cudaSuccess || cudaMalloc(&t, bytes) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(k), cudaFree(v), cudaFree(t);
        set_error("cudaMalloc of the %llu-entry count map failed", (unsigned long long)cap);
        return PB_ERR_OOM;
    }
    PB_CUDA(cudaMemsetAsync(k, 0, bytes, ctx->stream));
    PB_CUDA(cudaMemsetAsync(v, 0, bytes, ctx->stream));
    PB_CUDA(cudaMemsetAsync(t, 0, bytes, ctx->stream));
    uint64_t used = 0;
    if (c->cnt_cap) {
        PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
        unsigned long long *acc = counts_scratch(ctx);
        PB_CUDA(cudaMemsetAsync(acc, 0, 8, ctx->stream));
        FpCounts to;
        to.keys = k, to.vals = v, to.tickets = t, to.mask = cap - 1;
        launch_begin(ctx);
        counts_rehash_kernel<<<grid_for(ctx, c->cnt_cap + 1, 256, 8), 256, 0, ctx->stream>>>(counts_view(c), to, acc);
        PB_TRY(check_launch(ctx, "cuckoo_counts_rehash"));
        PB_CUDA(cudaMemcpyAsync(ctx->pinned_small, acc, 8, cudaMemcpyDeviceToHost, ctx->stream));
        PB_CUDA(cudaStreamSynchronize(ctx->stream));
        used = *(uint64_t *)ctx->pinned_small;
        cudaFree(c->cnt_keys), cudaFree(c->cnt_vals), cudaFree(c->cnt_tickets);
    }
    c->cnt_keys = k, c->cnt_vals = v, c->cnt_tickets = t;
    c->cnt_cap = cap;
    c->cnt_used = used;
    return PB_OK;
}

static uint64_t counts_cap_for(uint64_t entries) {
    uint64_t cap = 1024;
    while (cap < entries * 2) cap <<= 1;
    return cap;
}

// room for `incoming` more distinct fingerprints at a load of at most 3/4 (probe sequences stay short and end)
static int counts_reserve(pb_cuckoo *c, uint64_t incoming) {
    PB_REQUIRE(c->cnt_cap, "pb_cuckoo_counts_enable was not called on this filter");
    if ((c->cnt_used + incoming) * 4 <= c->cnt_cap * 3) return PB_OK;
    PB_TRY(counts_rebuild(c, c->cnt_cap));  // drops the entries whose count is 0
    if ((c->cnt_used + incoming) * 4 <= c->cnt_cap * 3) return PB_OK;
    return counts_rebuild(c, counts_cap_for(c->cnt_used + incoming));
}

static int chunk_fps(pb_ctx *ctx, pb_cuckoo *c, const DevKeys &dk, int slot, uint32_t **out) {
    PB_TRY(scratch_reserve(ctx, ctx->aux_stage[slot], dk.n * 8));
    uint32_t *fps = (uint32_t *)ctx->aux_stage[slot].p;  // [fps n][tickets n]
    launch_begin(ctx);
    if (is_fixed16(dk)) {
        cuckoo_fp_fixed16<<<grid_for(ctx, dk.n, 256, 8), 256, 0, ctx->stream>>>((const uint4 *)dk.data, dk.n, c->fp_bits, fps);
    } else {
        const uint64_t tiles = (dk.n + kTileKeys - 1) / kTileKeys;
        const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->num_sms * 4);
        if (dk.sym_width == 4) cuckoo_fp_staged<4><<<grid, kTileKeys, 0, ctx->stream>>>(dk, c->fp_bits, fps);
        else cuckoo_fp_staged<1><<<grid, kTileKeys, 0, ctx->stream>>>(dk, c->fp_bits, fps);
    }
    *out = fps;
    return check_launch(ctx, "cuckoo_fp");
}

static int counts_sync_used(pb_cuckoo *c, unsigned long long *acc) {
    pb_ctx *ctx = c->ctx;
    PB_CUDA(cudaMemcpyAsync(ctx->pinned_small, acc, 24, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    c->cnt_used += ((uint64_t *)ctx->pinned_small)[0];
    return PB_OK;
}

// Turns the filter into a counting one (or re-fits the map after pb_cuckoo_expand / pb_cuckoo_resize).
int pb_cuckoo_counts_enable(pb_cuckoo *c) {
    PB_REQUIRE(c, "handle is NULL");
    DeviceGuard g(c->ctx->device);
    const uint64_t want = counts_cap_for(c->nslots);  // (fingerprint 0 has its own entry past the hashed ones)
    if (c->cnt_cap >= want) return PB_OK;
    return counts_rebuild(c, want);
}

struct CountsArgs {
    pb_cuckoo *c;
    unsigned long long *acc;
    uint32_t *out_dev, *out_host;
    uint8_t *flag_dev, *flag_host;
};

// CountingCuckooFilter.add, the counting half: count[fp(key)] += 1 for every key (the set half is pb_cuckoo_add_keys)
int pb_cuckoo_counts_add_keys(pb_cuckoo *c, const pb_keys *keys) {
    PB_REQUIRE(c && keys, "NULL argument");
    PB_REQUIRE(c->cnt_cap, "pb_cuckoo_counts_enable was not called on this filter");
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    if (keys->n == 0) return validate_keys(keys);
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    CountsArgs a{c, counts_scratch(ctx), nullptr, nullptr, nullptr, nullptr};
    auto fn = [](pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) -> int {
        (void)first;
        CountsArgs *a = (CountsArgs *)user;
        // room for the case that every key of the chunk is a new fingerprint; chunks are at most a quarter of the map,
        // so a stream of repeats never makes the map grow with the batch size
        PB_TRY(counts_reserve(a->c, dk.n));
        PB_CUDA(cudaMemsetAsync(a->acc, 0, 24, ctx->stream));
        uint32_t *fps;
        PB_TRY(chunk_fps(ctx, a->c, dk, slot, &fps));
        launch_begin(ctx);
        counts_add_kernel<<<grid_for(ctx, dk.n, 256, 8), 256, 0, ctx->stream>>>(fps, nullptr, dk.n, a->c->fp_bits, counts_view(a->c), a->acc);
        PB_TRY(check_launch(ctx, "cuckoo_counts_add"));
        return counts_sync_used(a->c, a->acc);
    };
    const uint64_t chunk = std::min<uint64_t>(1ull << 28, std::max<uint64_t>(c->cnt_cap / 4, 1ull << 16));
    return for_each_chunk(ctx, keys, fn, &a, chunk);
}

// count[fp] += amounts[i] (NULL: 1) for host or device arrays of fingerprints (plugin hash path, load)
int pb_cuckoo_counts_add_fingerprints(pb_cuckoo *c, const uint32_t *fps, const uint32_t *amounts, uint64_t n, int on_device) {
    PB_REQUIRE(c && (fps || n == 0), "NULL argument");
    PB_REQUIRE(c->cnt_cap, "pb_cuckoo_counts_enable was not called on this filter");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    unsigned long long *acc = counts_scratch(ctx);
    for (uint64_t lo = 0; lo < n;) {
        const uint64_t cn = std::min<uint64_t>(n - lo, std::max<uint64_t>(c->cnt_cap / 4, 1ull << 16));
        PB_TRY(counts_reserve(c, cn));
        PB_CUDA(cudaMemsetAsync(acc, 0, 24, ctx->stream));
        const uint32_t *d_fps = fps + lo, *d_amt = amounts ? amounts + lo : nullptr;
        if (!on_device) {
            PB_TRY(scratch_reserve(ctx, ctx->aux_stage[0], cn * 8));
            uint32_t *d = (uint32_t *)ctx->aux_stage[0].p;
            PB_CUDA(cudaMemcpyAsync(d, fps + lo, cn * 4, cudaMemcpyHostToDevice, ctx->stream));
            d_fps = d;
            if (amounts) {
                PB_CUDA(cudaMemcpyAsync(d + cn, amounts + lo, cn * 4, cudaMemcpyHostToDevice, ctx->stream));
                d_amt = d + cn;
            }
        }
        launch_begin(ctx);
        counts_add_kernel<<<grid_for(ctx, cn, 256, 8), 256, 0, ctx->stream>>>(d_fps, d_amt, cn, c->fp_bits, counts_view(c), acc);
        PB_TRY(check_launch(ctx, "cuckoo_counts_add"));
        PB_TRY(counts_sync_used(c, acc));
        lo += cn;
    }
    return PB_OK;
}

// CountingCuckooFilter.check (:175-191): the stored count of every key's fingerprint, 0 when it is not stored
int pb_cuckoo_counts_get_keys(pb_cuckoo *c, const pb_keys *keys, uint32_t *out, int out_on_device) {
    PB_REQUIRE(c && keys && (out || keys->n == 0), "NULL argument");
    PB_REQUIRE(c->cnt_cap, "pb_cuckoo_counts_enable was not called on this filter");
    PB_REQUIRE(keys->on_device || !out_on_device, "device output needs device keys");
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    CountsArgs a{c, nullptr, out_on_device ? out : nullptr, out_on_device ? nullptr : out, nullptr, nullptr};
    auto fn = [](pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) -> int {
        CountsArgs *a = (CountsArgs *)user;
        uint32_t *fps;
        PB_TRY(chunk_fps(ctx, a->c, dk, slot, &fps));
        uint32_t *o = a->out_dev ? a->out_dev + first : nullptr;
        if (!o) {
            PB_TRY(scratch_reserve(ctx, ctx->out_stage[slot], dk.n * 4));
            o = (uint32_t *)ctx->out_stage[slot].p;
        }
        launch_begin(ctx);
        counts_get_kernel<<<grid_for(ctx, dk.n, 256, 8), 256, 0, ctx->stream>>>(fps, dk.n, a->c->fp_bits, counts_view(a->c), o);
        PB_TRY(check_launch(ctx, "cuckoo_counts_get"));
        if (a->out_host) PB_CUDA(cudaMemcpyAsync(a->out_host + first, o, dk.n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        return PB_OK;
    };
    return for_each_chunk(ctx, keys, fn, &a, 1ull << 28);
}

// the same for host arrays of fingerprints (export: the count next to every stored fingerprint; plugin hash path)
int pb_cuckoo_counts_get_fingerprints(pb_cuckoo *c, const uint32_t *fps, uint64_t n, uint32_t *out) {
    PB_REQUIRE(c && ((fps && out) || n == 0), "NULL argument");
    PB_REQUIRE(c->cnt_cap, "pb_cuckoo_counts_enable was not called on this filter");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    PB_TRY(scratch_reserve(ctx, ctx->aux_stage[0], n * 8));
    uint32_t *d = (uint32_t *)ctx->aux_stage[0].p;
    PB_CUDA(cudaMemcpyAsync(d, fps, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    launch_begin(ctx);
    counts_get_kernel<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(d, n, c->fp_bits, counts_view(c), d + n);
    PB_TRY(check_launch(ctx, "cuckoo_counts_get"));
    PB_CUDA(cudaMemcpyAsync(out, d + n, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}

// count[fp] = vals[i] (NULL: 0) for host arrays: load (:286-303), and forgetting fingerprints that were given up as
// homeless (:264-265 hands the bin back to the caller, who raises)
int pb_cuckoo_counts_set(pb_cuckoo *c, const uint32_t *fps, const uint32_t *vals, uint64_t n) {
    PB_REQUIRE(c && (fps || n == 0), "NULL argument");
    PB_REQUIRE(c->cnt_cap, "pb_cuckoo_counts_enable was not called on this filter");
    if (n == 0) return PB_OK;
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    unsigned long long *acc = counts_scratch(ctx);
    for (uint64_t lo = 0; lo < n;) {
        const uint64_t cn = std::min<uint64_t>(n - lo, std::max<uint64_t>(c->cnt_cap / 4, 1ull << 16));
        PB_TRY(counts_reserve(c, cn));
        PB_CUDA(cudaMemsetAsync(acc, 0, 24, ctx->stream));
        PB_TRY(scratch_reserve(ctx, ctx->aux_stage[0], cn * 8));
        uint32_t *d = (uint32_t *)ctx->aux_stage[0].p;
        PB_CUDA(cudaMemcpyAsync(d, fps + lo, cn * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (vals) PB_CUDA(cudaMemcpyAsync(d + cn, vals + lo, cn * 4, cudaMemcpyHostToDevice, ctx->stream));
        launch_begin(ctx);
        counts_set_kernel<<<grid_for(ctx, cn, 256, 8), 256, 0, ctx->stream>>>(d, vals ? d + cn : nullptr, cn, counts_view(c), acc);
        PB_TRY(check_launch(ctx, "cuckoo_counts_set"));
        PB_TRY(counts_sync_used(c, acc));
        lo += cn;
    }
    return PB_OK;
}

static int counts_remove_device(pb_cuckoo *c, const uint32_t *fps, const uint64_t *i2, uint64_t n, uint32_t *tickets, uint8_t *out,
                                unsigned long long *acc) {
    pb_ctx *ctx = c->ctx;
    launch_begin(ctx);
    counts_remove_tickets<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(fps, n, c->fp_bits, counts_view(c), tickets, out);
    PB_TRY(check_launch(ctx, "cuckoo_counts_remove_tickets"));
    const CuckooDev cd = dev_view(c);
    PB_BS_DISPATCH(c, counts_remove_settle, grid_for(ctx, n, 256, 8), 256, fps, i2, n, cd, counts_view(c), tickets, acc + 1);
    return check_launch(ctx, "cuckoo_counts_remove_settle");
}

// CountingCuckooFilter.remove (:193-210) for every key: out[i] = 1 when a stored count was decremented.  A key that
// occurs more often in the batch than its count reports 1 exactly `count` times.  n_removed = decrements,
// n_bins_removed = fingerprints whose count reached 0 and that left the table.
int pb_cuckoo_counts_remove_keys(pb_cuckoo *c, const pb_keys *keys, uint8_t *out, int out_on_device, uint64_t *n_removed,
                                 uint64_t *n_bins_removed) {
    PB_REQUIRE(c && keys && n_removed && n_bins_removed && (out || keys->n == 0), "NULL argument");
    PB_REQUIRE(c->cnt_cap, "pb_cuckoo_counts_enable was not called on this filter");
    PB_REQUIRE(keys->on_device || !out_on_device, "device output needs device keys");
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    *n_removed = *n_bins_removed = 0;
    if (keys->n == 0) return validate_keys(keys);
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    unsigned long long *acc = counts_scratch(ctx);
    PB_CUDA(cudaMemsetAsync(acc, 0, 24, ctx->stream));
    CountsArgs a{c, acc, nullptr, nullptr, out_on_device ? out : nullptr, out_on_device ? nullptr : out};
    auto fn = [](pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user) -> int {
        CountsArgs *a = (CountsArgs *)user;
        uint32_t *fps;
        PB_TRY(chunk_fps(ctx, a->c, dk, slot, &fps));
        uint8_t *o = a->flag_dev ? a->flag_dev + first : nullptr;
        if (!o) {
            PB_TRY(scratch_reserve(ctx, ctx->out_stage[slot], dk.n));
            o = (uint8_t *)ctx->out_stage[slot].p;
        }
        PB_TRY(counts_remove_device(a->c, fps, nullptr, dk.n, fps + dk.n, o, a->acc));
        if (a->flag_host) PB_CUDA(cudaMemcpyAsync(a->flag_host + first, o, dk.n, cudaMemcpyDeviceToHost, ctx->stream));
        return PB_OK;
    };
    PB_TRY(for_each_chunk(ctx, keys, fn, &a, 1ull << 28));
    PB_CUDA(cudaMemcpyAsync(ctx->pinned_small, acc, 24, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_removed = ((uint64_t *)ctx->pinned_small)[1];
    *n_bins_removed = ((uint64_t *)ctx->pinned_small)[2];
    return PB_OK;
}

// the same for host arrays of fingerprints; i2 (may be NULL) = the caller's idx_2 of every fingerprint (custom hash_function)
int pb_cuckoo_counts_remove_fingerprints(pb_cuckoo *c, const uint32_t *fps, const uint64_t *i2, uint64_t n, uint8_t *out,
                                         uint64_t *n_removed, uint64_t *n_bins_removed) {
    PB_REQUIRE(c && n_removed && n_bins_removed && ((fps && out) || n == 0), "NULL argument");
    PB_REQUIRE(c->cnt_cap, "pb_cuckoo_counts_enable was not called on this filter");
    *n_removed = *n_bins_removed = 0;
    if (n == 0) return PB_OK;
    pb_ctx *ctx = c->ctx;
    DeviceGuard g(ctx->device);
    PB_TRY(scratch_reserve(ctx, ctx->small, 4096));
    unsigned long long *acc = counts_scratch(ctx);
    PB_CUDA(cudaMemsetAsync(acc, 0, 24, ctx->stream));
    PB_TRY(scratch_reserve(ctx, ctx->aux_stage[0], n * 8));
    PB_TRY(scratch_reserve(ctx, ctx->out_stage[0], n * 9));
    uint32_t *d = (uint32_t *)ctx->aux_stage[0].p;
    uint64_t *d_i2 = i2 ? (uint64_t *)ctx->out_stage[0].p : nullptr;
    uint8_t *d_out = (uint8_t *)ctx->out_stage[0].p + n * 8;
    PB_CUDA(cudaMemcpyAsync(d, fps, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (i2) PB_CUDA(cudaMemcpyAsync(d_i2, i2, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    PB_TRY(counts_remove_device(c, d, d_i2, n, d + n, d_out, acc));
    PB_CUDA(cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaMemcpyAsync(ctx->pinned_small, acc, 24, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_removed = ((uint64_t *)ctx->pinned_small)[1];
    *n_bins_removed = ((uint64_t *)ctx->pinned_small)[2];
    return PB_OK;
}

}  // extern "C"
