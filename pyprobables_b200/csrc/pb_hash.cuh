// pb_hash.cuh -- integer building blocks of the hot path, usable from host and device so the
// arithmetic can be unit-tested on a GPU-less machine (pbt_* exports in pb_ctx.cu).
//   * seeded 64-bit FNV-1a                      (reference: probables/hashes.py:86-103)
//   * exact u64 % m by precomputed reciprocal   (reference: `%` at blooms/bloom.py:247, countminsketch.py:275)
//   * cuckoo fingerprint / index derivation     (reference: cuckoo/cuckoo.py:483-506, utilities.py:32-35)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD inline
#endif

namespace pb {

constexpr uint64_t kFnvBasis = 0xCBF29CE484222325ULL;  // hashes.py:96
constexpr uint64_t kFnvPrime = 0x100000001B3ULL;       // hashes.py:97  (= 2^40 + 0x1B3)

PB_HD uint64_t fnv_init(uint64_t seed) { return kFnvBasis + 31ULL * seed; }  // hashes.py:96
PB_HD uint64_t fnv_step(uint64_t h, uint32_t sym) { return (h ^ (uint64_t)sym) * kFnvPrime; }  // :100-102

#if defined(__CUDA_ARCH__)
// One FNV-1a step on a hash held as two 32-bit halves, pinned to four SASS instructions.  The 64-bit product
// h * (2^40 + 0x1B3) is  lo' = low32(x*0x1B3),  hi' = (hi*0x1B3 + high32(x*0x1B3)) + (x << 8)  with x = lo ^ sym.
// Left to itself nvcc reassociates the sum and emits five instructions per step (IMAD, IMAD x*0x100+.., IMAD.WIDE,
// IMAD.IADD, LOP3: 1032 instructions per 16-byte key at k = 7 in the round-1 profile, and the kernel is
// issue bound).  The explicit mad chain below makes ptxas emit exactly LOP3, IMAD.WIDE.U32, IMAD, LEA.
__device__ __forceinline__ void fnv_step_halves(uint32_t &lo, uint32_t &hi, uint32_t sym) {
    const uint32_t x = lo ^ sym;
    uint32_t plo, phi, t, h2;
    asm("{ .reg .b64 p; mul.wide.u32 p, %2, 0x1b3; mov.b64 {%0, %1}, p; }" : "=r"(plo), "=r"(phi) : "r"(x));
    asm("mad.lo.u32 %0, %1, 0x1b3, %2;" : "=r"(t) : "r"(hi), "r"(phi));
    asm("mad.lo.u32 %0, %1, 0x100, %2;" : "=r"(h2) : "r"(x), "r"(t));
    lo = plo;
    hi = h2;
}
#endif

PB_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// Exact h % m for any u64 h and 1 <= m < 2^64.
//   R = floor(2^64 / m) (m >= 2; 2^64-1 for m == 1)   =>   q = mulhi(h, R) is floor(h/m) or one less,
// so one conditional subtract finishes it.  Powers of two take the mask path.
struct FastMod {
    uint64_t m;
    uint64_t recip;
    uint64_t mask;  // m-1 when m is a power of two, else 0 (m == 1 -> mask 0 and recip path gives 0)
    uint32_t is_pow2;
    uint32_t pad;
};

inline FastMod make_fastmod(uint64_t m) {
    FastMod f;
    f.m = m;
    f.is_pow2 = (m & (m - 1)) == 0 ? 1u : 0u;
    f.mask = f.is_pow2 ? m - 1 : 0;
    f.recip = m <= 1 ? ~0ULL : (uint64_t)((((unsigned __int128)1) << 64) / m);
    f.pad = 0;
    return f;
}

PB_HD uint64_t fastmod(uint64_t h, const FastMod &f) {
    if (f.is_pow2) return h & f.mask;
    uint64_t q = mulhi64(h, f.recip);
    uint64_t r = h - q * f.m;
    return r >= f.m ? r - f.m : r;
}

// Exact floor(h / m) with the same reciprocal (used to find the owning shard of a bit index).
PB_HD uint64_t fastdiv(uint64_t h, const FastMod &f) {
    uint64_t q = mulhi64(h, f.recip);
    uint64_t r = h - q * f.m;
    return r >= f.m ? q + 1 : q;
}

// utilities.py:32-35 with right_bits=True as used at cuckoo.py:500
PB_HD uint32_t cuckoo_fingerprint(uint64_t h, uint32_t fp_bits) {
    return fp_bits >= 32 ? (uint32_t)h : (uint32_t)h & ((1u << fp_bits) - 1u);
}

// cuckoo.py:489: fnv_1a(str(fingerprint)) -- FNV-1a (seed 0) over the decimal ASCII digits.
PB_HD uint64_t fnv_of_decimal(uint32_t v) {
    // most significant digit first: peel digits with a descending power of ten
    uint64_t h = kFnvBasis;
    uint32_t p = 1000000000u;
    bool started = false;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t d = v / p;
        v -= d * p;
        if (d != 0 || started || i == 9) {
            h = fnv_step(h, (uint32_t)'0' + d);
            started = true;
        }
        p /= 10u;
    }
    return h;
}

// splitmix64 (synthetic key generator of SURVEY 8(d); not reference code)
PB_HD uint64_t sm64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

}  // namespace pb
