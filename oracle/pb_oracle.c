/*
 * pb_oracle.c -- CPU restatement of pyprobables' hash-then-scatter hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pyprobables_b200/ may link, import or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the timed CPU arm.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below against
 * (a) the golden vectors of the reference's own test-suite (hashes_test.py:27-55,
 * bloom_test.py:256-265 / :323-341, countminsketch_test.py:76-203 / :262-278,
 * cuckoo_test.py:248-266 / :489-498) and (b) fixtures produced by importing the
 * pure-Python reference (tests/golden/make_golden.py -> tests/golden/golden.json).
 *
 * Every function cites the reference lines (relative to /root/reference/) it follows.
 * Written from the behaviour of the reference; no reference source is reproduced
 * (the reference is Python; this is C).
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_FNV_BASIS 0xCBF29CE484222325ULL /* probables/hashes.py:96 */
#define ORC_FNV_PRIME 0x100000001B3ULL      /* probables/hashes.py:97 */

/* A batch of keys.  data = packed symbols (sym_width 1 = bytes / ASCII str,
 * 4 = little-endian u32 code points for non-ASCII str: hashes.py:98 hashes `ord(c)`).
 * offsets == NULL -> fixed `stride` symbols per key, else n+1 symbol offsets. */
typedef struct {
    const void *data;
    const uint64_t *offsets;
    uint64_t n;
    uint32_t stride;
    uint32_t sym_width;
} orc_keys;

static inline void key_span(const orc_keys *k, uint64_t i, uint64_t *beg, uint64_t *len) {
    if (k->offsets) {
        *beg = k->offsets[i];
        *len = k->offsets[i + 1] - k->offsets[i];
    } else {
        *beg = i * (uint64_t)k->stride;
        *len = k->stride;
    }
}

/* probables/hashes.py:86-103: hval0 = basis + 31*seed (mod 2^64); per symbol xor, multiply. */
static inline uint64_t fnv1a_syms(const orc_keys *k, uint64_t beg, uint64_t len, uint64_t seed) {
    uint64_t h = ORC_FNV_BASIS + 31ULL * seed;
    if (k->sym_width == 4) {
        const uint32_t *p = (const uint32_t *)k->data + beg;
        for (uint64_t j = 0; j < len; ++j) { h ^= p[j]; h *= ORC_FNV_PRIME; }
    } else {
        const uint8_t *p = (const uint8_t *)k->data + beg;
        for (uint64_t j = 0; j < len; ++j) { h ^= p[j]; h *= ORC_FNV_PRIME; }
    }
    return h;
}

uint64_t orc_fnv1a(const uint8_t *key, uint64_t len, uint64_t seed) {
    orc_keys k = {key, NULL, 1, (uint32_t)len, 1};
    return fnv1a_syms(&k, 0, len, seed);
}

uint64_t orc_fnv1a_u32(const uint32_t *key, uint64_t len, uint64_t seed) {
    orc_keys k = {key, NULL, 1, (uint32_t)len, 4};
    return fnv1a_syms(&k, 0, len, seed);
}

/* probables/hashes.py:71-83: [fnv_1a(key, s) for s in range(depth)], for a whole batch.
 * out is n x depth row-major. */
void orc_default_fnv1a_many(const orc_keys *k, uint32_t depth, uint64_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)k->n; ++i) {
        uint64_t beg, len;
        key_span(k, (uint64_t)i, &beg, &len);
        for (uint32_t s = 0; s < depth; ++s) out[(uint64_t)i * depth + s] = fnv1a_syms(k, beg, len, s);
    }
}

/* splitmix64 key generator of SURVEY.md 8(d) (not reference code; the shared synthetic input). */
static inline uint64_t sm64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
uint64_t orc_sm64(uint64_t x) { return sm64(x); }

/* uniform 16-byte key i = LE64(sm64(seed+2i)) || LE64(sm64(seed+2i+1)) */
void orc_gen_uniform_keys(uint64_t seed, uint64_t first, uint64_t n, uint8_t *out) {
    uint64_t *o = (uint64_t *)out;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        uint64_t g = first + (uint64_t)i;
        o[2 * i] = sm64(seed + 2 * g);
        o[2 * i + 1] = sm64(seed + 2 * g + 1);
    }
}

/* zipf key for rank r = LE64(r) || LE64(sm64(r)) */
void orc_gen_rank_keys(const uint64_t *ranks, uint64_t n, uint8_t *out) {
    uint64_t *o = (uint64_t *)out;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        o[2 * i] = ranks[i];
        o[2 * i + 1] = sm64(ranks[i]);
    }
}

/* ------------------------------------------------------------------ Bloom */

/* probables/blooms/bloom.py:234-250: for i<k: b = h_i % num_bits; bloom[b//8] |= 1 << (b%8).
 * threads > 1 uses atomic byte ORs (OR commutes, so the bitmap is thread-count independent). */
void orc_bloom_add(uint8_t *bloom, uint64_t num_bits, uint32_t k, const orc_keys *keys) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)keys->n; ++i) {
        uint64_t beg, len;
        key_span(keys, (uint64_t)i, &beg, &len);
        for (uint32_t s = 0; s < k; ++s) {
            uint64_t b = fnv1a_syms(keys, beg, len, s) % num_bits;
            __atomic_fetch_or(&bloom[b >> 3], (uint8_t)(1u << (b & 7)), __ATOMIC_RELAXED);
        }
    }
}

/* probables/blooms/bloom.py:252-272: AND over the k bits (early exit only changes cost). */
void orc_bloom_check(const uint8_t *bloom, uint64_t num_bits, uint32_t k, const orc_keys *keys, uint8_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)keys->n; ++i) {
        uint64_t beg, len;
        key_span(keys, (uint64_t)i, &beg, &len);
        uint8_t ok = 1;
        for (uint32_t s = 0; s < k && ok; ++s) {
            uint64_t b = fnv1a_syms(keys, beg, len, s) % num_bits;
            if (((bloom[b >> 3] >> (b & 7)) & 1u) == 0) ok = 0;
        }
        out[i] = ok;
    }
}

/* bloom.py:241-250 / :261-272 with caller-supplied hashes (n x k u64, already < 2^64). */
void orc_bloom_add_hashes(uint8_t *bloom, uint64_t num_bits, uint32_t k, const uint64_t *h, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i)
        for (uint32_t s = 0; s < k; ++s) {
            uint64_t b = h[i * k + s] % num_bits;
            bloom[b >> 3] |= (uint8_t)(1u << (b & 7));
        }
}
void orc_bloom_check_hashes(const uint8_t *bloom, uint64_t num_bits, uint32_t k, const uint64_t *h, uint64_t n,
                            uint8_t *out) {
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t ok = 1;
        for (uint32_t s = 0; s < k; ++s) {
            uint64_t b = h[i * k + s] % num_bits;
            if (((bloom[b >> 3] >> (b & 7)) & 1u) == 0) ok = 0;
        }
        out[i] = ok;
    }
}

/* bloom.py:552-557 */
uint64_t orc_popcount(const uint8_t *buf, uint64_t nbytes) {
    uint64_t c = 0;
#pragma omp parallel for reduction(+ : c) schedule(static)
    for (int64_t i = 0; i < (int64_t)nbytes; ++i) c += (uint64_t)__builtin_popcount(buf[i]);
    return c;
}

/* ------------------------------------------------------------------ Counting Bloom ("next" row) */

#define ORC_U32_MAX 4294967295ULL

/* probables/blooms/countingbloom.py:135-155, sequential over the batch.  The k counters are read first
 * (vals = bloom[idx] + n), then updated one by one with `bloom[idx] += n` -- so an index hit by two of a key's
 * hashes is incremented twice although both snapshots saw the old value (the reference's own NOTE at :143).
 * post_add (optional) receives min(vals) per key (:155).  A counter pushed past UINT32_MAX by the double
 * increment would raise OverflowError in the reference's array('I'); it is clamped here and counted in the
 * return value so tests can assert it never happens. */
uint64_t orc_cbloom_add(uint32_t *bloom, uint64_t length, uint32_t k, uint64_t *elements_added, const orc_keys *keys,
                        uint64_t num_els, uint64_t *post_add) {
    uint64_t clamped = 0, idx[64], vals[64];
    for (uint64_t i = 0; i < keys->n; ++i) {
        uint64_t beg, len;
        key_span(keys, i, &beg, &len);
        const uint32_t kk = k < 64 ? k : 64;
        for (uint32_t s = 0; s < kk; ++s) {
            idx[s] = fnv1a_syms(keys, beg, len, s) % length; /* :145 */
            vals[s] = (uint64_t)bloom[idx[s]] + num_els;     /* :146 */
        }
        uint64_t mn = ~0ULL;
        for (uint32_t s = 0; s < kk; ++s) {
            if (vals[s] > ORC_U32_MAX) { /* :149-151 */
                bloom[idx[s]] = (uint32_t)ORC_U32_MAX;
                vals[s] = ORC_U32_MAX;
            } else { /* :153 */
                uint64_t v = (uint64_t)bloom[idx[s]] + num_els;
                if (v > ORC_U32_MAX) { v = ORC_U32_MAX; ++clamped; }
                bloom[idx[s]] = (uint32_t)v;
            }
            if (vals[s] < mn) mn = vals[s];
        }
        /* :154 (UINT64 saturation) */
        *elements_added = (*elements_added > ~0ULL - num_els) ? ~0ULL : *elements_added + num_els;
        if (post_add) post_add[i] = mn;
    }
    return clamped;
}

/* countingbloom.py:157-175: min over the k counters */
void orc_cbloom_check(const uint32_t *bloom, uint64_t length, uint32_t k, const orc_keys *keys, uint64_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)keys->n; ++i) {
        uint64_t beg, len, mn = ~0ULL;
        key_span(keys, (uint64_t)i, &beg, &len);
        for (uint32_t s = 0; s < k; ++s) {
            uint64_t v = bloom[fnv1a_syms(keys, beg, len, s) % length];
            if (v < mn) mn = v;
        }
        out[i] = mn;
    }
}

/* countingbloom.py:177-208, sequential: remove min(num_els, current minimum) from every counter of the key that is
 * not saturated; post (optional) receives what remove() returns. */
void orc_cbloom_remove(uint32_t *bloom, uint64_t length, uint32_t k, int64_t *elements_added, const orc_keys *keys,
                       uint64_t num_els, uint64_t *post) {
    uint64_t idx[64];
    for (uint64_t i = 0; i < keys->n; ++i) {
        uint64_t beg, len, mn = ~0ULL;
        key_span(keys, i, &beg, &len);
        const uint32_t kk = k < 64 ? k : 64;
        for (uint32_t s = 0; s < kk; ++s) {
            idx[s] = fnv1a_syms(keys, beg, len, s) % length;
            if (bloom[idx[s]] < mn) mn = bloom[idx[s]];
        }
        uint64_t ret;
        if (mn == ORC_U32_MAX) ret = ORC_U32_MAX;      /* :196-197 */
        else if (mn == 0) ret = 0;                     /* :198-199 */
        else {
            const uint64_t take = mn > num_els ? num_els : mn; /* :201 */
            for (uint32_t s = 0; s < kk; ++s)
                if (bloom[idx[s]] < ORC_U32_MAX) bloom[idx[s]] -= (uint32_t)take; /* :202-204 (twice on a shared index) */
            *elements_added -= (int64_t)take;
            ret = mn - take;
        }
        if (post) post[i] = ret;
    }
}

/* ------------------------------------------------------------------ Count-Min */

#define ORC_I32_MAX 2147483647LL
#define ORC_I32_MIN (-2147483647LL - 1)
#define ORC_I64_MAX 9223372036854775807LL

/* Python floor division (countminsketch.py:436, :443, :450 use `//` on possibly negative ints). */
static inline int64_t floordiv(int64_t a, int64_t b) {
    int64_t q = a / b, r = a % b;
    if (r != 0 && ((r < 0) != (b < 0))) --q;
    return q;
}

static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* countminsketch.py:429-453.  vals = the depth bin values (any order; sorted here as :288/:340 do).
 * query_type: 0 = min (:430-432), 1 = mean (:434-436), 2 = mean-min (:438-453). */
static int64_t cms_query(int64_t *vals, uint32_t depth, uint32_t width, int64_t elements_added, int query_type) {
    qsort(vals, depth, sizeof(int64_t), cmp_i64);
    if (query_type == 1) {
        int64_t s = 0;
        for (uint32_t i = 0; i < depth; ++i) s += vals[i];
        return floordiv(s, (int64_t)depth);
    }
    if (query_type == 2) {
        if (vals[0] == 0 && vals[depth - 1] == 0) return 0;
        for (uint32_t i = 0; i < depth; ++i) {
            int64_t diff = elements_added - vals[i];
            /* width == 1 raises ZeroDivisionError in the reference; callers must not pass it */
            vals[i] = vals[i] - floordiv(diff, (int64_t)width - 1);
        }
        qsort(vals, depth, sizeof(int64_t), cmp_i64);
        if (depth % 2 == 0) return floordiv(vals[depth / 2] + vals[depth / 2 - 1], 2);
        return vals[depth / 2];
    }
    return vals[0];
}

/* countminsketch.py:257-288, sequential over the batch in order.  num_els == NULL -> scalar_num_els for all.
 * post_add (optional, n entries) receives what each add() call returns (:288).
 * Lower bound: the reference stores val unclamped when val <= INT32_MAX and array('i') would raise
 * OverflowError below INT32_MIN; the oracle clamps at INT32_MIN there and reports it via the return
 * value (number of such events) so a test can assert it never happened. */
uint64_t orc_cms_add(int32_t *bins, uint32_t width, uint32_t depth, int64_t *elements_added, const orc_keys *keys,
                     const int64_t *num_els, int64_t scalar_num_els, int query_type, int64_t *post_add) {
    uint64_t underflows = 0;
    int64_t vals[64];
    for (uint64_t i = 0; i < keys->n; ++i) {
        uint64_t beg, len;
        key_span(keys, i, &beg, &len);
        int64_t n = num_els ? num_els[i] : scalar_num_els;
        for (uint32_t r = 0; r < depth; ++r) {
            uint64_t idx = fnv1a_syms(keys, beg, len, r) % width + (uint64_t)r * width; /* :275 */
            int64_t v = (int64_t)bins[idx] + n;                                         /* :276 */
            if (v > ORC_I32_MAX) v = ORC_I32_MAX;                                       /* :280-282 */
            if (v < ORC_I32_MIN) { v = ORC_I32_MIN; ++underflows; }
            bins[idx] = (int32_t)v;
            if (r < 64) vals[r] = v;
        }
        /* :285-287 (Python ints do not wrap; emulate with a saturating add) */
        if (n > 0 && *elements_added > ORC_I64_MAX - n) *elements_added = ORC_I64_MAX;
        else *elements_added += n;
        if (post_add) post_add[i] = cms_query(vals, depth < 64 ? depth : 64, width, *elements_added, query_type);
    }
    return underflows;
}

/* countminsketch.py:323-340 */
void orc_cms_check(const int32_t *bins, uint32_t width, uint32_t depth, int64_t elements_added, const orc_keys *keys,
                   int query_type, int64_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)keys->n; ++i) {
        uint64_t beg, len;
        int64_t vals[64];
        key_span(keys, (uint64_t)i, &beg, &len);
        for (uint32_t r = 0; r < depth && r < 64; ++r)
            vals[r] = bins[fnv1a_syms(keys, beg, len, r) % width + (uint64_t)r * width];
        out[i] = cms_query(vals, depth < 64 ? depth : 64, width, elements_added, query_type);
    }
}

/* Order-free parallel add for the CPU baseline leg (non-negative num_els, no saturation handling beyond
 * the final clamp): same final table as orc_cms_add whenever no bin exceeds INT32_MAX mid-way. */
void orc_cms_add_parallel(int32_t *bins, uint32_t width, uint32_t depth, const orc_keys *keys, int32_t n_each) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)keys->n; ++i) {
        uint64_t beg, len;
        key_span(keys, (uint64_t)i, &beg, &len);
        for (uint32_t r = 0; r < depth; ++r) {
            uint64_t idx = fnv1a_syms(keys, beg, len, r) % width + (uint64_t)r * width;
            __atomic_fetch_add(&bins[idx], n_each, __ATOMIC_RELAXED);
        }
    }
}

/* ------------------------------------------------------------------ Cuckoo */

/* Buckets are kept as the reference keeps them: a list per bucket, append order preserved
 * (cuckoo.py:448-453), so that export() bytes match at loads where no eviction happens. */
typedef struct {
    uint64_t capacity;
    uint32_t bucket_size, max_swaps, fp_bits;
    uint32_t *slots; /* capacity * bucket_size */
    uint8_t *lens;   /* capacity */
    uint64_t inserted;
    uint64_t rng;
} orc_cuckoo;

orc_cuckoo *orc_cuckoo_new(uint64_t capacity, uint32_t bucket_size, uint32_t max_swaps, uint32_t fp_bits,
                           uint64_t rng_seed) {
    orc_cuckoo *c = (orc_cuckoo *)calloc(1, sizeof(orc_cuckoo));
    if (!c) return NULL;
    c->capacity = capacity;
    c->bucket_size = bucket_size;
    c->max_swaps = max_swaps;
    c->fp_bits = fp_bits;
    c->slots = (uint32_t *)calloc(capacity * bucket_size, sizeof(uint32_t));
    c->lens = (uint8_t *)calloc(capacity, 1);
    c->rng = rng_seed ? rng_seed : 0x9E3779B97F4A7C15ULL;
    if (!c->slots || !c->lens) { free(c->slots); free(c->lens); free(c); return NULL; }
    return c;
}
void orc_cuckoo_free(orc_cuckoo *c) {
    if (c) { free(c->slots); free(c->lens); free(c); }
}
uint64_t orc_cuckoo_inserted(const orc_cuckoo *c) { return c->inserted; }

/* utilities.py:32-35 (right_bits=True) applied as cuckoo.py:499-500 does */
static inline uint32_t cuckoo_fp(uint64_t h, uint32_t fp_bits) {
    return (uint32_t)(fp_bits >= 64 ? h : (h & ((1ULL << fp_bits) - 1)));
}

/* cuckoo.py:483-490: idx_1 = fp % capacity; idx_2 = fnv_1a(str(fp)) % capacity (decimal ASCII digits). */
static inline void cuckoo_indices(uint32_t fp, uint64_t capacity, uint64_t *i1, uint64_t *i2) {
    char buf[16];
    int n = 0;
    uint32_t v = fp;
    do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    uint64_t h = ORC_FNV_BASIS;
    while (n) { h ^= (uint8_t)buf[--n]; h *= ORC_FNV_PRIME; }
    *i1 = fp % capacity;
    *i2 = h % capacity;
}

void orc_cuckoo_fingerprint_info(const orc_keys *keys, uint64_t capacity, uint32_t fp_bits, uint32_t *fp,
                                 uint64_t *idx1, uint64_t *idx2) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)keys->n; ++i) {
        uint64_t beg, len;
        key_span(keys, (uint64_t)i, &beg, &len);
        fp[i] = cuckoo_fp(fnv1a_syms(keys, beg, len, 0), fp_bits);
        cuckoo_indices(fp[i], capacity, &idx1[i], &idx2[i]);
    }
}

static inline int bucket_has(const orc_cuckoo *c, uint64_t b, uint32_t fp) {
    const uint32_t *s = c->slots + b * c->bucket_size;
    for (uint32_t j = 0; j < c->lens[b]; ++j)
        if (s[j] == fp) return 1;
    return 0;
}
static inline int bucket_append(orc_cuckoo *c, uint64_t b, uint32_t fp) { /* cuckoo.py:448-453 */
    if (c->lens[b] < c->bucket_size) {
        c->slots[b * c->bucket_size + c->lens[b]++] = fp;
        return 1;
    }
    return 0;
}
static inline uint64_t rng_next(orc_cuckoo *c) { /* xorshift64*; the reference uses Python's global MT
    (cuckoo.py:373,377) which is unseeded in its tests: slot placement is not part of the contract */
    uint64_t x = c->rng;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    c->rng = x;
    return x * 0x2545F4914F6CDD1DULL;
}

/* cuckoo.py:361-392.  Returns 1 when placed; 0 with *homeless set when max_swaps ran out. */
static int cuckoo_insert_fp(orc_cuckoo *c, uint32_t fp, uint64_t i1, uint64_t i2, uint32_t *homeless) {
    if (bucket_append(c, i1, fp) || bucket_append(c, i2, fp)) { c->inserted++; return 1; }
    uint64_t idx = (rng_next(c) >> 33) & 1 ? i2 : i1; /* :373 */
    for (uint32_t s = 0; s < c->max_swaps; ++s) {
        uint32_t slot = (uint32_t)((rng_next(c) >> 32) % c->bucket_size); /* :377 */
        uint32_t *p = &c->slots[idx * c->bucket_size + slot];
        uint32_t victim = *p;
        *p = fp;
        fp = victim;
        uint64_t a, b;
        cuckoo_indices(fp, c->capacity, &a, &b); /* :383 */
        idx = (idx == a) ? b : a;                /* :385 */
        if (bucket_append(c, idx, fp)) { c->inserted++; return 1; }
    }
    *homeless = fp;
    return 0;
}

/* cuckoo.py:291-304 for a batch, in order; stops adding nothing on failure but records the homeless
 * fingerprints (what _deal_with_insertion :508-516 would expand with / raise about).
 * Returns the number of failures; failed[] holds up to failed_cap of them. */
uint64_t orc_cuckoo_add(orc_cuckoo *c, const orc_keys *keys, uint32_t *failed, uint64_t failed_cap) {
    uint64_t nfail = 0;
    for (uint64_t i = 0; i < keys->n; ++i) {
        uint64_t beg, len, i1, i2;
        key_span(keys, i, &beg, &len);
        uint32_t fp = cuckoo_fp(fnv1a_syms(keys, beg, len, 0), c->fp_bits);
        cuckoo_indices(fp, c->capacity, &i1, &i2);
        if (bucket_has(c, i1, fp) || bucket_has(c, i2, fp)) continue; /* :300-302 */
        uint32_t homeless = 0;
        if (!cuckoo_insert_fp(c, fp, i1, i2, &homeless)) {
            if (nfail < failed_cap) failed[nfail] = homeless;
            ++nfail;
        }
    }
    return nfail;
}

/* cuckoo.py:306-315 */
void orc_cuckoo_check(const orc_cuckoo *c, const orc_keys *keys, uint8_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)keys->n; ++i) {
        uint64_t beg, len, i1, i2;
        key_span(keys, (uint64_t)i, &beg, &len);
        uint32_t fp = cuckoo_fp(fnv1a_syms(keys, beg, len, 0), c->fp_bits);
        cuckoo_indices(fp, c->capacity, &i1, &i2);
        out[i] = (uint8_t)(bucket_has(c, i1, fp) || bucket_has(c, i2, fp));
    }
}

/* cuckoo.py:332-349 bucket area of export(): bucket_size u32 per bucket, list order, zero padded.
 * out must hold capacity*bucket_size u32. */
void orc_cuckoo_export_slots(const orc_cuckoo *c, uint32_t *out) {
    for (uint64_t b = 0; b < c->capacity; ++b)
        for (uint32_t j = 0; j < c->bucket_size; ++j)
            out[b * c->bucket_size + j] = j < c->lens[b] ? c->slots[b * c->bucket_size + j] : 0;
}

/* all stored fingerprints (unsorted); returns the count */
uint64_t orc_cuckoo_fingerprints(const orc_cuckoo *c, uint32_t *out) {
    uint64_t n = 0;
    for (uint64_t b = 0; b < c->capacity; ++b)
        for (uint32_t j = 0; j < c->lens[b]; ++j) out[n++] = c->slots[b * c->bucket_size + j];
    return n;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
