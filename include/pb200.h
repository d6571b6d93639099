/*
 * pb200.h -- C ABI of libpb200.so: the B200 (sm_100a) batch engine behind pyprobables'
 * hash-then-scatter hot path.
 *
 * The reference (barrust/pyprobables v0.7.0, pure Python) has no FFI.  Each entry point below is
 * what a ctypes binding for the cited reference method would call; the citing comments name the
 * reference lines (relative to the reference repo root) whose per-key semantics the call
 * reproduces for a whole batch.  INTEGRATION.md shows the reference-side ctypes stub.
 *
 * Conventions
 *  - plain C types only; every function returns PB_OK (0) or a negative pb_status;
 *    pb_last_error() returns a thread-local message for the last failure on this thread.
 *  - handles are opaque; the library owns device memory, the caller owns every buffer it passes
 *    and the library never keeps a caller pointer after the call returns.
 *  - calls are synchronous on return (results visible to the host) unless the buffer arguments
 *    are device pointers (on_device != 0), in which case work is only enqueued on the context's
 *    stream: use pb_ctx_synchronize() or order your own work on pb_ctx_stream().
 *  - a handle is not thread-safe; different handles may be driven from different host threads.
 */
#ifndef PB200_H
#define PB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_VERSION 100 /* 0.1.0 */

typedef enum pb_status {
    PB_OK = 0,
    PB_ERR_BAD_ARG = -1,
    PB_ERR_CUDA = -2,
    PB_ERR_OOM = -3,
    PB_ERR_NO_DEVICE = -4,
    PB_ERR_UNSUPPORTED = -5,
    PB_ERR_CUCKOO_FULL = -6 /* cuckoo/cuckoo.py:515-516; see pb_cuckoo_add_keys */
} pb_status;

typedef struct pb_ctx pb_ctx;
typedef struct pb_bloom pb_bloom;
typedef struct pb_cms pb_cms;
typedef struct pb_cuckoo pb_cuckoo;
typedef struct pb_cbloom pb_cbloom;

/* A batch of keys (KeyT = str | bytes, hashes.py:10).
 *   data      packed symbols of all keys, back to back.
 *   sym_width 1: one byte per symbol (bytes keys, ASCII/latin-1 str keys);
 *             4: one little-endian u32 per symbol (str keys holding code points > 255:
 *                hashes.py:98 hashes ord(c), not UTF-8 bytes).
 *   offsets   NULL: every key is `stride` symbols long; else n+1 symbol offsets into data.
 *   on_device 0: data/offsets are host pointers (pinned or pageable; copied in chunks, the copy
 *                overlapped with the kernels); 1: device pointers on the context's device. */
typedef struct pb_keys {
    const void *data;
    const uint64_t *offsets;
    uint64_t n;
    uint32_t stride;
    uint32_t sym_width;
    int32_t on_device;
    int32_t reserved;
} pb_keys;

/* ---------------------------------------------------------------- library / context */
int pb_version(void);
const char *pb_last_error(void);
int pb_device_count(int *out);
/* stream: a cudaStream_t to run on (e.g. torch's), or NULL to let the context create its own. */
int pb_ctx_create(int device, void *stream, pb_ctx **out);
int pb_ctx_destroy(pb_ctx *ctx);
int pb_ctx_synchronize(pb_ctx *ctx);
int pb_ctx_stream(pb_ctx *ctx, void **out_stream);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int pb_ctx_launch_count(pb_ctx *ctx, uint64_t *out);
/* Device time per kernel name for the launches made since the previous call, while the "kernel_timing"
 * option is 1: text lines "name launches total_ms\n" (CUDA events on the context's stream). */
int pb_ctx_kernel_times(pb_ctx *ctx, char *buf, size_t cap);
/* tuning knobs: "bloom_insert_mode" (0 = auto, 1 = direct RED.OR, 2 = partition + L2-window apply),
 * "bloom_window_log2_bits", "stage_bytes" (staging budget for partitioned insert),
 * "bloom_min_chunks" (chunks a large batch is split into so pass 2 overlaps the next pass 1), "p2p_timeout_ms",
 * "h2d_chunk_keys" (host-buffer pipeline chunk), "cms_aggregate" (warp-combine equal keys: 0 never, 1 always, 2 = default: only when the hot-counter cache is off),
 * "cuckoo_serial" (1 = one-thread in-order cuckoo insert that reproduces the reference's append order),
 * "kernel_timing" (1 = bracket hot kernels with events, see pb_ctx_kernel_times). */
int pb_ctx_set_option(pb_ctx *ctx, const char *name, int64_t value);
int pb_ctx_get_option(pb_ctx *ctx, const char *name, int64_t *out);

/* pinned host memory and raw device buffers, so a ctypes-only caller can stage inputs */
int pb_host_alloc(size_t bytes, void **out);
int pb_host_free(void *p);
int pb_dev_alloc(pb_ctx *ctx, size_t bytes, void **out);
int pb_dev_free(pb_ctx *ctx, void *p);
int pb_memcpy_h2d(pb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int pb_memcpy_d2h(pb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
int pb_memset_dev(pb_ctx *ctx, void *dst_dev, int value, size_t bytes);
/* L2 flush helper for benchmarks: writes a scratch buffer larger than L2 */
int pb_flush_l2(pb_ctx *ctx);

/* ---------------------------------------------------------------- hashing (hashes.py:71-103) */
/* default_fnv_1a(key, depth) for every key: out[i*depth + s] = fnv_1a(key_i, seed = s). */
int pb_hash_keys(pb_ctx *ctx, const pb_keys *keys, uint32_t depth, uint64_t *out, int out_on_device);

/* fnv_1a(key, seed) (bits = 64, hashes.py:86-103) / fnv_1a_32(key, seed) (bits = 32, :106-122) with ANY seed: one
 * hash per key from the explicit start value (offset basis + 31*seed, reduced by the caller): out[i]. */
int pb_hash_keys_from(pb_ctx *ctx, const pb_keys *keys, uint64_t start_value, int bits, uint64_t *out, int out_on_device);

/* synthetic inputs of SURVEY.md 8(d) generated on the device (bench/test utility, not reference
 * code): uniform key i = LE64(sm64(seed+2i)) || LE64(sm64(seed+2i+1)); rank key = LE64(r)||LE64(sm64(r)) */
int pb_gen_uniform_keys(pb_ctx *ctx, uint64_t seed, uint64_t first, uint64_t n, void *out_dev);
int pb_gen_rank_keys(pb_ctx *ctx, const uint64_t *ranks_dev, uint64_t n, void *out_dev);
/* Zipf(a) ranks (config 3's stream) drawn on the device: rank i depends only on (seed, first + i) */
int pb_gen_zipf_ranks(pb_ctx *ctx, uint64_t seed, uint64_t first, uint64_t n, double a, uint64_t *out_ranks_dev);

/* ---------------------------------------------------------------- Bloom (blooms/bloom.py) */
/* state = ceil(num_bits/8) bytes, bit b lives in byte b/8 at mask 1<<(b%8) (bloom.py:247-249).
 * num_bits / k come from the Python side (bloom.py:463-483 stays host float math). */
int pb_bloom_create(pb_ctx *ctx, uint64_t num_bits, uint32_t k, pb_bloom **out);
int pb_bloom_destroy(pb_bloom *b);
int pb_bloom_clear(pb_bloom *b);                                         /* bloom.py:217-221 */
int pb_bloom_upload(pb_bloom *b, const uint8_t *bytes, uint64_t nbytes); /* bloom.py:548 */
int pb_bloom_download(pb_bloom *b, uint8_t *bytes, uint64_t nbytes);     /* bloom.py:299 */
int pb_bloom_device_ptr(pb_bloom *b, void **out_dev, uint64_t *out_nbytes);
/* BloomFilter.add for every key (bloom.py:234-250, hashes via hashes.py:71-103) */
int pb_bloom_add_keys(pb_bloom *b, const pb_keys *keys);
/* BloomFilter.check for every key (bloom.py:252-272): out[i] = 1/0 */
int pb_bloom_check_keys(pb_bloom *b, const pb_keys *keys, uint8_t *out, int out_on_device);
/* BloomFilter.add_alt / check_alt (bloom.py:241-250, :261-272) with caller-made hashes,
 * hashes[i*k + s] (a custom hash_function ran on the host); reduced % num_bits on the device. */
int pb_bloom_add_hashes(pb_bloom *b, const uint64_t *hashes, uint64_t n, int on_device);
int pb_bloom_check_hashes(pb_bloom *b, const uint64_t *hashes, uint64_t n, int on_device, uint8_t *out,
                          int out_on_device);
int pb_bloom_popcount(pb_bloom *b, uint64_t *out); /* bloom.py:552-557 */
/* BloomFilter.union (op 0) / intersection (op 1), bloom.py:371-428: dst = a op b for same-shaped filters on one device */
int pb_bloom_combine(pb_bloom *dst, pb_bloom *a, pb_bloom *b, int op);
/* jaccard_index (bloom.py:430-460): counts[0] = popcount(a | b), counts[1] = popcount(a & b) */
int pb_bloom_pair_popcounts(pb_bloom *a, pb_bloom *b, uint64_t *counts);
/* multi-GPU range sharding: this handle owns bits [lo, hi) of a num_bits-bit filter.
 * Keys are hashed against the GLOBAL num_bits; only indices inside [lo, hi) touch this shard. */
int pb_bloom_create_shard(pb_ctx *ctx, uint64_t num_bits, uint32_t k, uint64_t lo_bit, uint64_t hi_bit,
                          pb_bloom **out);
/* route: hash keys -> global bit index -> owner = idx / shard_bits; writes each owner's indices
 * (u64, global) contiguously into out_idx at out_offsets[owner]; counts[owner] returned.
 * All buffers are device pointers (they feed an NCCL all-to-all). */
int pb_bloom_route_keys(pb_ctx *ctx, const pb_keys *keys, uint64_t num_bits, uint32_t k, uint64_t shard_bits,
                        uint32_t n_shards, uint64_t *out_idx_dev, uint64_t slot_cap, uint64_t *counts_dev);
/* The fast multi-GPU insert: partition + exchange over NVLink peer memory + apply (no NCCL on the data path).
 * Every rank owns a mailbox (window sublists per source rank, three rotating buffers) that all peers map through
 * CUDA IPC.  pb_p2p_partition_send hashes a chunk of keys and bins the bit indices by GLOBAL window of
 * 2^window_log2 bits (window g belongs to rank g / windows_per_rank) as window-local u32 into a local staging laid
 * out [window][n_sub][sub_cap] -- one private sublist per CTA of the launch, so the kernel needs no global atomics --
 * then the copy engines push every destination's block into its mailbox and a publish kernel raises per-source
 * flags; pb_p2p_apply on the owner waits for all sources' flags, ORs the lists into its shard (window L2 resident)
 * and raises the flags that let the sources reuse the buffer.  Both calls are stream-ordered and never synchronise
 * the host; every rank must call them the same number of times.  Indices that do not fit their sublist go to
 * ovf_list_dev as global u64 indices (*ovf_count_dev counts them) for the caller to route exactly.
 * Flag waits are bounded by the "p2p_timeout_ms" option; pb_p2p_check reports who gave up.
 * pb_bloom_partition_layout: the (n_sub, sub_cap) a mailbox needs for chunks of up to n_keys keys per rank. */
typedef struct pb_p2p pb_p2p;
int pb_bloom_partition_layout(pb_ctx *ctx, uint64_t n_keys, uint32_t k, uint64_t num_bits, uint32_t window_log2,
                              uint32_t n_windows, uint32_t *n_sub, uint32_t *sub_cap);
int pb_p2p_create(pb_ctx *send_ctx, uint32_t world, uint32_t rank, uint32_t windows_per_rank, uint32_t n_sub,
                  uint32_t sub_cap, pb_p2p **out);
int pb_p2p_export(pb_p2p *p, uint8_t *handle_out /* 64 bytes */);
int pb_p2p_connect(pb_p2p *p, const uint8_t *handles /* world * 64 bytes */);
/* ranks living in one process (several shards driven by one host thread): connect by pointer, peers[r] = rank r */
int pb_p2p_connect_local(pb_p2p *p, pb_p2p *const *peers);
int pb_p2p_partition_send(pb_p2p *p, const pb_keys *keys, uint64_t num_bits, uint32_t k, uint32_t window_log2,
                          uint64_t *ovf_list_dev, uint64_t ovf_cap, uint64_t *ovf_count_dev);
int pb_p2p_apply(pb_p2p *p, pb_bloom *shard, uint32_t active_windows, uint32_t window_log2);
/* *aborted_by = 0 when healthy, else 1 + the rank whose flag wait timed out first (synchronous device read) */
int pb_p2p_check(pb_p2p *p, int *aborted_by);
int pb_p2p_destroy(pb_p2p *p);
/* multi-GPU query: global bit indices of device keys in key order, out_idx_dev[i*k + s] = h_s(key_i) % num_bits
 * (they travel to the owners of their bits; pb_bloom_test_bit_indices answers; pb_bloom_and_rows folds the k
 * answers of each key at the source) */
int pb_bloom_index_keys(pb_ctx *ctx, const pb_keys *keys, uint64_t num_bits, uint32_t k, uint64_t *out_idx_dev);
int pb_bloom_and_rows(pb_ctx *ctx, const uint8_t *bits_dev, uint64_t n, uint32_t k, uint8_t *out_dev);
/* apply global bit indices that fall into this shard's [lo, hi) (others are an error count) */
int pb_bloom_add_bit_indices(pb_bloom *b, const uint64_t *idx_dev, uint64_t n);
int pb_bloom_test_bit_indices(pb_bloom *b, const uint64_t *idx_dev, uint64_t n, uint8_t *out_dev);

/* Ordered "check, then add" batches: ExpandingBloomFilter / RotatingBloomFilter.add_alt (blooms/expandingbloom.py:159-169,
 * :320-330) add a key to the newest filter of the stack only if no filter finds it.  Rows = k bit indices per key
 * (pb_bloom_index_keys, or plugin hashes reduced mod num_bits), all on the device.
 * pb_bloom_novel_rows: novel_dev[i] = 1 iff adding the rows one at a time, in order, would add row i to this filter
 *   (rows with skip_dev[i] != 0 -- found in an older filter -- take no part; skip_dev may be NULL).  Read-only on the
 *   bits; at most 2^24 rows per call; keeps a 4-byte-per-bit table on the handle until pb_bloom_release_scratch.
 * pb_bloom_add_rows: bloom.py:241-250 for the rows with mask_dev[i] != 0 (NULL: all rows). */
int pb_bloom_novel_rows(pb_bloom *b, const uint64_t *idx_dev, uint64_t n, const uint8_t *skip_dev, uint8_t *novel_dev);
int pb_bloom_add_rows(pb_bloom *b, const uint64_t *idx_dev, uint64_t n, const uint8_t *mask_dev);
int pb_bloom_release_scratch(pb_bloom *b);
/* the table of pb_bloom_novel_rows moves on to the next filter of the stack (same geometry) instead of being freed */
int pb_bloom_move_scratch(pb_bloom *from, pb_bloom *to);
/* ExpandingBloomFilter.check_alt (:140-147): found_dev[i] = 1 iff one of the filters holds every bit of row i */
int pb_bloom_rows_in_any(pb_bloom *const *filters, uint32_t n_filters, const uint64_t *idx_dev, uint64_t n, uint8_t *found_dev);

/* ---------------------------------------------------------------- Counting Bloom (blooms/countingbloom.py) */
/* state = uint32[num_counters] (array('I'), bloom_length == number_bits, countingbloom.py:77-78);
 * counter of hash i = h_i % num_counters (:143). */
int pb_cbloom_create(pb_ctx *ctx, uint64_t num_counters, uint32_t k, pb_cbloom **out);
int pb_cbloom_destroy(pb_cbloom *b);
int pb_cbloom_clear(pb_cbloom *b);
int pb_cbloom_upload(pb_cbloom *b, const uint32_t *counts, uint64_t n);
int pb_cbloom_download(pb_cbloom *b, uint32_t *counts, uint64_t n);
int pb_cbloom_device_ptr(pb_cbloom *b, void **out_dev, uint64_t *out_count);
/* CountingBloomFilter.add for every key (:125-153): each (key, hash) pair adds num_els to its counter -- colliding
 * hashes of one key increment twice, as in the reference -- saturating at UINT32_MAX (:147-149). */
int pb_cbloom_add_keys(pb_cbloom *b, const pb_keys *keys, uint64_t num_els);
int pb_cbloom_add_hashes(pb_cbloom *b, const uint64_t *hashes, uint64_t n, int on_device, uint64_t num_els);
/* CountingBloomFilter.check (:155-174): out[i] = smallest of the key's counters */
int pb_cbloom_check_keys(pb_cbloom *b, const pb_keys *keys, uint32_t *out, int out_on_device);
int pb_cbloom_check_hashes(pb_cbloom *b, const uint64_t *hashes, uint64_t n, int on_device, uint32_t *out,
                           int out_on_device);
/* CountingBloomFilter.remove for every key IN ORDER (:176-208): a key removes min(num_els, its smallest counter)
 * from each of its counters, skips saturated counters and does nothing when its smallest counter is 0 or
 * UINT32_MAX.  *removed_total = sum of what the keys removed (what elements_added shrinks by, :207). */
int pb_cbloom_remove_keys(pb_cbloom *b, const pb_keys *keys, uint64_t num_els, uint64_t *removed_total);
int pb_cbloom_remove_hashes(pb_cbloom *b, const uint64_t *hashes, uint64_t n, int on_device, uint64_t num_els,
                            uint64_t *removed_total);
/* out4[0] = non-zero counters (_cnt_number_bits_set, :328-330), [1] = sum of all counters, [2] = largest counter,
 * [3] = first index holding it (__str__, :100-123) */
int pb_cbloom_stats(pb_cbloom *b, uint64_t *out4);
/* union (op 0, :300-326) / intersection (op 1, :210-243): dst[i] = a[i] + b[i] (saturating), for the intersection
 * only where both are non-zero */
int pb_cbloom_combine(pb_cbloom *dst, pb_cbloom *a, pb_cbloom *b, int op);
/* jaccard_index (:245-272): counts[0] = counters non-zero in a or b, counts[1] = non-zero in both */
int pb_cbloom_pair_counts(pb_cbloom *a, pb_cbloom *b, uint64_t *counts);

/* ---------------------------------------------------------------- Count-Min (countminsketch.py) */
/* state = int32[depth][width] row-major; bin of row i = (h_i % width) + i*width (:275). */
int pb_cms_create(pb_ctx *ctx, uint32_t width, uint32_t depth, pb_cms **out);
int pb_cms_destroy(pb_cms *c);
int pb_cms_clear(pb_cms *c); /* :240-244 */
int pb_cms_upload(pb_cms *c, const int32_t *bins, uint64_t count);
int pb_cms_download(pb_cms *c, int32_t *bins, uint64_t count);
int pb_cms_device_ptr(pb_cms *c, void **out_dev, uint64_t *out_count);
/* CountMinSketch.add for every key (:257-288): bins saturate at INT32_MAX (:280-282);
 * num_els == NULL -> every key adds scalar_num_els.  *elements_added is updated with the
 * saturating sum (:285-287).  num_els follows keys->on_device. */
int pb_cms_add_keys(pb_cms *c, const pb_keys *keys, const int64_t *num_els, int64_t scalar_num_els,
                    int64_t *elements_added_inout);
/* CountMinSketch.check (:323-340) with query_type 0 = min (:430-432), 1 = mean (:434-436),
 * 2 = mean-min (:438-453; needs elements_added). */
int pb_cms_check_keys(pb_cms *c, const pb_keys *keys, int query_type, int64_t elements_added, int64_t *out,
                      int out_on_device);
/* add_alt / check_alt (:267-288, :332-340) with caller-made hashes[i*depth + r] */
int pb_cms_add_hashes(pb_cms *c, const uint64_t *hashes, uint64_t n, int on_device, const int64_t *num_els,
                      int64_t scalar_num_els, int64_t *elements_added_inout);
int pb_cms_check_hashes(pb_cms *c, const uint64_t *hashes, uint64_t n, int on_device, int query_type,
                        int64_t elements_added, int64_t *out, int out_on_device);
/* CountMinSketch.join (:380-391) against a second table already on this device (multi-GPU merge) */
int pb_cms_join_buffer(pb_cms *c, const int32_t *other_dev, uint64_t count);
/* multi-GPU merge as ONE all-reduce: the table widened to int64, and a table loaded from int64 sums with the
 * reference's saturation (for non-negative tables a chain of saturating joins equals min(sum, INT32_MAX)) */
int pb_cms_widen(pb_cms *c, int64_t *out_dev, uint64_t count);
int pb_cms_load_sums(pb_cms *c, const int64_t *sums_dev, uint64_t count);

/* ---------------------------------------------------------------- Cuckoo (cuckoo/cuckoo.py) */
/* state = u32[capacity][bucket_size], 0 = empty slot; fingerprint 0 (legal, utilities.py:35) is
 * kept as a side flag because it is indistinguishable from "empty" in the reference's own export
 * format (cuckoo.py:346, :429).  fp = low fp_bits of fnv_1a(key) (:499-500);
 * idx_1 = fp % capacity; idx_2 = fnv_1a(str(fp)) % capacity (:488-489). */
int pb_cuckoo_create(pb_ctx *ctx, uint64_t capacity, uint32_t bucket_size, uint32_t max_swaps, uint32_t fp_bits,
                     uint64_t rng_seed, pb_cuckoo **out);
int pb_cuckoo_destroy(pb_cuckoo *c);
int pb_cuckoo_clear(pb_cuckoo *c);
/* CuckooFilter.add for every key (:291-304): fingerprints already present are skipped (:300-302),
 * the rest are placed (idx_1, then idx_2, then an eviction walk of <= max_swaps, :361-392).
 * *n_added = fingerprints newly stored (what elements_added grows by).  Fingerprints left homeless
 * after max_swaps are returned in failed_fps (up to failed_cap) with *n_failed their total count
 * and the call returns PB_ERR_CUCKOO_FULL so the caller can expand or raise (:508-516). */
int pb_cuckoo_add_keys(pb_cuckoo *c, const pb_keys *keys, uint64_t *n_added, uint64_t *n_failed,
                       uint32_t *failed_fps, uint64_t failed_cap);
/* same, starting from fingerprints (used by expand, :467-481) */
int pb_cuckoo_add_fingerprints(pb_cuckoo *c, const uint32_t *fps, uint64_t n, int on_device, uint64_t *n_added,
                               uint64_t *n_failed, uint32_t *failed_fps, uint64_t failed_cap);
/* CuckooFilter.check (:306-315) */
int pb_cuckoo_check_keys(pb_cuckoo *c, const pb_keys *keys, uint8_t *out, int out_on_device);
/* _check_if_present (:440-446) from fingerprints (custom hash_function ran on the host) */
int pb_cuckoo_check_fingerprints(pb_cuckoo *c, const uint32_t *fps, uint64_t n, int on_device, uint8_t *out,
                                 int out_on_device);
/* CuckooFilter.remove (:317-330) for every key: out[i] = 1 when a stored copy of the key's fingerprint was cleared
 * (idx_1 first, then idx_2); *n_removed is what elements_added shrinks by.  Equal keys inside one batch: exactly one
 * of them wins. */
int pb_cuckoo_remove_keys(pb_cuckoo *c, const pb_keys *keys, uint8_t *out, int out_on_device, uint64_t *n_removed);
/* Pre-indexed filters -- a custom hash_function decides idx_2 = hash_function(str(fp)) % capacity (:489), which the
 * device cannot evaluate: the caller passes idx_2 with every fingerprint (host arrays), the table keeps the idx_2
 * of every stored fingerprint for the eviction walk (:383-385).  Inserts run in key order on one thread (the
 * reference's append order); failed_* return the homeless (fingerprint, idx_2) pairs with PB_ERR_CUCKOO_FULL. */
int pb_cuckoo_add_indexed(pb_cuckoo *c, const uint32_t *fps, const uint64_t *idx2, uint64_t n, uint64_t *n_added,
                          uint64_t *n_failed, uint32_t *failed_fps, uint64_t *failed_idx2, uint64_t failed_cap);
int pb_cuckoo_check_indexed(pb_cuckoo *c, const uint32_t *fps, const uint64_t *idx2, uint64_t n, uint8_t *out);
int pb_cuckoo_remove_indexed(pb_cuckoo *c, const uint32_t *fps, const uint64_t *idx2, uint64_t n, uint8_t *out,
                             uint64_t *n_removed);
/* idx_2 per slot for an uploaded table (a file written by the reference with a custom hash), and an empty table
 * of a new capacity (pre-indexed filters expand by re-inserting from the host, :455-481) */
int pb_cuckoo_set_alt(pb_cuckoo *c, const uint64_t *idx2_per_slot, uint64_t count);
int pb_cuckoo_resize(pb_cuckoo *c, uint64_t new_capacity);
/* _generate_fingerprint_info (:492-506): per key fp, idx_1, idx_2 */
int pb_cuckoo_fingerprint_info(pb_cuckoo *c, const pb_keys *keys, uint32_t *fp, uint64_t *idx1, uint64_t *idx2,
                               int out_on_device);
int pb_cuckoo_count(pb_cuckoo *c, uint64_t *out); /* stored fingerprints incl. the fp-0 flag */
int pb_cuckoo_download(pb_cuckoo *c, uint32_t *slots, uint64_t count, int *has_zero_fp); /* :340-347 */
int pb_cuckoo_upload(pb_cuckoo *c, const uint32_t *slots, uint64_t count, int has_zero_fp);
int pb_cuckoo_device_ptr(pb_cuckoo *c, void **out_dev, uint64_t *out_count);
int pb_cuckoo_capacity(pb_cuckoo *c, uint64_t *out);
/* _expand_logic (:455-481): the table becomes new_capacity buckets and every stored fingerprint is
 * re-inserted on the device.  Fingerprints that find no home are returned like pb_cuckoo_add_keys does
 * (PB_ERR_CUCKOO_FULL; the reference raises "The CuckooFilter failed to expand", :463-465). */
int pb_cuckoo_expand(pb_cuckoo *c, uint64_t new_capacity, uint64_t *n_failed, uint32_t *failed_fps, uint64_t failed_cap);

/* ---------------------------------------------------------------- Counting Cuckoo (cuckoo/countingcuckoo.py) */
/* A CountingCuckooFilter stores every fingerprint once, next to a count (countingcuckoo.py:156-173).  Here the
 * fingerprints stay in the pb_cuckoo table (all entry points above apply unchanged) and the counts live in a
 * fingerprint -> count map on the same handle:
 *   add    (:156-173)  = pb_cuckoo_add_keys (stores the new fingerprints) + pb_cuckoo_counts_add_keys (+1 per key)
 *   check  (:175-191)  = pb_cuckoo_counts_get_keys -> uint32 count per key, 0 when absent
 *   remove (:193-210)  = pb_cuckoo_counts_remove_keys: out[i] = 1 when a count was decremented; a fingerprint whose count
 *                        reaches 0 leaves the table.  n_removed = decrements, n_bins_removed = fingerprints that left.
 *   expand (:212-214, :305-316) = pb_cuckoo_expand, then pb_cuckoo_counts_enable again (re-fits the map).
 * The *_fingerprints variants take host arrays (custom hash_function path, load/export of the (fp, count) bins :216-228,
 * :286-303); pb_cuckoo_counts_set stores explicit counts (vals NULL: 0 = forget the fingerprint). */
int pb_cuckoo_counts_enable(pb_cuckoo *c);
int pb_cuckoo_counts_add_keys(pb_cuckoo *c, const pb_keys *keys);
int pb_cuckoo_counts_add_fingerprints(pb_cuckoo *c, const uint32_t *fps, const uint32_t *amounts, uint64_t n, int on_device);
int pb_cuckoo_counts_get_keys(pb_cuckoo *c, const pb_keys *keys, uint32_t *out, int out_on_device);
int pb_cuckoo_counts_get_fingerprints(pb_cuckoo *c, const uint32_t *fps, uint64_t n, uint32_t *out);
int pb_cuckoo_counts_set(pb_cuckoo *c, const uint32_t *fps, const uint32_t *vals, uint64_t n);
int pb_cuckoo_counts_remove_keys(pb_cuckoo *c, const pb_keys *keys, uint8_t *out, int out_on_device, uint64_t *n_removed,
                                 uint64_t *n_bins_removed);
int pb_cuckoo_counts_remove_fingerprints(pb_cuckoo *c, const uint32_t *fps, const uint64_t *i2, uint64_t n, uint8_t *out,
                                         uint64_t *n_removed, uint64_t *n_bins_removed);

/* ---------------------------------------------------------------- roofline micro-benchmarks */
/* n random RED.OR.b32 / atomicAdd.s32 over `words` 32-bit words with pre-generated uniform
 * indices and no hashing: the empirical random-atomic ceiling SURVEY 8(d) asks for. ms = device time. */
int pb_microbench_random_atomic(pb_ctx *ctx, uint64_t words, uint64_t n,
                                int op /*0=RED.OR 1=RED.ADD 2=random 32-bit load 3=streaming copy of `words`*/, int reps,
                                float *ms_best);

#ifdef __cplusplus
}
#endif
#endif /* PB200_H */
