// pb_common.cuh -- context, error plumbing and launch helpers shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/pb200.h"

namespace pb {

constexpr int kNumSMsB200 = 148;

void set_error(const char *fmt, ...);

#define PB_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            pb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return _e == cudaErrorMemoryAllocation ? PB_ERR_OOM : PB_ERR_CUDA;                     \
        }                                                                                          \
    } while (0)

#define PB_REQUIRE(cond, ...)                \
    do {                                     \
        if (!(cond)) {                       \
            pb::set_error(__VA_ARGS__);      \
            return PB_ERR_BAD_ARG;           \
        }                                    \
    } while (0)

#define PB_TRY(expr)                \
    do {                            \
        int _s = (expr);            \
        if (_s != PB_OK) return _s; \
    } while (0)

}  // namespace pb

// Scratch device buffer that only ever grows (reused across calls: no cudaMalloc on the hot path).
struct pb_scratch {
    void *p = nullptr;
    size_t cap = 0;
};

// one timed kernel launch (only while the "kernel_timing" option is on): events bracket the launch on
// the compute stream; pb_ctx_kernel_times() resolves them
struct pb_timed_launch {
    const char *name;
    cudaEvent_t e0, e1;
};

struct pb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;   // compute stream (all kernels)
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;  // H2D staging for host-buffer batches
    cudaStream_t aux_stream = nullptr;   // pass 2 of the partitioned Bloom insert (overlaps pass 1 of the next chunk)
    cudaEvent_t ev_part[2] = {nullptr, nullptr};   // pass 1 of the chunk using staging half i is done
    cudaEvent_t ev_apply[2] = {nullptr, nullptr};  // pass 2 ... is done (the half may be refilled)
    bool apply_pending[2] = {false, false};
    int num_sms = pb::kNumSMsB200;
    size_t l2_bytes = 0;
    uint64_t launches = 0;
    // options
    int64_t bloom_insert_mode = 0;       // 0 auto, 1 direct, 2 partitioned
    int64_t bloom_window_log2_bits = 27; // 2^27 bits = 16 MiB of bitmap per L2 window (r1 sweeps: best with overlap)
    int64_t bloom_apply_cpw_per_sm = 4;  // apply pass: CTAs per window = this x SMs (at most ~2 windows in flight, so the
                                         // windows being updated stay L2 resident; 4 x 32 MiB at once thrashed: r1 sweep)
    int64_t bloom_part_tile = 0;         // keys per pass-1 tile: 0 auto (512 beyond 112 windows), 256, 512
    int64_t bloom_overlap = 1;           // run pass 2 of chunk i on aux_stream while pass 1 of chunk i+1 runs
    int64_t bloom_part_ctas_per_sm = 0;  // pass 1: resident 256-thread CTAs per SM; 0 = auto (3 beside pass 2, else what registers allow)
    int64_t bloom_check_mode = 0;        // query of device keys: 0 auto (partitioned when most sampled keys are members), 1 direct, 2 partitioned
    int64_t bloom_min_chunks = 8;        // overlapped partitioned insert: split a batch into at least this many chunks
    int64_t stage_bytes = 8ll << 30;     // staging budget for partitioned insert
    int64_t h2d_chunk_keys = 1ll << 24;  // keys per H2D pipeline chunk
    int64_t cms_aggregate = 2;           // warp-aggregate equal keys before the atomics: 0 never, 1 always, 2 auto (only without the hot cache)
    int64_t cms_hot_cache = 1;           // per-CTA shared-memory write-back cache for hot counters (safe path)
    int64_t cuckoo_serial = 0;           // 1: one-thread in-order cuckoo insert (reference append order)
    int64_t p2p_copy_lanes = 4;          // multi-GPU exchange: streams (copy engines) the pushes of one chunk are spread over
    int64_t p2p_timeout_ms = 20000;      // multi-GPU flag waits give up after this long (pb_p2p_check reports it)
    int64_t kernel_timing = 0;           // 1: bracket the hot kernels with CUDA events (bench roofline)
    std::vector<pb_timed_launch> timed;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t pending_e0 = nullptr;
    // reusable buffers
    pb_scratch key_stage[2];   // device staging of host key data (double buffered)
    pb_scratch off_stage[2];   // device staging of host offsets
    pb_scratch aux_stage[2];   // per-key side input (num_els / hashes)
    pb_scratch out_stage[2];   // per-key device output before D2H
    pb_scratch part_stage;     // partitioned-insert index staging
    pb_scratch part_cursors;   // bucket cursors
    pb_scratch small;          // counters and tiny results
    pb_scratch flush;          // L2 flush buffer
    pb_scratch claim_set;      // cuckoo in-batch dedupe set (small batches)
    pb_scratch claim_bitmap;   // cuckoo in-batch dedupe bitmap, 2^fp_bits bits (large batches); all zero between calls
    void *pinned[2] = {nullptr, nullptr};  // pinned bounce buffers for pageable host memory
    size_t pinned_cap[2] = {0, 0};
    void *pinned_small = nullptr;          // 4 KiB pinned result area
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_done[2] = {nullptr, nullptr};
};

namespace pb {

int scratch_reserve(pb_ctx *ctx, pb_scratch &s, size_t bytes);
void scratch_release(pb_scratch &s);

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// grid sizing: persistent-ish grids in multiples of the SM count
inline int grid_for(const pb_ctx *ctx, uint64_t work_items, int block, int ctas_per_sm) {
    uint64_t need = (work_items + (uint64_t)block - 1) / (uint64_t)block;
    uint64_t cap = (uint64_t)ctx->num_sms * (uint64_t)ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

inline cudaEvent_t timing_event(pb_ctx *ctx) {
    cudaEvent_t e = nullptr;
    if (!ctx->event_pool.empty()) {
        e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}

// call right before the <<<>>> of a kernel that should show up in pb_ctx_kernel_times()
inline void launch_begin(pb_ctx *ctx, cudaStream_t on = nullptr) {
    if (!ctx->kernel_timing || ctx->pending_e0) return;
    ctx->pending_e0 = timing_event(ctx);
    cudaEventRecord(ctx->pending_e0, on ? on : ctx->stream);
}

inline int check_launch(pb_ctx *ctx, const char *what, cudaStream_t on = nullptr) {
    ctx->launches++;
    if (ctx->pending_e0) {
        pb_timed_launch t{what, ctx->pending_e0, timing_event(ctx)};
        cudaEventRecord(t.e1, on ? on : ctx->stream);
        ctx->timed.push_back(t);
        ctx->pending_e0 = nullptr;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return PB_ERR_CUDA;
    }
    return PB_OK;
}

// ---- host-batch pipeline -------------------------------------------------------------------
// A key batch as the kernels see it (device pointers only).
struct DevKeys {
    const uint8_t *data;      // symbols (byte address even when sym_width == 4)
    const uint64_t *offsets;  // nullptr => fixed stride
    uint64_t n;
    uint32_t stride;
    uint32_t sym_width;
    uint64_t base_symbol;  // offsets are relative to this symbol index of `data` (chunked host batches)
    uint64_t total_bytes;  // bytes readable from data (for bulk-copy clamping)
};

// Calls fn(dev_keys, first_key_index, slot) for successive chunks of `keys`.  Host batches are copied
// chunk by chunk (pinned memory directly, pageable through a pinned bounce buffer) on the copy stream
// and handed to fn on the compute stream with event ordering, double buffered.
// fn enqueues work on ctx->stream only.  After the last chunk the compute stream is synchronized when
// the batch came from the host.
typedef int (*chunk_fn)(pb_ctx *ctx, const DevKeys &dk, uint64_t first, int slot, void *user);
int for_each_chunk(pb_ctx *ctx, const pb_keys *keys, chunk_fn fn, void *user, uint64_t max_chunk_keys = 0);

int validate_keys(const pb_keys *keys);

// default_fnv_1a rows of device-resident keys (pb_ctx.cu): out[i*depth + s] = fnv_1a(key_i, seed s)
int hash_dev_keys(pb_ctx *ctx, const DevKeys &dk, uint32_t depth, uint64_t *out);

}  // namespace pb
