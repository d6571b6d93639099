// pb_p2p.cu -- the multi-GPU Bloom insert: partition + exchange over NVLink peer memory + apply, no NCCL on the
// data path.
//
// Every rank owns a "mailbox" in its own HBM (cudaMalloc'd, exported with CUDA IPC and mapped by every peer):
//     recv_stage[3][n_src][wps][n_sub][sub_cap]  window sublists, one block per source rank; three buffers rotate
//     recv_cnt  [3][n_src][wps][n_sub]           entries per sublist
//     data_flag [3][n_src]    written by source s: "chunk seq of mine is complete in your buffer h"
//     done_flag [3][n_dst]    written by destination d into the SOURCE's mailbox: "I applied your chunk seq"
//     abort_flag              raised by any rank whose wait timed out: everybody stops waiting
// Three buffers, not two: with two, pass 1 of chunk c+2 would have to wait for pass 2 of chunk c on every
// destination (pass 1 + copy + pass 2 > 2 x pass 1), a bubble in every second chunk and no slack for rank skew.
// Pass 1 (bloom_part4) hashes a chunk of local keys and bins the bit indices by GLOBAL window into a local staging
// laid out [window][n_sub][sub_cap]; the windows of one destination are contiguous, so the copy engines push each
// destination's block into its mailbox (cudaMemcpyAsync peer-to-peer over NVLink) while the SMs go on with pass 1
// of the next chunk and pass 2 of the previous one.  A tiny publish kernel then writes the sublist lengths and
// raises the data flags (release, system scope).  Pass 2 on the owner waits for the flags of all sources
// (acquire), ORs the lists into its shard with the window L2 resident, and raises the done flags so the sources
// may reuse that buffer three chunks later.  No host synchronisation; ordering is flags in peer memory.
// Waits are bounded: a flag that does not arrive within "p2p_timeout_ms" raises the abort flag in every mailbox
// and the error is reported by pb_p2p_check (a dead peer cannot hang the GPU).
#include <algorithm>
#include <new>

#include "pb_bloom_part.cuh"
#include "pb_common.cuh"
#include "pb_hash.cuh"

using namespace pb;

struct pb_bloom;  // pb_bloom.cu
namespace pb {
uint32_t *bloom_words(pb_bloom *b);
pb_ctx *bloom_ctx(pb_bloom *b);
}  // namespace pb

constexpr int kBufs = 3;
constexpr int kMaxRanks = 16;
constexpr int kCopyLanes = 4;

struct pb_p2p {
    pb_ctx *send_ctx = nullptr;
    uint32_t world = 0, rank = 0, wps = 0, n_sub = 0, sub_cap = 0;
    uint8_t *local = nullptr;             // this rank's mailbox
    uint8_t *peer[kMaxRanks] = {nullptr};  // mapped mailboxes (peer[rank] == local)
    bool opened[kMaxRanks] = {false};     // mapped through CUDA IPC (to be closed)
    size_t stage_bytes = 0, cnt_bytes = 0, total_bytes = 0;
    uint32_t *scnt[kBufs] = {nullptr, nullptr, nullptr};    // counts of the chunk being partitioned [world*wps][n_sub]
    uint32_t *lstage[kBufs] = {nullptr, nullptr, nullptr};  // local staging [world*wps][n_sub][sub_cap], allocated on first use
    uint64_t send_seq = 0, apply_seq = 0;
    cudaStream_t copy_stream = nullptr;
    // the pushes of one chunk are spread over several streams so that several copy engines work at once (one
    // cudaMemcpyAsync at a time does not fill NVLink: r2, N = 8 was bound by the serialised pushes)
    cudaStream_t copy_lane[kCopyLanes] = {nullptr};
    cudaEvent_t ev_lane_go = nullptr, ev_lane_done[kCopyLanes] = {nullptr};
    cudaEvent_t ev_part[kBufs] = {nullptr, nullptr, nullptr}, ev_copy[kBufs] = {nullptr, nullptr, nullptr};
};

namespace pb {

static size_t off_stage(const pb_p2p *p, int h) { return (size_t)h * p->stage_bytes; }
static size_t off_cnt(const pb_p2p *p, int h) { return kBufs * p->stage_bytes + (size_t)h * p->cnt_bytes; }
static size_t off_flags(const pb_p2p *p) { return kBufs * p->stage_bytes + kBufs * p->cnt_bytes; }
static size_t off_data_flag(const pb_p2p *p, int h) { return off_flags(p) + (size_t)h * kMaxRanks * 8; }
static size_t off_done_flag(const pb_p2p *p, int h) { return off_flags(p) + (size_t)(kBufs + h) * kMaxRanks * 8; }
static size_t off_abort_flag(const pb_p2p *p) { return off_flags(p) + (size_t)2 * kBufs * kMaxRanks * 8; }
constexpr size_t kFlagBytes = (size_t)2 * kBufs * kMaxRanks * 8 + 64;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct P2PPeers {
    uint8_t *base[kMaxRanks];
};

// lane r spins until flags[r] >= want (flags are raised by other GPUs through NVLink).  The wait is bounded: after
// timeout_ns -- or as soon as somebody else gave up -- the lane raises the abort flag in every mailbox and returns,
// so a peer that died or skipped a collective call costs an error (pb_p2p_check), never a hung GPU.
__global__ void p2p_wait_flags(const unsigned long long *flags, uint32_t n, unsigned long long want, P2PPeers peers, uint32_t world,
                               size_t abort_off, uint32_t rank, unsigned long long timeout_ns) {
    const uint32_t r = threadIdx.x;
    if (r >= n) return;
    const unsigned long long *my_abort = reinterpret_cast<const unsigned long long *>(peers.base[rank] + abort_off);
    const unsigned long long t0 = global_timer_ns();
    uint32_t spins = 0;
    while (ld_acquire_sys(flags + r) < want) {
        __nanosleep(200);
        if ((++spins & 255u) == 0) {
            const bool timed_out = global_timer_ns() - t0 > timeout_ns;
            if (timed_out || ld_acquire_sys(my_abort) != 0) {
                for (uint32_t d = 0; d < world; ++d)
                    st_release_sys(reinterpret_cast<unsigned long long *>(peers.base[d] + abort_off), 1ull + rank);
                return;
            }
        }
    }
}

// block d: hand rank d the lengths of the sublists this rank just wrote into its mailbox, then raise the flag
__global__ void p2p_publish(P2PPeers peers, const uint32_t *__restrict__ scnt, uint32_t per_dest, uint32_t rank, size_t cnt_off,
                            size_t flag_off, unsigned long long seq) {
    const uint32_t d = blockIdx.x;
    uint32_t *rc = reinterpret_cast<uint32_t *>(peers.base[d] + cnt_off) + (size_t)rank * per_dest;
    for (uint32_t j = threadIdx.x; j < per_dest; j += blockDim.x) rc[j] = scnt[(size_t)d * per_dest + j];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(reinterpret_cast<unsigned long long *>(peers.base[d] + flag_off) + rank, seq);
}

// lane s: tell source s that its chunk `seq` has been applied here (its buffer may be refilled)
__global__ void p2p_done(P2PPeers peers, uint32_t n, uint32_t rank, size_t flag_off, unsigned long long seq) {
    const uint32_t s = threadIdx.x;
    __threadfence_system();
    if (s < n) st_release_sys(reinterpret_cast<unsigned long long *>(peers.base[s] + flag_off) + rank, seq);
}

static P2PPeers peers_of(const pb_p2p *p) {
    P2PPeers q;
    for (int i = 0; i < kMaxRanks; ++i) q.base[i] = p->peer[i];
    return q;
}

static int launch_wait(pb_p2p *p, pb_ctx *ctx, cudaStream_t st, size_t flag_off, unsigned long long want) {
    const unsigned long long timeout_ns = (unsigned long long)std::max<int64_t>(ctx->p2p_timeout_ms, 1) * 1000000ull;
    p2p_wait_flags<<<1, 32, 0, st>>>(reinterpret_cast<const unsigned long long *>(p->local + flag_off), p->world, want, peers_of(p),
                                     p->world, off_abort_flag(p), p->rank, timeout_ns);
    return check_launch(ctx, "p2p_wait", st);
}

}  // namespace pb

extern "C" {

int pb_p2p_create(pb_ctx *send_ctx, uint32_t world, uint32_t rank, uint32_t windows_per_rank, uint32_t n_sub, uint32_t sub_cap,
                  pb_p2p **out) {
    PB_REQUIRE(send_ctx && out, "NULL argument");
    PB_REQUIRE(world >= 1 && world <= (uint32_t)kMaxRanks && rank < world, "world must be 1..16 and rank < world");
    PB_REQUIRE(windows_per_rank >= 1 && (uint64_t)windows_per_rank * world <= (uint64_t)kMaxWindows2, "too many windows");
    PB_REQUIRE(n_sub >= 1 && (sub_cap & 3u) == 0 && sub_cap >= 4, "bad sublist layout");
    PB_REQUIRE((uint64_t)sub_cap * n_sub * windows_per_rank * world <= 0xFFFFFFF0ull, "staging too large for 32-bit entry numbers");
    DeviceGuard g(send_ctx->device);
    pb_p2p *p = new (std::nothrow) pb_p2p();
    if (!p) return PB_ERR_OOM;
    p->send_ctx = send_ctx;
    p->world = world;
    p->rank = rank;
    p->wps = windows_per_rank;
    p->n_sub = n_sub;
    p->sub_cap = sub_cap;
    const size_t lists = (size_t)world * windows_per_rank * n_sub;
    p->stage_bytes = (lists * sub_cap * 4 + 255) & ~(size_t)255;
    p->cnt_bytes = (lists * 4 + 255) & ~(size_t)255;
    p->total_bytes = kBufs * p->stage_bytes + kBufs * p->cnt_bytes + kFlagBytes;
    cudaError_t e = cudaMalloc(&p->local, p->total_bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc of the %zu-byte P2P mailbox failed: %s", p->total_bytes, cudaGetErrorString(e));
        delete p;
        return PB_ERR_OOM;
    }
    for (int h = 0; h < kBufs; ++h) {
        e = cudaMalloc(&p->scnt[h], lists * 4);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaFree(p->local);
            delete p;
            set_error("cudaMalloc failed");
            return PB_ERR_OOM;
        }
        PB_CUDA(cudaEventCreateWithFlags(&p->ev_part[h], cudaEventDisableTiming));
        PB_CUDA(cudaEventCreateWithFlags(&p->ev_copy[h], cudaEventDisableTiming));
    }
    PB_CUDA(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    PB_CUDA(cudaEventCreateWithFlags(&p->ev_lane_go, cudaEventDisableTiming));
    for (int l = 0; l < kCopyLanes; ++l) {
        PB_CUDA(cudaStreamCreateWithFlags(&p->copy_lane[l], cudaStreamNonBlocking));
        PB_CUDA(cudaEventCreateWithFlags(&p->ev_lane_done[l], cudaEventDisableTiming));
    }
    // flags and counts start at zero; the lists need no initialisation
    PB_CUDA(cudaMemset(p->local + kBufs * p->stage_bytes, 0, kBufs * p->cnt_bytes + kFlagBytes));
    p->peer[rank] = p->local;
    *out = p;
    return PB_OK;
}

int pb_p2p_export(pb_p2p *p, uint8_t *handle_out) {
    PB_REQUIRE(p && handle_out, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard g(p->send_ctx->device);
    cudaIpcMemHandle_t h;
    PB_CUDA(cudaIpcGetMemHandle(&h, p->local));
    memcpy(handle_out, &h, 64);
    return PB_OK;
}

int pb_p2p_connect(pb_p2p *p, const uint8_t *handles) {
    PB_REQUIRE(p && handles, "NULL argument");
    DeviceGuard g(p->send_ctx->device);
    for (uint32_t r = 0; r < p->world; ++r) {
        if (r == p->rank || p->peer[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("cudaIpcOpenMemHandle for rank %u failed: %s", r, cudaGetErrorString(e));
            return PB_ERR_CUDA;
        }
        p->peer[r] = (uint8_t *)ptr;
        p->opened[r] = true;
    }
    return PB_OK;
}

// Ranks that live in ONE process (several shards driven by one host thread, e.g. a test that runs the whole
// exchange on a single GPU): connect by pointer instead of CUDA IPC.  peers[r] is rank r's pb_p2p.
int pb_p2p_connect_local(pb_p2p *p, pb_p2p *const *peers) {
    PB_REQUIRE(p && peers, "NULL argument");
    for (uint32_t r = 0; r < p->world; ++r) {
        PB_REQUIRE(peers[r] != nullptr, "peer %u is NULL", r);
        PB_REQUIRE(peers[r]->world == p->world && peers[r]->rank == r && peers[r]->wps == p->wps && peers[r]->n_sub == p->n_sub &&
                       peers[r]->sub_cap == p->sub_cap,
                   "peer %u has a different layout", r);
        if (r == p->rank) continue;
        if (peers[r]->send_ctx->device != p->send_ctx->device) {
            DeviceGuard g(p->send_ctx->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(peers[r]->send_ctx->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
                set_error("cudaDeviceEnablePeerAccess(%d) failed: %s", peers[r]->send_ctx->device, cudaGetErrorString(e));
                return PB_ERR_CUDA;
            }
            cudaGetLastError();
        }
        p->peer[r] = peers[r]->local;
    }
    return PB_OK;
}

// 0 = healthy; otherwise 1 + the rank whose wait gave up first (timeout or a dead peer).  Synchronizes both streams.
int pb_p2p_check(pb_p2p *p, int *aborted_by) {
    PB_REQUIRE(p && aborted_by, "NULL argument");
    DeviceGuard g(p->send_ctx->device);
    unsigned long long v = 0;
    PB_CUDA(cudaMemcpy(&v, p->local + off_abort_flag(p), 8, cudaMemcpyDeviceToHost));
    *aborted_by = (int)v;
    return PB_OK;
}

int pb_p2p_destroy(pb_p2p *p) {
    if (!p) return PB_OK;
    DeviceGuard g(p->send_ctx->device);
    cudaDeviceSynchronize();
    for (uint32_t r = 0; r < p->world; ++r)
        if (p->opened[r]) cudaIpcCloseMemHandle(p->peer[r]);
    for (int h = 0; h < kBufs; ++h) {
        cudaFree(p->scnt[h]);
        if (p->lstage[h]) cudaFree(p->lstage[h]);
        if (p->ev_part[h]) cudaEventDestroy(p->ev_part[h]);
        if (p->ev_copy[h]) cudaEventDestroy(p->ev_copy[h]);
    }
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    if (p->ev_lane_go) cudaEventDestroy(p->ev_lane_go);
    for (int l = 0; l < kCopyLanes; ++l) {
        if (p->copy_lane[l]) cudaStreamDestroy(p->copy_lane[l]);
        if (p->ev_lane_done[l]) cudaEventDestroy(p->ev_lane_done[l]);
    }
    cudaFree(p->local);
    delete p;
    return PB_OK;
}

// Pass 1 + exchange of one chunk of this rank's keys (stream of the send context).
int pb_p2p_partition_send(pb_p2p *p, const pb_keys *keys, uint64_t num_bits, uint32_t k, uint32_t window_log2,
                          uint64_t *ovf_list_dev, uint64_t ovf_cap, uint64_t *ovf_count_dev) {
    PB_REQUIRE(p && keys && ovf_list_dev && ovf_count_dev, "NULL argument");
    PB_REQUIRE(keys->on_device, "pb_p2p_partition_send takes device keys");
    PB_REQUIRE(keys->offsets == nullptr && keys->sym_width == 1 && keys->stride == 16 && ((uintptr_t)keys->data & 15u) == 0,
               "partitioned routing takes fixed 16-byte keys");
    PB_REQUIRE(k >= 1 && k <= kMaxPartK, "k must be in 1..%u", kMaxPartK);
    PB_REQUIRE(window_log2 >= 5 && window_log2 <= 31, "window_log2 must be in 5..31");
    const uint32_t W = p->world * p->wps;
    PB_REQUIRE(((uint64_t)W << window_log2) >= num_bits, "windows do not cover the filter");
    for (uint32_t r = 0; r < p->world; ++r) PB_REQUIRE(p->peer[r] != nullptr, "rank %u is not connected (pb_p2p_connect)", r);
    pb_ctx *ctx = p->send_ctx;
    DeviceGuard g(ctx->device);
    const PartLayout need = part_layout(ctx, std::max<uint64_t>(keys->n, 1), k, num_bits, window_log2, W, true, true);
    PB_REQUIRE(need.sub_cap <= p->sub_cap, "chunk of %llu keys needs sublists of %u entries, the mailbox has %u",
               (unsigned long long)keys->n, need.sub_cap, p->sub_cap);
    const unsigned long long seq = ++p->send_seq;
    const int h = (int)(seq % kBufs);
    const size_t block_bytes = (size_t)p->wps * p->n_sub * p->sub_cap * 4;  // one destination's windows
    for (int q = 0; q < kBufs; ++q) {
        if (!p->lstage[q]) {
            cudaError_t e = cudaMalloc(&p->lstage[q], block_bytes * p->world);
            if (e != cudaSuccess) {
                cudaGetLastError();
                set_error("cudaMalloc of the local staging failed: %s", cudaGetErrorString(e));
                return PB_ERR_OOM;
            }
        }
    }
    // the copies that read this local buffer kBufs chunks ago are done
    if (seq > kBufs) PB_CUDA(cudaStreamWaitEvent(ctx->stream, p->ev_copy[h], 0));
    {
        PartDev pd;
        pd.stage = p->lstage[h];
        pd.counts = p->scnt[h];
        pd.words = nullptr;
        part_set_modulus(pd, num_bits);
        pd.sub_cap = p->sub_cap;
        pd.n_sub = p->n_sub;
        pd.window_log2 = window_log2;
        pd.n_windows = W;
        pd.k = k;
        pd.ovf_list = ovf_list_dev;
        pd.ovf_count = (unsigned long long *)ovf_count_dev;
        pd.ovf_cap = ovf_cap;
        DevKeys dk;
        dk.data = (const uint8_t *)keys->data;
        dk.offsets = nullptr;
        dk.n = keys->n;
        dk.stride = 16;
        dk.sym_width = 1;
        dk.base_symbol = 0;
        dk.total_bytes = keys->n * 16;
        // the tile size belongs to the layout the mailbox was created with (same decision on every rank)
        const int block = part_big_tile(ctx, W, k, true) ? 512 : 256;
        launch_begin(ctx);
        cudaError_t e = launch_part4(block, ctx->stream, dk, pd);
        if (e != cudaSuccess) {
            set_error("launch of bloom_part4 (k=%u) failed: %s", k, cudaGetErrorString(e));
            return PB_ERR_CUDA;
        }
        PB_TRY(check_launch(ctx, "bloom_part"));
    }
    // the copy engines push every destination's block into its mailbox while the SMs go on with pass 1 of the next
    // chunk and pass 2 of the previous one
    cudaStream_t cs = p->copy_stream;
    PB_CUDA(cudaEventRecord(p->ev_part[h], ctx->stream));
    PB_CUDA(cudaStreamWaitEvent(cs, p->ev_part[h], 0));
    if (seq > kBufs) PB_TRY(launch_wait(p, ctx, cs, off_done_flag(p, h), seq - kBufs));  // every owner applied what this buffer held
    launch_begin(ctx, cs);  // "p2p_copy" in pb_ctx_kernel_times: the pushes of this chunk, first byte to last
    PB_CUDA(cudaEventRecord(p->ev_lane_go, cs));
    const int lanes = (int)std::max<int64_t>(1, std::min<int64_t>(ctx->p2p_copy_lanes, kCopyLanes));
    for (int l = 0; l < lanes; ++l) PB_CUDA(cudaStreamWaitEvent(p->copy_lane[l], p->ev_lane_go, 0));
    for (uint32_t i = 0; i < p->world; ++i) {
        const uint32_t d = (p->rank + 1 + i) % p->world;  // stagger the destinations across ranks
        PB_CUDA(cudaMemcpyAsync(p->peer[d] + off_stage(p, h) + (size_t)p->rank * block_bytes,
                                reinterpret_cast<const uint8_t *>(p->lstage[h]) + (size_t)d * block_bytes, block_bytes,
                                cudaMemcpyDeviceToDevice, p->copy_lane[i % lanes]));
    }
    for (int l = 0; l < lanes; ++l) {
        PB_CUDA(cudaEventRecord(p->ev_lane_done[l], p->copy_lane[l]));
        PB_CUDA(cudaStreamWaitEvent(cs, p->ev_lane_done[l], 0));
    }
    PB_TRY(check_launch(ctx, "p2p_copy", cs));
    ctx->launches--;  // (not a kernel)
    p2p_publish<<<p->world, 256, 0, cs>>>(peers_of(p), p->scnt[h], p->wps * p->n_sub, p->rank, off_cnt(p, h), off_data_flag(p, h), seq);
    PB_TRY(check_launch(ctx, "p2p_publish", cs));
    PB_CUDA(cudaEventRecord(p->ev_copy[h], cs));
    return PB_OK;
}

// Pass 2 of the next chunk on this rank's shard (stream of the shard's context): waits for every source.
int pb_p2p_apply(pb_p2p *p, pb_bloom *shard, uint32_t active_windows, uint32_t window_log2) {
    PB_REQUIRE(p, "NULL argument");
    PB_REQUIRE(active_windows <= p->wps, "active_windows exceeds windows_per_rank");
    PB_REQUIRE(shard || active_windows == 0, "a shard handle is needed when this rank owns windows");
    pb_ctx *ctx = shard ? bloom_ctx(shard) : p->send_ctx;
    DeviceGuard g(ctx->device);
    const unsigned long long seq = ++p->apply_seq;
    const int h = (int)(seq % kBufs);
    PB_TRY(launch_wait(p, ctx, ctx->stream, off_data_flag(p, h), seq));
    if (active_windows) {
        // windows of 32 MiB and more: one window in flight (every resident CTA slot works on it) or L2 thrashes
        const int64_t per_sm = window_log2 >= 28 ? std::max<int64_t>(ctx->bloom_apply_cpw_per_sm, 8) : ctx->bloom_apply_cpw_per_sm;
        const uint32_t cpw = (uint32_t)ctx->num_sms * (uint32_t)std::max<int64_t>(1, std::min<int64_t>(per_sm, 32));
        PB_CUDA(apply_prefer_max_smem());
        launch_begin(ctx);
        bloom_apply_sources<<<active_windows * cpw, 256, 0, ctx->stream>>>(
            bloom_words(shard), reinterpret_cast<const uint32_t *>(p->local + off_stage(p, h)),
            reinterpret_cast<const uint32_t *>(p->local + off_cnt(p, h)), p->world, p->wps, p->n_sub, p->sub_cap, window_log2, cpw);
        PB_TRY(check_launch(ctx, "bloom_apply_windows"));
    }
    p2p_done<<<1, 32, 0, ctx->stream>>>(peers_of(p), p->world, p->rank, off_done_flag(p, h), seq);
    return check_launch(ctx, "p2p_done");
}

}  // extern "C"
