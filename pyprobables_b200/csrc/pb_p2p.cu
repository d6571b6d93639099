// pb_p2p.cu -- the multi-GPU Bloom insert as ONE fused compute + exchange kernel over NVLink peer memory.
//
// Every rank owns a "mailbox" in its own HBM (cudaMalloc'd, exported with CUDA IPC and mapped by every peer):
//     recv_stage[3][n_src][wps][cap]   window lists, one block per source rank; three buffers rotate by chunk number
//     recv_cur  [3][n_src][wps]        entries per list
//     data_flag [3][n_src]             written by source s: "chunk seq of mine is complete in your buffer h"
//     done_flag [3][n_dst]             written by destination d into the SOURCE's mailbox: "I applied your chunk seq"
// Three buffers, not two: with two, pass 1 of chunk c+2 would have to wait for pass 2 of chunk c on every
// destination (pass 1 + copy + pass 2 > 2 x pass 1), a bubble in every second chunk and no slack for rank skew.
// Pass 1 (bloom_part3_fixed16<.., P2P = true>) hashes a chunk of local keys, bins the bit indices by global
// window and stores every entry straight into the list of its window inside the OWNER's mailbox -- the
// all-to-all is the kernel's own coalesced stores travelling over NVLink while the next tile is being hashed.
// A tiny publish kernel then writes the list lengths and raises the data flags (release, system scope).
// Pass 2 on the owner waits for the flags of all sources (acquire), ORs the lists into its shard with the
// window L2 resident, and raises the done flags so the sources may reuse that half two chunks later.
// No NCCL on the data path, no host synchronisation; ordering is flags in peer memory.
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

struct pb_p2p {
    pb_ctx *send_ctx = nullptr;
    uint32_t world = 0, rank = 0, wps = 0, cap = 0;
    uint8_t *local = nullptr;      // this rank's mailbox
    uint8_t *peer[16] = {nullptr};  // mapped mailboxes (peer[rank] == local)
    bool opened[16] = {false};
    size_t stage_bytes = 0, cur_bytes = 0, total_bytes = 0;
    unsigned int *scur[kBufs] = {nullptr, nullptr, nullptr};  // local cursors of the chunk being partitioned [world*wps], per half
    uint64_t send_seq = 0, apply_seq = 0;
    // exchange variant: 1 = pass 1 stores straight into the owners' mailboxes (SM stores over NVLink);
    // 0 = pass 1 fills a local staging and the copy engines push every destination's block (DMA over NVLink),
    // which leaves the SMs to the two compute passes
    int direct = 0;
    uint32_t *lstage[kBufs] = {nullptr, nullptr, nullptr};  // local staging halves (DMA variant), allocated on first use
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_part[kBufs] = {nullptr, nullptr, nullptr}, ev_copy[kBufs] = {nullptr, nullptr, nullptr};
};

namespace pb {

static size_t off_stage(const pb_p2p *p, int h) { return (size_t)h * p->stage_bytes; }
static size_t off_cur(const pb_p2p *p, int h) { return kBufs * p->stage_bytes + (size_t)h * p->cur_bytes; }
static size_t off_data_flag(const pb_p2p *p, int h) { return kBufs * p->stage_bytes + kBufs * p->cur_bytes + (size_t)h * 16 * 8; }
static size_t off_done_flag(const pb_p2p *p, int h) {
    return kBufs * p->stage_bytes + kBufs * p->cur_bytes + (size_t)kBufs * 16 * 8 + (size_t)h * 16 * 8;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// lane r spins until flags[r] >= want (flags are raised by other GPUs through NVLink)
__global__ void p2p_wait_flags(const unsigned long long *flags, uint32_t n, unsigned long long want) {
    const uint32_t r = threadIdx.x;
    if (r < n) {
        while (ld_acquire_sys(flags + r) < want) __nanosleep(200);
    }
}

struct P2PPeers {
    uint8_t *base[16];
};

// block d: hand rank d the lengths of the lists this rank just wrote into its mailbox, then raise the flag
__global__ void p2p_publish(P2PPeers peers, const unsigned int *__restrict__ scur, uint32_t wps, uint32_t cap, uint32_t rank,
                            size_t cur_off, size_t flag_off, unsigned long long seq) {
    const uint32_t d = blockIdx.x;
    unsigned int *rc = reinterpret_cast<unsigned int *>(peers.base[d] + cur_off) + (size_t)rank * wps;
    for (uint32_t j = threadIdx.x; j < wps; j += blockDim.x) {
        const unsigned int c = scur[d * wps + j];
        rc[j] = c < cap ? c : cap;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(reinterpret_cast<unsigned long long *>(peers.base[d] + flag_off) + rank, seq);
}

// lane s: tell source s that its chunk `seq` has been applied here (its half may be refilled)
__global__ void p2p_done(P2PPeers peers, uint32_t n, uint32_t rank, size_t flag_off, unsigned long long seq) {
    const uint32_t s = threadIdx.x;
    __threadfence_system();
    if (s < n) st_release_sys(reinterpret_cast<unsigned long long *>(peers.base[s] + flag_off) + rank, seq);
}

static P2PPeers peers_of(const pb_p2p *p) {
    P2PPeers q;
    for (int i = 0; i < 16; ++i) q.base[i] = p->peer[i];
    return q;
}

}  // namespace pb

extern "C" {

int pb_p2p_create(pb_ctx *send_ctx, uint32_t world, uint32_t rank, uint32_t windows_per_rank, uint32_t cap, pb_p2p **out) {
    PB_REQUIRE(send_ctx && out, "NULL argument");
    PB_REQUIRE(world >= 1 && world <= 16 && rank < world, "world must be 1..16 and rank < world");
    PB_REQUIRE(windows_per_rank >= 1 && (uint64_t)windows_per_rank * world <= (uint64_t)kMaxWindows2, "too many windows");
    PB_REQUIRE((cap & 3u) == 0 && cap >= 4 && (uint64_t)cap * windows_per_rank * world <= 0xFFFFFFF0ull, "bad cap");
    DeviceGuard g(send_ctx->device);
    pb_p2p *p = new (std::nothrow) pb_p2p();
    if (!p) return PB_ERR_OOM;
    p->send_ctx = send_ctx;
    p->world = world;
    p->rank = rank;
    p->wps = windows_per_rank;
    p->cap = cap;
    p->stage_bytes = ((size_t)world * windows_per_rank * cap * 4 + 255) & ~(size_t)255;
    p->cur_bytes = ((size_t)world * windows_per_rank * 4 + 255) & ~(size_t)255;
    p->total_bytes = kBufs * p->stage_bytes + kBufs * p->cur_bytes + 2 * (size_t)kBufs * 16 * 8;
    cudaError_t e = cudaMalloc(&p->local, p->total_bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc of the %zu-byte P2P mailbox failed: %s", p->total_bytes, cudaGetErrorString(e));
        delete p;
        return PB_ERR_OOM;
    }
    for (int h = 0; h < kBufs; ++h) {
        e = cudaMalloc(&p->scur[h], (size_t)world * windows_per_rank * 4);
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
    // flags and counts start at zero; the lists need no initialisation
    PB_CUDA(cudaMemset(p->local + kBufs * p->stage_bytes, 0, kBufs * p->cur_bytes + 2 * (size_t)kBufs * 16 * 8));
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
        if (r == p->rank || p->opened[r]) continue;
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

int pb_p2p_destroy(pb_p2p *p) {
    if (!p) return PB_OK;
    DeviceGuard g(p->send_ctx->device);
    cudaDeviceSynchronize();
    for (uint32_t r = 0; r < p->world; ++r)
        if (p->opened[r]) cudaIpcCloseMemHandle(p->peer[r]);
    for (int h = 0; h < kBufs; ++h) {
        cudaFree(p->scur[h]);
        if (p->lstage[h]) cudaFree(p->lstage[h]);
        if (p->ev_part[h]) cudaEventDestroy(p->ev_part[h]);
        if (p->ev_copy[h]) cudaEventDestroy(p->ev_copy[h]);
    }
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    cudaFree(p->local);
    delete p;
    return PB_OK;
}

int pb_p2p_set_direct(pb_p2p *p, int direct_stores) {
    PB_REQUIRE(p, "NULL argument");
    PB_REQUIRE(p->send_seq == 0, "choose the exchange variant before the first chunk");
    p->direct = direct_stores ? 1 : 0;
    return PB_OK;
}

// Pass 1 + exchange of one chunk of this rank's keys (stream of the send context).
int pb_p2p_partition_send(pb_p2p *p, const pb_keys *keys, uint64_t num_bits, uint32_t k, uint32_t window_log2,
                          uint64_t *ovf_list_dev, uint64_t ovf_cap, uint64_t *ovf_count_dev) {
    PB_REQUIRE(p && keys && ovf_list_dev && ovf_count_dev, "NULL argument");
    PB_REQUIRE(keys->on_device, "pb_p2p_partition_send takes device keys");
    PB_REQUIRE(keys->offsets == nullptr && keys->sym_width == 1 && keys->stride == 16 && ((uintptr_t)keys->data & 15u) == 0,
               "partitioned routing takes fixed 16-byte keys");
    PB_REQUIRE(k >= 1 && k <= 16, "k must be in 1..16");
    PB_REQUIRE(window_log2 >= 5 && window_log2 <= 31, "window_log2 must be in 5..31");
    const uint32_t W = p->world * p->wps;
    PB_REQUIRE(((uint64_t)W << window_log2) >= num_bits, "windows do not cover the filter");
    PB_REQUIRE(keys->n * (uint64_t)k < 0xE0000000ull, "too many keys for one partition call");
    for (uint32_t r = 0; r < p->world; ++r) PB_REQUIRE(p->peer[r] != nullptr, "rank %u is not connected (pb_p2p_connect)", r);
    pb_ctx *ctx = p->send_ctx;
    DeviceGuard g(ctx->device);
    const unsigned long long seq = ++p->send_seq;
    const int h = (int)(seq % kBufs);
    const size_t block_bytes = (size_t)p->wps * p->cap * 4;  // one destination's windows
    if (!p->direct) {
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
    } else if (seq > kBufs) {
        // direct stores: every destination must have applied the chunk that used this mailbox buffer last
        p2p_wait_flags<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long *>(p->local + off_done_flag(p, h)), p->world,
                                                 seq - kBufs);
        PB_TRY(check_launch(ctx, "p2p_wait"));
    }
    PB_CUDA(cudaMemsetAsync(p->scur[h], 0, (size_t)W * 4, ctx->stream));
    if (keys->n) {
        Part2Dev pd;
        pd.stage = p->direct ? nullptr : p->lstage[h];
        pd.cursors = p->scur[h];
        pd.words = nullptr;
        part_set_modulus(pd, num_bits);
        pd.cap = p->cap;
        pd.window_log2 = window_log2;
        pd.n_windows = W;
        pd.k = k;
        const int grid = grid_for(ctx, keys->n, 256, 4);
        const bool big_tile = grid >= 2 && (ctx->bloom_part_tile == 512 || (ctx->bloom_part_tile != 256 && W > 112));
        {
            const double per = (double)keys->n * k / ((double)W * (double)std::max(grid, 1));
            uint32_t q = 32;
            while (q < kQuota && (double)q * 16.0 < per) q <<= 1;
            pd.quota = q;
        }
        pd.ovf_list = ovf_list_dev;
        pd.ovf_count = (unsigned long long *)ovf_count_dev;
        pd.ovf_cap = ovf_cap;
        P2PDst dst;
        for (int i = 0; i < 16; ++i) dst.stage[i] = nullptr;
        for (uint32_t r = 0; r < p->world; ++r) dst.stage[r] = reinterpret_cast<uint32_t *>(p->peer[r] + off_stage(p, h));
        dst.wps = p->wps;
        dst.src_rank = p->rank;
        const uint4 *k4 = (const uint4 *)keys->data;
        launch_begin(ctx);
        if (!p->direct && ctx->bloom_part_version != 3) {
            // DMA variant: the staging is local and contiguous, exactly the single-GPU layout
            DevKeys dk;
            dk.data = (const uint8_t *)keys->data;
            dk.offsets = nullptr;
            dk.n = keys->n;
            dk.stride = 16;
            dk.sym_width = 1;
            dk.base_symbol = 0;
            dk.total_bytes = keys->n * 16;
            cudaError_t e = launch_part4(k, big_tile, grid, ctx->stream, dk, pd);
            if (e != cudaSuccess) {
                set_error("launch of bloom_part4 (k=%u) failed: %s", k, cudaGetErrorString(e));
                return PB_ERR_CUDA;
            }
        } else {
            const int kg = k <= 8 ? (int)k : (int)((k + 1) / 2), ng = k <= 8 ? 1 : 2;
            switch (ng * 100 + kg) {
#define PB_P3(KG, NG)                                                                                        \
    case NG * 100 + KG:                                                                                      \
        if (p->direct) launch_part3<KG, NG, true>(big_tile, grid, ctx->stream, k4, keys->n, pd, dst);        \
        else launch_part3<KG, NG, false>(big_tile, grid, ctx->stream, k4, keys->n, pd, dst);                 \
        break;
                PB_P3(1, 1) PB_P3(2, 1) PB_P3(3, 1) PB_P3(4, 1) PB_P3(5, 1) PB_P3(6, 1) PB_P3(7, 1) PB_P3(8, 1)
                PB_P3(5, 2) PB_P3(6, 2) PB_P3(7, 2) PB_P3(8, 2)
#undef PB_P3
                default: set_error("internal: no partition kernel for k=%u", k); return PB_ERR_UNSUPPORTED;
            }
        }
        PB_TRY(check_launch(ctx, "bloom_part"));
    }
    if (p->direct) {
        p2p_publish<<<p->world, 128, 0, ctx->stream>>>(peers_of(p), p->scur[h], p->wps, p->cap, p->rank, off_cur(p, h),
                                                      off_data_flag(p, h), seq);
        return check_launch(ctx, "p2p_publish");
    }
    // DMA variant: the copy engines push every destination's block into its mailbox while the SMs go on with
    // pass 1 of the next chunk and pass 2 of the previous one
    cudaStream_t cs = p->copy_stream;
    PB_CUDA(cudaEventRecord(p->ev_part[h], ctx->stream));
    PB_CUDA(cudaStreamWaitEvent(cs, p->ev_part[h], 0));
    if (seq > kBufs) {
        p2p_wait_flags<<<1, 32, 0, cs>>>(reinterpret_cast<const unsigned long long *>(p->local + off_done_flag(p, h)), p->world,
                                         seq - kBufs);
        PB_TRY(check_launch(ctx, "p2p_wait", cs));
    }
    for (uint32_t i = 0; i < p->world; ++i) {
        const uint32_t d = (p->rank + 1 + i) % p->world;  // stagger the destinations across ranks
        PB_CUDA(cudaMemcpyAsync(p->peer[d] + off_stage(p, h) + (size_t)p->rank * block_bytes,
                                reinterpret_cast<const uint8_t *>(p->lstage[h]) + (size_t)d * block_bytes, block_bytes,
                                cudaMemcpyDeviceToDevice, cs));
    }
    p2p_publish<<<p->world, 128, 0, cs>>>(peers_of(p), p->scur[h], p->wps, p->cap, p->rank, off_cur(p, h), off_data_flag(p, h), seq);
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
    p2p_wait_flags<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<const unsigned long long *>(p->local + off_data_flag(p, h)), p->world, seq);
    PB_TRY(check_launch(ctx, "p2p_wait"));
    if (active_windows) {
        const int64_t per_sm = window_log2 >= 28 ? std::max<int64_t>(ctx->bloom_apply_cpw_per_sm, 8) : ctx->bloom_apply_cpw_per_sm;
        const uint32_t cpw = (uint32_t)ctx->num_sms * (uint32_t)std::max<int64_t>(1, std::min<int64_t>(per_sm, 32));
        launch_begin(ctx);
        bloom_apply_sources<<<active_windows * cpw, 256, 0, ctx->stream>>>(
            bloom_words(shard), reinterpret_cast<const uint32_t *>(p->local + off_stage(p, h)),
            reinterpret_cast<const unsigned int *>(p->local + off_cur(p, h)), p->world, p->wps, p->cap, window_log2, cpw);
        PB_TRY(check_launch(ctx, "bloom_apply_windows"));
    }
    p2p_done<<<1, 32, 0, ctx->stream>>>(peers_of(p), p->world, p->rank, off_done_flag(p, h), seq);
    return check_launch(ctx, "p2p_done");
}

}  // extern "C"
