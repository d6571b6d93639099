// pb_micro.cu -- roofline micro-benchmarks: the empirical random-atomic ceiling the Bloom /
// Count-Min kernels are judged against.  No hashing: pre-generated uniform indices, one atomic each.
#include <algorithm>

#include "pb_common.cuh"
#include "pb_hash.cuh"

using namespace pb;

namespace pb {

__global__ void __launch_bounds__(256) micro_gen_idx(uint32_t *__restrict__ idx, uint64_t n, FastMod fm, uint64_t salt) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        idx[i] = (uint32_t)fastmod(sm64(i + salt), fm);
}

// op 0: RED.OR.b32 of a random bit, op 1: RED.ADD.s32 of 1, op 2: 32-bit gather (sum kept alive)
template <int OP>
__global__ void __launch_bounds__(256) micro_random(uint32_t *__restrict__ words, const uint32_t *__restrict__ idx, uint64_t n,
                                                    unsigned long long *sink) {
    uint32_t acc = 0;
    const uint64_t n4 = n >> 2;
    const uint4 *idx4 = reinterpret_cast<const uint4 *>(idx);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcs(idx4 + i);
        if (OP == 0) {
            atomicOr(words + v.x, 1u << (v.y & 31));
            atomicOr(words + v.y, 1u << (v.z & 31));
            atomicOr(words + v.z, 1u << (v.w & 31));
            atomicOr(words + v.w, 1u << (v.x & 31));
        } else if (OP == 1) {
            atomicAdd((int *)words + v.x, 1);
            atomicAdd((int *)words + v.y, 1);
            atomicAdd((int *)words + v.z, 1);
            atomicAdd((int *)words + v.w, 1);
        } else {
            acc += __ldg(words + v.x) + __ldg(words + v.y) + __ldg(words + v.z) + __ldg(words + v.w);
        }
    }
    if (OP == 2 && acc == 0x12345679u) atomicAdd(sink, 1ull);
}

// op 4: atomicOr into a CTA-private shared-memory tile (2^20 bits) at random bit positions: the ceiling of
// a shared-memory apply pass.  idx values are reused as bit positions (mod tile bits).
__global__ void __launch_bounds__(1024) micro_smem_or(const uint32_t *__restrict__ idx, uint64_t n, unsigned long long *sink) {
    extern __shared__ uint32_t tile[];
    constexpr uint32_t kWords = 32768;  // 128 KB
    for (uint32_t i = threadIdx.x; i < kWords; i += blockDim.x) tile[i] = 0;
    __syncthreads();
    const uint64_t n4 = n >> 2;
    const uint4 *idx4 = reinterpret_cast<const uint4 *>(idx);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcs(idx4 + i);
        atomicOr(tile + ((v.x >> 5) & (kWords - 1)), 1u << (v.x & 31));
        atomicOr(tile + ((v.y >> 5) & (kWords - 1)), 1u << (v.y & 31));
        atomicOr(tile + ((v.z >> 5) & (kWords - 1)), 1u << (v.z & 31));
        atomicOr(tile + ((v.w >> 5) & (kWords - 1)), 1u << (v.w & 31));
    }
    __syncthreads();
    uint32_t acc = 0;
    for (uint32_t i = threadIdx.x; i < kWords; i += blockDim.x) acc += __popc(tile[i]);
    if (acc == 0xFFFFFFFFu) atomicAdd(sink, 1ull);
}

__global__ void __launch_bounds__(256) micro_copy(const uint4 *__restrict__ src, uint4 *__restrict__ dst, uint64_t n4) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x)
        __stcs(dst + i, __ldcs(src + i));
}

}  // namespace pb

extern "C" {

// op: 0 = RED.OR.b32, 1 = RED.ADD.s32, 2 = random 32-bit load, 3 = streaming copy of `words` words (n ignored).
// Returns the best device time of `reps` runs (after one warm-up) in *ms_best.
int pb_microbench_random_atomic(pb_ctx *ctx, uint64_t words, uint64_t n, int op, int reps, float *ms_best) {
    PB_REQUIRE(ctx && ms_best, "NULL argument");
    PB_REQUIRE(op >= 0 && op <= 4, "op must be 0..4");
    PB_REQUIRE(words >= 4 && words <= 0xFFFFFFFFull, "words must be in 4..2^32-1");
    PB_REQUIRE(reps >= 1, "reps must be >= 1");
    DeviceGuard g(ctx->device);
    n = (n + 3) & ~(uint64_t)3;
    uint32_t *buf = nullptr, *idx = nullptr, *dst = nullptr;
    unsigned long long *sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int st = PB_OK;
    float best = 1e30f;
    auto fail = [&](const char *what, cudaError_t e) {
        set_error("%s failed: %s", what, cudaGetErrorString(e));
        st = e == cudaErrorMemoryAllocation ? PB_ERR_OOM : PB_ERR_CUDA;
    };
    cudaError_t e;
    const uint64_t words4 = (words + 3) & ~(uint64_t)3;
    if ((e = cudaMalloc(&buf, words4 * 4)) != cudaSuccess) fail("cudaMalloc", e);
    if (st == PB_OK && op == 3 && (e = cudaMalloc(&dst, words4 * 4)) != cudaSuccess) fail("cudaMalloc", e);
    if (st == PB_OK && op != 3 && (e = cudaMalloc(&idx, std::max<uint64_t>(n, 4) * 4)) != cudaSuccess) fail("cudaMalloc", e);
    if (st == PB_OK && (e = cudaMalloc(&sink, 8)) != cudaSuccess) fail("cudaMalloc", e);
    if (st == PB_OK) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaMemsetAsync(buf, 0, words4 * 4, ctx->stream);
        cudaMemsetAsync(sink, 0, 8, ctx->stream);
        const int grid = ctx->num_sms * 8;
        for (int r = 0; r <= reps && st == PB_OK; ++r) {
            if (op != 3) {
                micro_gen_idx<<<grid, 256, 0, ctx->stream>>>(idx, n, make_fastmod(words), 0x1234567ull * (uint64_t)(r + 1));
                ctx->launches++;
            }
            if (op == 0 || op == 1) cudaMemsetAsync(buf, 0, words4 * 4, ctx->stream);
            cudaEventRecord(e0, ctx->stream);
            switch (op) {
                case 0: micro_random<0><<<grid, 256, 0, ctx->stream>>>(buf, idx, n, sink); break;
                case 1: micro_random<1><<<grid, 256, 0, ctx->stream>>>(buf, idx, n, sink); break;
                case 2: micro_random<2><<<grid, 256, 0, ctx->stream>>>(buf, idx, n, sink); break;
                case 4:
                    cudaFuncSetAttribute(micro_smem_or, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
                    micro_smem_or<<<ctx->num_sms, 1024, 131072, ctx->stream>>>(idx, n, sink);
                    break;
                default: micro_copy<<<grid, 256, 0, ctx->stream>>>((const uint4 *)buf, (uint4 *)dst, words4 / 4); break;
            }
            ctx->launches++;
            cudaEventRecord(e1, ctx->stream);
            if ((e = cudaEventSynchronize(e1)) != cudaSuccess) {
                fail("microbench kernel", e);
                break;
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0 && ms < best) best = ms;
        }
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(idx);
    cudaFree(dst);
    cudaFree(sink);
    if (st == PB_OK) *ms_best = best;
    return st;
}

}  // extern "C"
