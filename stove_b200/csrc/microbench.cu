// Micro-benchmarks behind the roofline denominators that MEASURED_PEAKS.json does not carry (SURVEY 8d asks for a
// measured FP32-FMA peak; VERDICT r01 for a measured tcgen05 TF32 peak): measurement infrastructure, not on the
// product path.  scripts/microbench.py runs them and writes profiles/r02_microbench.json.
#include <cuda.h>
#include "common.cuh"

namespace mb {
// 16 independent FMA chains per thread: no dependency stalls, no memory traffic
__global__ void __launch_bounds__(1024) ffma_kernel(float* out, int iters, float b, float c) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-6f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456f) out[0] = s;          // keeps the chains alive
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {      // K-major, 128-byte swizzle (see lstm_tc.cu)
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// one CTA per SM; one elected thread issues `iters` x 4 tcgen05.mma kind::tf32 (128 x N x 8 each) on resident
// shared-memory tiles: the issue-rate peak of the tensor pipe with operands from shared memory (SS mode)
template <int BN>
__global__ void __launch_bounds__(128, 1) tf32_mma_kernel(int iters) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const uint32_t tiles = smem_u32(smem);
    constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    float* f = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x)
        f[i] = __uint_as_float(__float_as_uint(1.0f + (float)((i * 2654435761u) >> 20) * 1e-4f) & 0xffffe000u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(slot);
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 0 && lane == 0) {
        const uint64_t ad = smem_desc(tiles), bd = smem_desc(tiles + A_BYTES);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t acc = (uint32_t)((it | k) != 0);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(IDESC), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
        } while (!done);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN) : "memory");
}

template <typename F>
static int time_ms(F launch, float* ms, cudaStream_t s) {
    cudaEvent_t a, b;
    STOVE_CUDA(cudaEventCreate(&a));
    STOVE_CUDA(cudaEventCreate(&b));
    launch();                                   // warm-up
    STOVE_CUDA(cudaEventRecord(a, s));
    launch();
    STOVE_CUDA(cudaEventRecord(b, s));
    STOVE_CUDA(cudaEventSynchronize(b));
    STOVE_CUDA(cudaEventElapsedTime(ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}
}  // namespace mb

// FP32 FMA throughput: returns TFLOP/s (2 FLOPs per FMA) of `iters` x 16 FMAs per thread on ctas x 1024 threads
extern "C" int stove_microbench_ffma(int ctas, int iters, float* scratch, double* tflops, void* stream) {
    STOVE_CHECK_ARG(ctas > 0 && iters > 0 && scratch && tflops, "bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    float ms = 0.f;
    const int rc = mb::time_ms([&] { mb::ffma_kernel<<<ctas, 1024, 0, s>>>(scratch, iters, 0.999f, 1e-3f); }, &ms, s);
    if (rc) return rc;
    *tflops = 2.0 * 16.0 * iters * 1024.0 * ctas / (ms * 1e-3) / 1e12;
    return STOVE_OK;
}

// tcgen05 kind::tf32 issue peak (operands resident in shared memory): TFLOP/s over ctas CTAs, N = 128 or 256
extern "C" int stove_microbench_tf32(int ctas, int iters, int n_cols, double* tflops, void* stream) {
    STOVE_CHECK_ARG(ctas > 0 && iters > 0 && tflops && (n_cols == 128 || n_cols == 256), "bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    float ms = 0.f;
    int rc;
    if (n_cols == 128) {
        const int smem = 128 * 128 + 128 * 128 + 64 + 1024;
        STOVE_CUDA(cudaFuncSetAttribute(mb::tf32_mma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        rc = mb::time_ms([&] { mb::tf32_mma_kernel<128><<<ctas, 128, smem, s>>>(iters); }, &ms, s);
    } else {
        const int smem = 128 * 128 + 256 * 128 + 64 + 1024;
        STOVE_CUDA(cudaFuncSetAttribute(mb::tf32_mma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        rc = mb::time_ms([&] { mb::tf32_mma_kernel<256><<<ctas, 128, smem, s>>>(iters); }, &ms, s);
    }
    if (rc) return rc;
    *tflops = 2.0 * 128.0 * n_cols * 8.0 * 4.0 * iters * ctas / (ms * 1e-3) / 1e12;
    return STOVE_OK;
}
