// Error reporting + trivial entry points of the C ABI (include/stove_b200.h).
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void stove_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* stove_last_error(void) { return g_err; }
extern "C" int stove_abi_version(void) { return 1; }

// ---------------------------------------------------------------------------------------
// bw_transform (model/utils/utils.py:10-15): y = clamp(sum_c x[:, c], 0, 1).
// Pure streaming: reads C*4 B, writes 4 B per pixel; float4 vectorised when hw % 4 == 0.
// ---------------------------------------------------------------------------------------
__global__ void bw_transform_kernel(const float* __restrict__ x, float* __restrict__ y,
                                    int64_t n, int C, int64_t hw) {
    int64_t total = n * hw;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t b = i / hw, p = i - b * hw;
        const float* src = x + (b * C) * hw + p;
        float acc = 0.f;
        for (int c = 0; c < C; ++c) acc += __ldg(src + c * hw);
        y[i] = fminf(fmaxf(acc, 0.f), 1.f);
    }
}
__global__ void bw_transform_kernel_v4(const float4* __restrict__ x, float4* __restrict__ y,
                                       int64_t n, int C, int64_t hw4) {
    int64_t total = n * hw4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t b = i / hw4, p = i - b * hw4;
        const float4* src = x + (b * C) * hw4 + p;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < C; ++c) {
            float4 v = __ldg(src + c * hw4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x = fminf(fmaxf(acc.x, 0.f), 1.f);
        acc.y = fminf(fmaxf(acc.y, 0.f), 1.f);
        acc.z = fminf(fmaxf(acc.z, 0.f), 1.f);
        acc.w = fminf(fmaxf(acc.w, 0.f), 1.f);
        y[i] = acc;
    }
}

extern "C" int stove_bw_transform(const float* x, float* y, int64_t n, int channels, int64_t hw,
                                  void* stream) {
    STOVE_CHECK_ARG(x && y && n >= 0 && channels > 0 && hw > 0, "bad argument");
    if (n == 0) return STOVE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 256;
    bool v4 = (hw % 4 == 0) && (((uintptr_t)x | (uintptr_t)y) % 16 == 0);
    int64_t items = v4 ? n * (hw / 4) : n * hw;
    int blocks = (int)((items + threads - 1) / threads);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (v4)
        bw_transform_kernel_v4<<<blocks, threads, 0, st>>>((const float4*)x, (float4*)y, n, channels, hw / 4);
    else
        bw_transform_kernel<<<blocks, threads, 0, st>>>(x, y, n, channels, hw);
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}
