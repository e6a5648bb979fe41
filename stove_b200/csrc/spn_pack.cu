// RAT-SPN parameter packing: raw nn.Parameters -> the layouts the fused kernels read.
//   leaf : var = vmin + (vmax - vmin) * sigmoid(sigma_params)      (rat_torch.py:85-87,98-99)
//          packed (mu, a = 1/(2 var), b = 0.5 log var + 0.5 log 2pi)
//   sums : column-wise log_softmax over the K inputs                (rat_torch.py:209-210)
// These are a few 10k elements, once per step; they exist so the hot kernels never touch
// sigmoid/log/softmax of parameters.
#include "common.cuh"

__global__ void pack_leaf_fwd_kernel(const float* __restrict__ means, const float* __restrict__ sig,
                                     const int32_t* __restrict__ dst_row, int rows, int G, int GP,
                                     float vmin, float vmax, float* __restrict__ packed) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * G) return;
    int row = i / G, g = i - row * G;
    float var = vmin + (vmax - vmin) * sigmoidf_(sig[i]);
    float* dst = packed + (int64_t)dst_row[row] * 3 * GP;
    dst[g] = means[i];
    dst[GP + g] = 0.5f / var;
    dst[2 * GP + g] = 0.5f * logf(var) + HALF_LOG_2PI;
}

__global__ void pack_leaf_bwd_kernel(const float* __restrict__ sig, const int32_t* __restrict__ dst_row,
                                     int rows, int G, int GP, float vmin, float vmax,
                                     const float* __restrict__ g_packed, float* __restrict__ g_means,
                                     float* __restrict__ g_sig) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * G) return;
    int row = i / G, g = i - row * G;
    float s = sigmoidf_(sig[i]);
    float var = vmin + (vmax - vmin) * s;
    const float* src = g_packed + (int64_t)dst_row[row] * 3 * GP;
    g_means[i] = src[g];
    // a = 0.5/var -> da/dvar = -0.5/var^2 ; b = 0.5 log var -> db/dvar = 0.5/var
    float gvar = src[GP + g] * (-0.5f / (var * var)) + src[2 * GP + g] * (0.5f / var);
    g_sig[i] = gvar * (vmax - vmin) * s * (1.f - s);
}

// one warp per (block, column)
__global__ void pack_sum_fwd_kernel(const float* __restrict__ raw, int nb, int K, int S, int SP,
                                    float* __restrict__ wlog, float* __restrict__ wlin) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nb * S) return;
    int b = warp / S, s = warp - b * S;
    const float* src = raw + (int64_t)b * K * S + s;
    float m = -INFINITY;
    for (int k = lane; k < K; k += 32) m = fmaxf(m, src[(int64_t)k * S]);
    m = warp_max(m);
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc += expf(src[(int64_t)k * S] - m);
    float lse = m + logf(warp_sum(acc));
    for (int k = lane; k < K; k += 32) {
        float lw = src[(int64_t)k * S] - lse;
        int64_t o = ((int64_t)b * K + k) * SP + s;
        wlog[o] = lw;
        wlin[o] = expf(lw);
    }
}

// g_raw[k] = G[k] - softmax[k] * sum_k' G[k']   (G = gradient w.r.t. the log-weights)
__global__ void pack_sum_bwd_kernel(const float* __restrict__ wlog, int nb, int K, int S, int SP,
                                    const float* __restrict__ g_wlog, float* __restrict__ g_raw) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nb * S) return;
    int b = warp / S, s = warp - b * S;
    float tot = 0.f;
    for (int k = lane; k < K; k += 32) tot += g_wlog[((int64_t)b * K + k) * SP + s];
    tot = warp_sum(tot);
    for (int k = lane; k < K; k += 32) {
        int64_t o = ((int64_t)b * K + k) * SP + s;
        g_raw[((int64_t)b * K + k) * S + s] = g_wlog[o] - expf(wlog[o]) * tot;
    }
}

extern "C" int stove_spn_pack_leaf_fwd(const float* means, const float* sigma_params,
                                       const int32_t* dst_row, int rows, int G, int GP,
                                       float min_var, float max_var, float* packed, void* stream) {
    STOVE_CHECK_ARG(means && sigma_params && dst_row && packed && rows > 0 && G > 0 && GP >= G, "bad argument");
    int total = rows * G;
    STOVE_KERNEL(K_PACK_LEAF_FWD, (cudaStream_t)stream, pack_leaf_fwd_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        means, sigma_params, dst_row, rows, G, GP, min_var, max_var, packed));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" int stove_spn_pack_leaf_bwd(const float* sigma_params, const int32_t* dst_row, int rows,
                                       int G, int GP, float min_var, float max_var,
                                       const float* g_packed, float* g_means, float* g_sigma_params,
                                       void* stream) {
    STOVE_CHECK_ARG(sigma_params && dst_row && g_packed && g_means && g_sigma_params && rows > 0, "bad argument");
    int total = rows * G;
    STOVE_KERNEL(K_PACK_LEAF_BWD, (cudaStream_t)stream, pack_leaf_bwd_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        sigma_params, dst_row, rows, G, GP, min_var, max_var, g_packed, g_means, g_sigma_params));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" int stove_spn_pack_sum_fwd(const float* raw, int nb, int K, int S, int SP, float* wlog,
                                      float* wlin, void* stream) {
    STOVE_CHECK_ARG(raw && wlog && wlin && nb > 0 && K > 0 && S > 0 && SP >= S, "bad argument");
    int warps = nb * S;
    STOVE_KERNEL(K_PACK_SUM_FWD, (cudaStream_t)stream, pack_sum_fwd_kernel<<<(warps * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(raw, nb, K, S, SP, wlog, wlin));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" int stove_spn_pack_sum_bwd(const float* wlog, int nb, int K, int S, int SP,
                                      const float* g_wlog, float* g_raw, void* stream) {
    STOVE_CHECK_ARG(wlog && g_wlog && g_raw && nb > 0 && K > 0 && S > 0 && SP >= S, "bad argument");
    int warps = nb * S;
    STOVE_KERNEL(K_PACK_SUM_BWD, (cudaStream_t)stream, pack_sum_bwd_kernel<<<(warps * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(wlog, nb, K, S, SP, g_wlog, g_raw));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}
