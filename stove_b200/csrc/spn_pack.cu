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
    // the padding Gaussians of the row (G .. GP - 1) are zero: written here, so the caller need not clear the table
    if (g == 0)
        for (int k = G; k < GP; ++k) { dst[k] = 0.f; dst[GP + k] = 0.f; dst[2 * GP + k] = 0.f; }
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
        // padding columns (S .. SP - 1) of the row: zero, written by the warp of column 0
        if (s == 0)
            for (int c = S; c < SP; ++c) { wlog[o + c] = 0.f; wlin[o + c] = 0.f; }
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

// Lane-interleaved copy of a packed leaf table with 8-float parameter rows (24 floats = 6 float4 per row): blocks of 32
// rows stored [6 parts][32 rows] float4.  The fused scene-likelihood kernels (scene_ll*.cu) walk the whole background
// leaf table in every CTA with lane = row: with the plain layout every 16-byte load of a warp touched 32 different
// cache lines and the pass was bound by the load unit (43 % of its stall samples: lg_throttle).
__global__ void interleave_leaf_kernel(const float4* __restrict__ src, const int32_t* __restrict__ row_map, int64_t rows,
                                       float4* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;       // (block, part, lane)
    const int64_t nblk = (rows + 31) / 32;
    if (i >= nblk * 6 * 32) return;
    const int lane = (int)(i & 31), part = (int)((i >> 5) % 6);
    const int64_t row = (i / (6 * 32)) * 32 + lane;
    const int sr = row < rows ? __ldg(row_map + row) : -1;
    out[i] = sr >= 0 ? __ldg(src + (int64_t)sr * 6 + part) : make_float4(0.f, 0.f, 0.f, 0.f);
}

extern "C" int stove_spn_interleave_leaf(const float* leaf, const int32_t* row_map, int64_t rows, float* out, void* stream) {
    STOVE_CHECK_ARG(leaf && row_map && out && rows >= 0, "bad argument");
    STOVE_CHECK_ARG(((uintptr_t)leaf & 15) == 0 && ((uintptr_t)out & 15) == 0, "tables must be 16-byte aligned");
    if (rows == 0) return STOVE_OK;
    const int64_t n = (rows + 31) / 32 * 6 * 32;
    cudaStream_t s = (cudaStream_t)stream;
    STOVE_KERNEL(K_PACK_LEAF_FWD, s, interleave_leaf_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<const float4*>(leaf), row_map, rows, reinterpret_cast<float4*>(out)));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}
