// Optimizer step of the reference trainer on the flat gradient bucket
// (model/video_prediction/train.py:46-49 Adam(amsgrad), :471-473 clip_grad_norm_(..., 1) + step):
// two launches instead of torch's multi-tensor foreach passes.
//   adam_norm  partial sums of squares of the bucket (fixed grid, fixed order) and step += 1
//   adam_step  every CTA folds the partials in the same order -> global norm -> clip coefficient; then
//              m, v, (vmax) and the parameters are updated in place; the clipped gradient is written back
//              (clip_grad_norm_ scales .grad in place).  Parameters are addressed through a pointer table
//              passed by value (they keep their own storage), moments live in flat buffers.
#include "common.cuh"

namespace opt {
constexpr int NORM_CTAS = 128, MAX_T = 128, CHUNK = 4096;

// partial[0 .. NORM_CTAS) = partial sums of squares; partial[NORM_CTAS] = lr / (1 - beta1^k), partial[NORM_CTAS + 1]
// = sqrt(1 - beta2^k) for the step count k after the increment: the double-precision powers are evaluated ONCE
// here (FP64 runs at 1/64 rate on this part; per CTA of the update kernel they cost ~5 us of latency each)
__global__ void __launch_bounds__(256) adam_norm_kernel(const float* __restrict__ g, int64_t total,
                                                        float* __restrict__ partial, float* __restrict__ step,
                                                        const float* __restrict__ lr, float b1, float b2) {
    __shared__ float red[8];
    float s = 0.f;
    const int64_t n4 = total >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0) {
        for (int64_t i = (n4 << 2) + threadIdx.x; i < total; i += blockDim.x) s += g[i] * g[i];
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        partial[blockIdx.x] = t;
        if (blockIdx.x == 0) {
            const float k = step[0] + 1.f;
            step[0] = k;
            partial[NORM_CTAS] = (float)((double)lr[0] / (1.0 - pow((double)b1, (double)k)));
            partial[NORM_CTAS + 1] = (float)sqrt(1.0 - pow((double)b2, (double)k));
        }
    }
}

struct Table {
    float* p[MAX_T];
    int64_t off[MAX_T];
    int64_t num[MAX_T];
};
struct Hyper {
    float b1, b2, eps, max_norm;
    int amsgrad;
};

__global__ void __launch_bounds__(256) adam_step_kernel(const __grid_constant__ Table tb, Hyper h, float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        float* __restrict__ vmax, const float* __restrict__ partial) {
    const int t = blockIdx.y;
    const int64_t num = tb.num[t];
    const int64_t lo = (int64_t)blockIdx.x * CHUNK;
    if (lo >= num) return;
    __shared__ float sh[3];
    if (threadIdx.x < 32) {
        float s = 0.f;
        for (int i = threadIdx.x; i < NORM_CTAS; i += 32) s += partial[i];
        s = warp_sum(s);
        if (threadIdx.x == 0) {
            const float norm = sqrtf(s);
            const float coef = h.max_norm > 0.f ? fminf(h.max_norm / (norm + 1e-6f), 1.f) : 1.f;
            sh[0] = coef;
            sh[1] = partial[NORM_CTAS];                 // step size lr / (1 - beta1^k)
            sh[2] = partial[NORM_CTAS + 1];             // sqrt(1 - beta2^k)
        }
    }
    __syncthreads();
    const float coef = sh[0], step_size = sh[1], bc2_sqrt = sh[2];
    const int64_t hi = lo + CHUNK < num ? lo + CHUNK : num;
    float* __restrict__ p = tb.p[t];
    const int64_t base = tb.off[t];
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int64_t f = base + i;
        const float gi = g[f] * coef;
        const float m0 = m[f];
        const float mi = m0 + (1.f - h.b1) * (gi - m0);          // torch: exp_avg.lerp_(grad, 1 - beta1)
        const float vi = h.b2 * v[f] + (1.f - h.b2) * gi * gi;
        float den = vi;
        if (h.amsgrad) {
            den = fmaxf(vmax[f], vi);
            vmax[f] = den;
        }
        g[f] = gi;
        m[f] = mi;
        v[f] = vi;
        p[i] -= step_size * (mi / (sqrtf(den) / bc2_sqrt + h.eps));
    }
}
}  // namespace opt

extern "C" int stove_adam_workspace_floats(void) { return opt::NORM_CTAS + 2; }

extern "C" int stove_adam_step(const void* const* params, const int64_t* offsets, const int64_t* numels, int count,
                               int64_t total, float* flat_grad, float* exp_avg, float* exp_avg_sq,
                               float* max_exp_avg_sq, float* partial, const float* lr, float* step, float beta1,
                               float beta2, float eps, float max_norm, void* stream) {
    using namespace opt;
    STOVE_CHECK_ARG(count >= 0 && total >= 0 && flat_grad && exp_avg && exp_avg_sq && partial && lr && step,
                    "bad argument");
    STOVE_CHECK_ARG((((uintptr_t)flat_grad) & 15) == 0, "flat_grad must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    STOVE_KERNEL(K_ADAM_NORM, st, adam_norm_kernel<<<NORM_CTAS, 256, 0, st>>>(flat_grad, total, partial, step, lr, beta1, beta2));
    STOVE_LAUNCH_CHECK();
    Hyper h;
    h.b1 = beta1; h.b2 = beta2; h.eps = eps; h.max_norm = max_norm; h.amsgrad = max_exp_avg_sq != nullptr;
    for (int pass = 0; pass < 2; ++pass) {
        Table tb;
        int k = 0;
        int64_t mx = 0;
        auto flush = [&]() -> int {
            if (k == 0) return STOVE_OK;
            const dim3 grid((unsigned)((mx + CHUNK - 1) / CHUNK), (unsigned)k);
            STOVE_KERNEL(K_ADAM_STEP, st, adam_step_kernel<<<grid, 256, 0, st>>>(tb, h, flat_grad, exp_avg, exp_avg_sq,
                                                                                 max_exp_avg_sq, partial));
            STOVE_LAUNCH_CHECK();
            k = 0;
            mx = 0;
            return STOVE_OK;
        };
        for (int i = 0; i < count; ++i) {
            const bool big = numels[i] > 4 * CHUNK;
            if (numels[i] <= 0 || big != (pass == 1)) continue;
            STOVE_CHECK_ARG(params[i] != nullptr && offsets[i] >= 0 && offsets[i] + numels[i] <= total, "bad table entry");
            tb.p[k] = (float*)params[i];
            tb.off[k] = offsets[i];
            tb.num[k] = numels[i];
            if (numels[i] > mx) mx = numels[i];
            if (++k == MAX_T) {
                const int rc = flush();
                if (rc != STOVE_OK) return rc;
            }
        }
        const int rc = flush();
        if (rc != STOVE_OK) return rc;
    }
    return STOVE_OK;
}
