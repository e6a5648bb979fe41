// Fused sequence glue of Stove.stove_forward: everything between the encoder output and the
// dynamics loop, one thread per sequence, forward and backward.
//
// Replaces (reference file:line, all in model/video_prediction/):
//   Supair.constrain_zp                 supair.py:112-149
//   Stove.match_objects (3 variants)    stove.py:200-329 (3_only), 331-430 (volatile), 432-514 (greedy)
//   Stove.fix_supair                    stove.py:516-571
//   Stove.v_from_state / v_std_from_pos stove.py:54-101
// ~215 ATen launches forward and ~300 backward in the tensor version (and one host sync per
// time step in the reference, stove.py:271-273) become two launches.  The matching is index
// arithmetic on detached values; gradients flow through the gather, the smoothing select and
// the finite differences only.
#include "common.cuh"

#define SG_MAX_T 16
#define SG_MAX_O 9

struct SupParams {
    int T, O, match_kind;     // 0 = 3_only, 1 = greedy, 2 = volatile
    int app_dim;              // 0 or 3
    int match_app;            // use appearance in the matching distance
    float scale_lo, scale_hi, ratio_lo, ratio_hi, pos_bound, scale_var, pos_var, fix_eps;
    int fix;                  // debug_fix_supair
};

__device__ __forceinline__ float sq(float v) { return v * v; }

// errors err[a][j] = |prev_a - cur_j|^2 over the matching features.  OC > 0: compile-time object count
// (all loops unroll and every array index is static, so the working set stays in registers); OC = 0:
// run-time count, arrays in local memory.
template <int OC>
__device__ __forceinline__ void match_errors(const SupParams& p, const float* prev_pos, const float* prev_app,
                                             const float* cur_pos, const float* cur_app, float* err) {
    const int O = OC ? OC : p.O;
#pragma unroll
    for (int a = 0; a < O; ++a)
#pragma unroll
        for (int j = 0; j < O; ++j) {
            float e = sq(prev_pos[a * 2] - cur_pos[j * 2]) + sq(prev_pos[a * 2 + 1] - cur_pos[j * 2 + 1]);
            if (p.match_app)
#pragma unroll
                for (int c = 0; c < 3; ++c) e += sq(prev_app[a * 3 + c] - cur_app[j * 3 + c]);
            err[a * O + j] = e;
        }
}

template <int OC>
__device__ __forceinline__ int row_argmin(const float* err, int a, int O) {
    int best = 0;
    float bv = err[a * O];
#pragma unroll
    for (int j = 1; j < (OC ? OC : O); ++j)
        if (err[a * O + j] < bv) {
            bv = err[a * O + j];
            best = j;
        }
    return best;
}

template <int OC>
__device__ __forceinline__ void match_step(const SupParams& p, float* err, int* idx) {
    const int O = OC ? OC : p.O;
    if (p.match_kind == 2) {                         // volatile: argmin over current for each previous
#pragma unroll
        for (int a = 0; a < O; ++a) idx[a] = row_argmin<OC>(err, a, O);
        return;
    }
    if (p.match_kind == 0) {                         // 3_only
#pragma unroll
        for (int a = 0; a < O; ++a) idx[a] = row_argmin<OC>(err, a, O);
        const bool valid = idx[0] != idx[1] && idx[1] != idx[2] && idx[0] != idx[2];
        if (valid) return;
        // greedy repair (stove.py:278-295): row o takes its current argmin, that column is
        // then knocked out for every row
#pragma unroll
        for (int o = 0; o < O; ++o) {
            const int best = row_argmin<OC>(err, o, O);
            idx[o] = best;
#pragma unroll
            for (int a = 0; a < O; ++a)
#pragma unroll
                for (int j = 0; j < O; ++j)
                    if (j == best) err[a * O + j] = 1e12f;
        }
        return;
    }
    // greedy bipartite (stove.py:488-494): repeatedly take the global minimum, retire its row/column
#pragma unroll 1
    for (int it = 0; it < O; ++it) {
        int ba = 0, bj = 0;
        float bv = err[0], mx = err[0];
#pragma unroll
        for (int a = 0; a < O; ++a)
#pragma unroll
            for (int j = 0; j < O; ++j) {
                const float e = err[a * O + j];
                if (e < bv) {
                    bv = e;
                    ba = a;
                    bj = j;
                }
                mx = fmaxf(mx, e);
            }
#pragma unroll
        for (int a = 0; a < O; ++a)
            if (a == ba) idx[a] = bj;
        const float big = mx + 1.f, big2 = big + 1.f;      // the reference re-evaluates max(errors) + 1
#pragma unroll
        for (int a = 0; a < O; ++a)
#pragma unroll
            for (int j = 0; j < O; ++j) {
                if (a == ba) err[a * O + j] = big;
                if (j == bj) err[a * O + j] = big2;
            }
    }
}

// One CTA per sequence; its working set lives in shared memory.  Only the
// matching itself (T-1 tiny assignment problems on detached values) runs on a single lane.
// raw encoder output zp [n][T][O][8] -> z_sup [n][T][O][4], z_full / std_full [n][T][O][6],
// app_out [n][T][O][3] (matched appearances), idx [n][T][O] (int32), flag [n][T][O] (bit0, bit1)
#define SGB_THREADS 256

template <int OC>
__global__ void __launch_bounds__(SGB_THREADS) sup_prepare_fwd_kernel(
    SupParams p, int64_t n, const float* __restrict__ zp, const float* __restrict__ app,
    float* __restrict__ z_sup, float* __restrict__ z_full, float* __restrict__ std_full,
    float* __restrict__ app_out, int32_t* __restrict__ idx_out, int32_t* __restrict__ flag_out) {
    extern __shared__ float sg_smem[];
    const int lane = threadIdx.x;              // one CTA per sequence, as in the backward kernel
    const int64_t b = blockIdx.x;
    (void)n;
    const int T = p.T, O = OC ? OC : p.O, TO = T * O;
    float* z = sg_smem;                                    // constrained, later the smoothed tensor
    float* zm = z + TO * 8;                                 // matched
    int* idx = reinterpret_cast<int*>(zm + TO * 8);         // [T][O]
    int* flg = idx + TO;
    const float* src = zp + b * TO * 8;
    const float* asrc = app ? app + b * TO * 3 : nullptr;
    for (int q = lane; q < TO * 8; q += SGB_THREADS) {
        const int f = q & 7;
        const float sgm = sigmoidf_(__ldg(src + q));
        float v;
        if (f == 0) v = p.scale_lo + (p.scale_hi - p.scale_lo) * sgm;
        else if (f == 1) v = p.ratio_lo + (p.ratio_hi - p.ratio_lo) * sgm;
        else if (f < 4) v = p.pos_bound * (2.f * sgm - 1.f);
        else if (f < 6) v = p.scale_var * sgm;
        else v = p.pos_var * sgm;
        z[q] = v;
    }
    __syncthreads();
    if (lane == 0) {
        // matching on positions scaled to [0, 1] ((z + 1) / 2, stove.py:220), detached
        constexpr int OM = OC ? OC : SG_MAX_O;
        float prev_pos[OM * 2], cur_pos[OM * 2], prev_app[OM * 3], cur_app[OM * 3], err[OM * OM];
        int cur[OM];
#pragma unroll
        for (int a = 0; a < O; ++a) {
            idx[a] = a;
#pragma unroll
            for (int d = 0; d < 2; ++d) prev_pos[a * 2 + d] = (z[a * 8 + 2 + d] + 1.f) * 0.5f;
#pragma unroll
            for (int c = 0; c < 3; ++c) prev_app[a * 3 + c] = asrc ? asrc[a * 3 + c] : 0.f;
        }
#pragma unroll 1
        for (int t = 1; t < T; ++t) {
#pragma unroll
            for (int a = 0; a < O; ++a) {
#pragma unroll
                for (int d = 0; d < 2; ++d) cur_pos[a * 2 + d] = (z[(t * O + a) * 8 + 2 + d] + 1.f) * 0.5f;
#pragma unroll
                for (int c = 0; c < 3; ++c) cur_app[a * 3 + c] = asrc ? asrc[(t * O + a) * 3 + c] : 0.f;
            }
            match_errors<OC>(p, prev_pos, prev_app, cur_pos, cur_app, err);
            match_step<OC>(p, err, cur);
#pragma unroll
            for (int a = 0; a < O; ++a) {
                idx[t * O + a] = cur[a];
#pragma unroll
                for (int d = 0; d < 2; ++d) prev_pos[a * 2 + d] = (z[(t * O + cur[a]) * 8 + 2 + d] + 1.f) * 0.5f;
                if (asrc)
#pragma unroll
                    for (int c = 0; c < 3; ++c) prev_app[a * 3 + c] = asrc[(t * O + cur[a]) * 3 + c];
            }
        }
    }
    __syncthreads();
    for (int q = lane; q < TO * 8; q += SGB_THREADS) {
        const int ta = q >> 3, f = q & 7, t = ta / O;
        zm[q] = z[(t * O + idx[ta]) * 8 + f];
    }
    for (int q = lane; q < TO; q += SGB_THREADS) idx_out[b * TO + q] = idx[q];
    if (app_out && asrc)
        for (int q = lane; q < TO * 3; q += SGB_THREADS) {
            const int ta = q / 3, c = q - ta * 3, t = ta / O;
            app_out[b * TO * 3 + q] = asrc[(t * O + idx[ta]) * 3 + c];
        }
    __syncthreads();
    // smoothing of glitches (stove.py:516-571): flags from the first two features
    for (int q = lane; q < TO; q += SGB_THREADS) {
        const int t = q / O;
        int fl = 0;
        if (p.fix && t >= 1 && t + 1 < T)
            for (int d = 0; d < 2; ++d) {
                const float before = fabsf(zm[q * 8 + d] - zm[(q - O) * 8 + d]);
                const float after = fabsf(zm[(q + O) * 8 + d] - zm[q * 8 + d]);
                if (before > p.fix_eps && after > p.fix_eps) fl |= 1 << d;
            }
        flg[q] = fl;
        flag_out[b * TO + q] = fl;
    }
    __syncthreads();
    for (int q = lane; q < TO * 8; q += SGB_THREADS) {
        const int ta = q >> 3, f = q & 7;
        float v = zm[q];
        if (flg[ta] & (1 << (f & 1))) v = 0.5f * (zm[q - O * 8] + zm[q + O * 8]);
        z[q] = v;
    }
    __syncthreads();
    for (int q = lane; q < TO * 6; q += SGB_THREADS) {
        const int ta = q / 6, f = q - ta * 6, t = ta / O;
        const int64_t o6 = (b * TO + ta) * 6 + f;
        float zf = 0.f, sf = 0.f;
        if (t > 0) {
            if (f < 4) {
                zf = z[ta * 8 + f];
                sf = z[ta * 8 + 4 + f];
            } else {
                const int d = f - 4;
                zf = z[ta * 8 + 2 + d] - z[(ta - O) * 8 + 2 + d];
                sf = sqrtf(sq(z[ta * 8 + 6 + d]) + sq(z[(ta - O) * 8 + 6 + d]));
            }
        }
        z_full[o6] = zf;
        std_full[o6] = sf;
        if (f < 4) z_sup[(b * TO + ta) * 4 + f] = z[ta * 8 + f];
    }
}

// One CTA per sequence (the warp-per-sequence version ran 2.5 k dependent instructions per warp at 6 % occupancy:
// 14 us for a few hundred elements per sequence); every phase is one pass over the CTA's threads.
__global__ void __launch_bounds__(SGB_THREADS) sup_prepare_bwd_kernel(
    SupParams p, int64_t n, const float* __restrict__ zp, const int32_t* __restrict__ idx_in,
    const int32_t* __restrict__ flag_in, const float* __restrict__ std_full, const float* __restrict__ g_z_sup,
    const float* __restrict__ g_z_full, const float* __restrict__ g_std_full, float* __restrict__ g_zp) {
    extern __shared__ float sg_smem[];
    const int lane = threadIdx.x;
    const int64_t b = blockIdx.x;
    (void)n;
    const int T = p.T, O = p.O, TO = T * O;
    float* gf = sg_smem;                 // gradient w.r.t. the smoothed tensor
    float* gm = gf + TO * 8;                                // w.r.t. the matched tensor
    float* sm_ = gm + TO * 8;                               // matched position stds [TO][2]
    float* gzs = sm_ + TO * 2;                              // staged inputs: every global load of the sequence is
    float* gzf = gzs + TO * 4;                              // issued up front (independent, coalesced) instead of
    float* gsf = gzf + TO * 6;                              // one dependent round trip to L2 per use
    float* sfs = gsf + TO * 6;
    float* zs = sfs + TO * 6;                               // raw encoder output of the sequence
    int* idx = reinterpret_cast<int*>(zs + TO * 8);         // [TO] matching permutation, [TO] smoothing flags
    int* flg = idx + TO;
    const float* src = zp + b * TO * 8;
    for (int q = lane; q < TO; q += SGB_THREADS) {
        idx[q] = idx_in[b * TO + q];
        flg[q] = flag_in[b * TO + q];
    }
    for (int q = lane; q < TO * 4; q += SGB_THREADS) gzs[q] = g_z_sup ? __ldg(g_z_sup + b * TO * 4 + q) : 0.f;
    for (int q = lane; q < TO * 6; q += SGB_THREADS) {
        gzf[q] = g_z_full ? __ldg(g_z_full + b * TO * 6 + q) : 0.f;
        gsf[q] = g_std_full ? __ldg(g_std_full + b * TO * 6 + q) : 0.f;
        sfs[q] = __ldg(std_full + b * TO * 6 + q);
    }
    for (int q = lane; q < TO * 8; q += SGB_THREADS) zs[q] = __ldg(src + q);
    __syncthreads();
    // matched position stds (features 6, 7); their smoothed values are recomputed on the fly
    for (int q = lane; q < TO * 2; q += SGB_THREADS) {
        const int ta = q >> 1, d = q & 1, t = ta / O;
        sm_[q] = p.pos_var * sigmoidf_(zs[(t * O + idx[ta]) * 8 + 6 + d]);
    }
    __syncthreads();
    auto fixed_std = [&](int ta, int d) {
        return (flg[ta] & (1 << d)) ? 0.5f * (sm_[(ta - O) * 2 + d] + sm_[(ta + O) * 2 + d]) : sm_[ta * 2 + d];
    };
    // outputs -> smoothed tensor, gather form (each element written once)
    for (int q = lane; q < TO * 8; q += SGB_THREADS) {
        const int ta = q >> 3, f = q & 7, t = ta / O;
        const int o6 = ta * 6;
        float g = 0.f;
        if (f < 4) g += gzs[ta * 4 + f];
        if (t > 0) {
            if (f < 4) g += gzf[o6 + f];
            if (f >= 4) g += gsf[o6 + f - 4];
        }
        if (f == 2 || f == 3) {
            const int d = f - 2;
            if (t > 0) g += gzf[o6 + 4 + d];
            if (t + 1 < T) g -= gzf[o6 + O * 6 + 4 + d];
        }
        if (f >= 6 && g_std_full) {
            const int d = f - 6;
            const float mine = fixed_std(ta, d);
            if (t > 0) g += gsf[o6 + 4 + d] / sfs[o6 + 4 + d] * mine;
            if (t + 1 < T) g += gsf[o6 + O * 6 + 4 + d] / sfs[o6 + O * 6 + 4 + d] * mine;
        }
        gf[q] = g;
    }
    __syncthreads();
    // smoothing select -> matched tensor
    for (int q = lane; q < TO * 8; q += SGB_THREADS) {
        const int ta = q >> 3, f = q & 7, t = ta / O, bit = 1 << (f & 1);
        float g = (flg[ta] & bit) ? 0.f : gf[q];
        if (t + 1 < T && (flg[ta + O] & bit)) g += 0.5f * gf[q + O * 8];
        if (t > 0 && (flg[ta - O] & bit)) g += 0.5f * gf[q - O * 8];
        gm[q] = g;
    }
    __syncthreads();
    // gather -> constrained tensor (`volatile` matching may pick an object twice), then constrain_zp
    for (int q = lane; q < TO * 8; q += SGB_THREADS) {
        const int ta = q >> 3, f = q & 7, t = ta / O, j = ta - t * O;
        float g = 0.f;
        for (int a = 0; a < O; ++a)
            if (idx[t * O + a] == j) g += gm[(t * O + a) * 8 + f];
        const float sc = f == 0 ? p.scale_hi - p.scale_lo
                         : f == 1 ? p.ratio_hi - p.ratio_lo
                         : f < 4 ? 2.f * p.pos_bound
                         : f < 6 ? p.scale_var : p.pos_var;
        const float sgm = sigmoidf_(zs[q]);
        g_zp[b * TO * 8 + q] = g * sc * sgm * (1.f - sgm);
    }
}

static int sup_check(const SupParams& p, int64_t n) {
    STOVE_CHECK_ARG(n >= 0 && p.T >= 2 && p.T <= SG_MAX_T && p.O >= 1 && p.O <= SG_MAX_O, "T or O out of range");
    STOVE_CHECK_ARG(p.match_kind >= 0 && p.match_kind <= 2, "bad match kind");
    STOVE_CHECK_ARG(p.match_kind != 0 || p.O == 3, "3_only matching needs 3 objects");
    return STOVE_OK;
}

static SupParams make_params(const stove_sup_cfg* c) {
    SupParams p;
    p.T = c->T; p.O = c->num_obj; p.match_kind = c->match_kind; p.app_dim = c->app_dim;
    p.match_app = c->match_appearance;
    p.scale_lo = c->min_obj_scale; p.scale_hi = c->max_obj_scale;
    p.ratio_lo = c->min_y_scale; p.ratio_hi = c->max_y_scale;
    p.pos_bound = c->obj_pos_bound; p.scale_var = c->scale_var; p.pos_var = c->pos_var;
    p.fix_eps = c->fix_eps; p.fix = c->fix_supair;
    return p;
}

extern "C" int stove_sup_prepare_fwd(const stove_sup_cfg* cfg, int64_t n, const float* zp, const float* app,
                                     float* z_sup, float* z_full, float* std_full, float* app_out,
                                     int32_t* idx, int32_t* flag, void* stream) {
    STOVE_CHECK_ARG(cfg && zp && z_sup && z_full && std_full && idx && flag, "null pointer");
    SupParams p = make_params(cfg);
    int rc = sup_check(p, n);
    if (rc) return rc;
    STOVE_CHECK_ARG(!(p.match_app && !app), "appearance matching without appearances");
    if (n == 0) return STOVE_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * (size_t)(2 * p.T * p.O * 8 + 2 * p.T * p.O);
    const unsigned blocks = (unsigned)n;
    if (p.O == 3) {
        STOVE_KERNEL(K_SUP_PREPARE_FWD, s, sup_prepare_fwd_kernel<3><<<blocks, SGB_THREADS, smem, s>>>(
            p, n, zp, app, z_sup, z_full, std_full, app_out, idx, flag));
    } else {
        STOVE_KERNEL(K_SUP_PREPARE_FWD, s, sup_prepare_fwd_kernel<0><<<blocks, SGB_THREADS, smem, s>>>(
            p, n, zp, app, z_sup, z_full, std_full, app_out, idx, flag));
    }
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" int stove_sup_prepare_bwd(const stove_sup_cfg* cfg, int64_t n, const float* zp, const int32_t* idx,
                                     const int32_t* flag, const float* std_full, const float* g_z_sup,
                                     const float* g_z_full, const float* g_std_full, float* g_zp, void* stream) {
    STOVE_CHECK_ARG(cfg && zp && idx && flag && std_full && g_zp, "null pointer");
    SupParams p = make_params(cfg);
    int rc = sup_check(p, n);
    if (rc) return rc;
    if (n == 0) return STOVE_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * (size_t)(50 * p.T * p.O);
    STOVE_KERNEL(K_SUP_PREPARE_BWD, s, sup_prepare_bwd_kernel<<<(unsigned)n, SGB_THREADS, smem, s>>>(
        p, n, zp, idx, flag, std_full, g_z_sup, g_z_full, g_std_full, g_zp));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

// ------------------------------------------------------------------------------------
// z of every scored frame (stove.py:731-736): frames t = 1 .. T-1 use the SuPAIR state for
// t < skip and the sampled state z_t for t >= skip; [sx, sy/sx, x, y] -> [sx, sy, x, y]
// (Supair.sy_from_quotient, supair.py:151-158).  Replaces 2 cat + slice + mul + cat + copy
// forward and ~20 ATen launches backward.
//   z_sup [n][T][O][4], z_s [n][S][O][Z] (S = T - skip)  ->  z_all [n][T-1][O][4]
// ------------------------------------------------------------------------------------
__global__ void zall_fwd_kernel(int64_t n, int T, int skip, int O, int Z, const float* __restrict__ z_sup,
                                const float* __restrict__ z_s, float* __restrict__ z_all) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (b, t', o)
    if (i >= n * (T - 1) * O) return;
    const int o = (int)(i % O);
    const int tp = (int)((i / O) % (T - 1));
    const int64_t b = i / ((int64_t)O * (T - 1));
    const int t = tp + 1, S = T - skip;
    const float* src = t < skip ? z_sup + ((b * T + t) * O + o) * 4 : z_s + ((b * S + (t - skip)) * O + o) * Z;
    const float sx = __ldg(src), q = __ldg(src + 1);
    reinterpret_cast<float4*>(z_all)[i] = make_float4(sx, sx * q, __ldg(src + 2), __ldg(src + 3));
}

// g_z_sup and g_z_s are fully written (zeros where z_all does not depend on them)
__global__ void zall_bwd_kernel(int64_t n, int T, int skip, int O, int Z, const float* __restrict__ z_sup,
                                const float* __restrict__ z_s, const float* __restrict__ g_z_all,
                                float* __restrict__ g_z_sup, float* __restrict__ g_z_s) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (b, t, o), t = 0 .. T-1
    if (i >= n * T * O) return;
    const int o = (int)(i % O);
    const int t = (int)((i / O) % T);
    const int64_t b = i / ((int64_t)O * T);
    const int S = T - skip;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 1) {
        const float4 ga = reinterpret_cast<const float4*>(g_z_all)[(b * (T - 1) + (t - 1)) * O + o];
        const float* src = t < skip ? z_sup + ((b * T + t) * O + o) * 4 : z_s + ((b * S + (t - skip)) * O + o) * Z;
        const float sx = __ldg(src), q = __ldg(src + 1);
        g = make_float4(ga.x + ga.y * q, ga.y * sx, ga.z, ga.w);
    }
    if (t < skip) {
        reinterpret_cast<float4*>(g_z_sup)[i] = g;
    } else {
        reinterpret_cast<float4*>(g_z_sup)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        float* dst = g_z_s + ((b * S + (t - skip)) * O + o) * Z;
        dst[0] = g.x; dst[1] = g.y; dst[2] = g.z; dst[3] = g.w;
        for (int k = 4; k < Z; ++k) dst[k] = 0.f;
    }
}

extern "C" int stove_zall_fwd(int64_t n, int T, int skip, int O, int Z, const float* z_sup, const float* z_s,
                              float* z_all, void* stream) {
    STOVE_CHECK_ARG(n >= 0 && T > skip && skip >= 1 && O > 0 && Z >= 4 && z_sup && z_s && z_all, "bad argument");
    const int64_t items = n * (T - 1) * O;
    if (items == 0) return STOVE_OK;
    STOVE_KERNEL(K_ZALL_FWD, (cudaStream_t)stream, zall_fwd_kernel<<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, T, skip, O, Z, z_sup, z_s, z_all));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" int stove_zall_bwd(int64_t n, int T, int skip, int O, int Z, const float* z_sup, const float* z_s,
                              const float* g_z_all, float* g_z_sup, float* g_z_s, void* stream) {
    STOVE_CHECK_ARG(n >= 0 && T > skip && skip >= 1 && O > 0 && Z >= 4 && z_sup && z_s && g_z_all && g_z_sup && g_z_s,
                    "bad argument");
    const int64_t items = n * T * O;
    if (items == 0) return STOVE_OK;
    STOVE_KERNEL(K_ZALL_BWD, (cudaStream_t)stream, zall_bwd_kernel<<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, T, skip, O, Z, z_sup, z_s, g_z_all, g_z_sup, g_z_s));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

// ------------------------------------------------------------------------------------
// ELBO assembly (supair.py:84-110 + stove.py:737-748):
//   lik_f = bg_f + sum_o patch_fo * sx_fo * sy_fo + sum_o (log beta - beta * overlap_fo)
//   elbo  = mean_{t >= skip}(lik_f) + mean(trans - log q) + mean_{1 <= t < skip}(lik_f)
// One CTA, fixed summation order (deterministic), double accumulation.  stats[8] =
// {elbo, mean bg, mean patch, mean overlap (t >= skip), mean log q, mean trans, mean lik_sup, 0}.
// ~40 ATen launches forward + backward become two.
// ------------------------------------------------------------------------------------
#define ELBO_THREADS 1024
__global__ void __launch_bounds__(ELBO_THREADS) elbo_fwd_kernel(int64_t n, int T, int skip, int O, float beta,
                                                                const float* __restrict__ bg,
                                                                const float* __restrict__ patch,
                                                                const float* __restrict__ z_all,
                                                                const float* __restrict__ overlap,
                                                                const float* __restrict__ logq,
                                                                const float* __restrict__ trans,
                                                                float* __restrict__ stats,
                                                                float* __restrict__ elbo_out) {
    __shared__ double red[6][ELBO_THREADS / 32];
    double acc[6] = {0., 0., 0., 0., 0., 0.};      // bg, patch, overlap (elbo frames), lik_sup, logq, trans
    const int S = T - skip, nsup = skip - 1;
    const float logb = logf(beta);
    const int64_t frames = n * (T - 1);
    for (int64_t f = threadIdx.x; f < frames; f += blockDim.x) {
        const int tp = (int)(f % (T - 1));
        float p = 0.f, ov = 0.f;
        for (int o = 0; o < O; ++o) {
            float w = 1.f;                        // z_all == NULL: the patch terms are already weighted by sx * sy
            if (z_all) {
                const float4 z = reinterpret_cast<const float4*>(z_all)[f * O + o];
                w = z.x * z.y;
            }
            p += __ldg(patch + f * O + o) * w;
            ov += logb - beta * __ldg(overlap + f * O + o);
        }
        const float b = __ldg(bg + f);
        if (tp >= nsup) { acc[0] += b; acc[1] += p; acc[2] += ov; }
        else acc[3] += (double)b + p + ov;
    }
    for (int64_t i = threadIdx.x; i < n * S; i += blockDim.x) {
        acc[4] += __ldg(logq + i);
        acc[5] += __ldg(trans + i);
    }
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        double v = acc[q];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[q][wp] = v;
    }
    __syncthreads();
    // second stage by the first warp as a shuffle tree (a serial loop of 6 x 32 double additions in one thread
    // was half of this kernel's 6 us: fp64 latency); the order is fixed, the result deterministic
    double tot[6];
    if (wp == 0) {
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            double v = red[q][lane];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            tot[q] = v;
        }
    }
    if (threadIdx.x == 0) {
        const double ne = (double)n * S, ns = (double)n * nsup;
        const double lik_sup = nsup > 0 ? tot[3] / ns : 0.;
        stats[0] = (float)((tot[0] + tot[1] + tot[2] + tot[5] - tot[4]) / ne + lik_sup);
        *elbo_out = stats[0];
        stats[1] = (float)(tot[0] / ne);
        stats[2] = (float)(tot[1] / ne);
        stats[3] = (float)(tot[2] / ne);
        stats[4] = (float)(tot[4] / ne);
        stats[5] = (float)(tot[5] / ne);
        stats[6] = (float)lik_sup;
        stats[7] = 0.f;
    }
}

// g = d loss / d elbo (device scalar)
__global__ void elbo_bwd_kernel(int64_t n, int T, int skip, int O, float beta, const float* __restrict__ g,
                                const float* __restrict__ patch, const float* __restrict__ z_all,
                                float* __restrict__ g_bg, float* __restrict__ g_patch, float* __restrict__ g_z_all,
                                float* __restrict__ g_overlap, float* __restrict__ g_logq,
                                float* __restrict__ g_trans) {
    const int S = T - skip, nsup = skip - 1;
    const float gv = __ldg(g);
    const float we = gv / (float)((double)n * S), ws = nsup > 0 ? gv / (float)((double)n * nsup) : 0.f;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t items = n * (T - 1) * O;
    if (i < items) {
        const int64_t f = i / O;
        const int tp = (int)(f % (T - 1));
        const float w = tp >= nsup ? we : ws;
        const float4 z = reinterpret_cast<const float4*>(z_all)[i];
        const float p = __ldg(patch + i);
        g_patch[i] = w * z.x * z.y;
        reinterpret_cast<float4*>(g_z_all)[i] = make_float4(w * p * z.y, w * p * z.x, 0.f, 0.f);
        g_overlap[i] = -beta * w;
        if (i % O == 0) g_bg[f] = w;
    }
    if (i < n * S) {
        g_logq[i] = -we;
        g_trans[i] = we;
    }
}

extern "C" int stove_elbo_fwd(int64_t n, int T, int skip, int O, float beta, const float* bg, const float* patch,
                              const float* z_all, const float* overlap, const float* logq, const float* trans,
                              float* stats, float* elbo_out, void* stream) {
    STOVE_CHECK_ARG(n > 0 && T > skip && skip >= 1 && O > 0 && bg && patch && overlap && logq && trans && stats && elbo_out,
                    "bad argument");
    STOVE_KERNEL(K_ELBO_FWD, (cudaStream_t)stream, elbo_fwd_kernel<<<1, ELBO_THREADS, 0, (cudaStream_t)stream>>>(
        n, T, skip, O, beta, bg, patch, z_all, overlap, logq, trans, stats, elbo_out));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" int stove_elbo_bwd(int64_t n, int T, int skip, int O, float beta, const float* g, const float* patch,
                              const float* z_all, float* g_bg, float* g_patch, float* g_z_all, float* g_overlap,
                              float* g_logq, float* g_trans, void* stream) {
    STOVE_CHECK_ARG(n > 0 && T > skip && skip >= 1 && O > 0 && g && patch && z_all && g_bg && g_patch && g_z_all &&
                        g_overlap && g_logq && g_trans, "bad argument");
    int64_t items = n * (T - 1) * O;
    if (n * (T - skip) > items) items = n * (T - skip);
    STOVE_KERNEL(K_ELBO_BWD, (cudaStream_t)stream, elbo_bwd_kernel<<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, T, skip, O, beta, g, patch, z_all, g_bg, g_patch, g_z_all, g_overlap, g_logq, g_trans));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}
