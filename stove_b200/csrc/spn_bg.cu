// Fused background RAT-SPN ("D1" structure), forward and backward, fp32, sm_100a.
//
// Replaces RatSpn.forward (model/spn/rat_torch.py:333-357) for the structure built by
// probabilistic_models.py:25-39: R root partitions, each the product of two Gauss leaf
// vectors that together cover all D = C*W*H inputs; root sum over R*G*G products.
//
// This is the only SPN stage with HBM-like traffic (2 * D floats per frame in, the leaf
// parameters D*R*3*G floats are the larger operand when N is small), so the leaf pass is
// tiled as a skinny GEMM: CTA = (32-pixel chunk) x (128 frames); thread = frame; the chunk's
// parameters are broadcast float4 loads shared by all 128 frames; the frame tile is staged
// transposed in shared memory.  Partial leaf sums per chunk are combined by a per-frame
// root kernel in a fixed order (deterministic).
#include <stdlib.h>
#include "common.cuh"

#define BG_PXC 32      // pixels per chunk
#define BG_FR 128      // frames per CTA

template <int G>
struct GPB_ {
    static constexpr int v = (G + 3) / 4 * 4;
};

// 16-byte asynchronous global -> shared copies (LDGSTS): every copy of a CTA is in flight at once,
// so a parameter block costs one memory round trip instead of one per use (ncu: the leaf kernels
// spent 12-17 cycles per issued instruction waiting on parameter loads from L2).
__device__ __forceinline__ void bg_cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void bg_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void bg_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// stage the leaf parameters and the side table of one pixel chunk: ps [BG_PXC * R * 3 * GP], ss [BG_PXC * R]
template <int R, int GP>
__device__ __forceinline__ void bg_stage_chunk_params(const float* __restrict__ leaf, const int32_t* __restrict__ side,
                                                      int D, int c0, float* ps, int* ss) {
    const int npx = min(BG_PXC, D - c0);
    const int n4 = npx * R * 3 * GP / 4;
    const float4* src = reinterpret_cast<const float4*>(leaf + (int64_t)c0 * R * 3 * GP);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) bg_cp_async16(reinterpret_cast<float4*>(ps) + i, src + i);
    bg_cp_async_commit();
    for (int i = threadIdx.x; i < npx * R; i += blockDim.x) ss[i] = __ldg(side + c0 * R + i);
}

template <int G>
__device__ __forceinline__ void bg_params_smem(const float* lp, float (&mu)[GPB_<G>::v], float (&a)[GPB_<G>::v],
                                               float (&b)[GPB_<G>::v]) {
    constexpr int GP = GPB_<G>::v;
    const float4* p4 = reinterpret_cast<const float4*>(lp);
#pragma unroll
    for (int v = 0; v < GP / 4; ++v) {
        float4 t = p4[v];
        mu[4 * v] = t.x; mu[4 * v + 1] = t.y; mu[4 * v + 2] = t.z; mu[4 * v + 3] = t.w;
        t = p4[GP / 4 + v];
        a[4 * v] = t.x; a[4 * v + 1] = t.y; a[4 * v + 2] = t.z; a[4 * v + 3] = t.w;
        t = p4[2 * (GP / 4) + v];
        b[4 * v] = t.x; b[4 * v + 1] = t.y; b[4 * v + 2] = t.z; b[4 * v + 3] = t.w;
    }
}

template <int G>
__device__ __forceinline__ void bg_load_params(const float* __restrict__ lp, float (&mu)[GPB_<G>::v],
                                               float (&a)[GPB_<G>::v], float (&b)[GPB_<G>::v]) {
    constexpr int GP = GPB_<G>::v;
    const float4* p4 = reinterpret_cast<const float4*>(lp);
#pragma unroll
    for (int v = 0; v < GP / 4; ++v) {
        float4 t = __ldg(p4 + v);
        mu[4 * v] = t.x; mu[4 * v + 1] = t.y; mu[4 * v + 2] = t.z; mu[4 * v + 3] = t.w;
        t = __ldg(p4 + GP / 4 + v);
        a[4 * v] = t.x; a[4 * v + 1] = t.y; a[4 * v + 2] = t.z; a[4 * v + 3] = t.w;
        t = __ldg(p4 + 2 * (GP / 4) + v);
        b[4 * v] = t.x; b[4 * v + 1] = t.y; b[4 * v + 2] = t.z; b[4 * v + 3] = t.w;
    }
}

__device__ __noinline__ float bg_slow_logsumexp(const float* in0, const float* in1, int64_t stride, int G,
                                                const float* wlog) {
    float M = -INFINITY;
    for (int j = 0; j < G; ++j)
        for (int i = 0; i < G; ++i) M = fmaxf(M, in0[i * stride] + in1[j * stride] + wlog[j * G + i]);
    if (!(M > -INFINITY)) return M;
    float acc = 0.f;
    for (int j = 0; j < G; ++j)
        for (int i = 0; i < G; ++i) acc += expf(in0[i * stride] + in1[j * stride] + wlog[j * G + i] - M);
    return M + logf(acc);
}

// stage a [BG_FR frames][BG_PXC pixels] tile transposed: t[p * (BG_FR + 1) + f]
template <bool HAS_MARG, bool KEEP_RAW>
__device__ __forceinline__ void bg_load_tile(const float* __restrict__ x, const float* __restrict__ marg,
                                             int64_t N, int D, int64_t f0, int c0, float* xs, float* ws,
                                             float* ms) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int px = c0 + lane;
#pragma unroll 8
    for (int f = warp; f < BG_FR; f += nwarp) {
        const int64_t n = f0 + f;
        float xv = 0.f, wv = 0.f, mv = 0.f;
        if (n < N && px < D) {
            xv = __ldg(x + n * D + px);
            if (HAS_MARG) {
                mv = __ldg(marg + n * D + px);
                wv = 1.f - fminf(fmaxf(mv, 0.f), 1.f);
            } else {
                wv = 1.f;
            }
        }
        xs[lane * (BG_FR + 1) + f] = xv;
        // KEEP_RAW: the backward needs "mask inside [0, 1]" (derivative of the clamp); it travels in the
        // sign bit of the weight (-0.0f for a saturated mask) instead of a third tile
        ws[lane * (BG_FR + 1) + f] = (KEEP_RAW && !(mv >= 0.f && mv <= 1.f)) ? -wv : wv;
    }
}

// ------------------------------------------------------------------------------------
// forward: leaf pass -> per-chunk partial sums part[(chunk*NI + idx)][npad], NI = R*2*G
// ------------------------------------------------------------------------------------
template <int R, int G, bool HAS_MARG>
__global__ void __launch_bounds__(BG_FR) spn1_fwd_leaf_kernel(
    int D, const int32_t* __restrict__ side, int64_t N, int64_t npad, const float* __restrict__ x,
    const float* __restrict__ marg, const float* __restrict__ leaf, float* __restrict__ part) {
    constexpr int GP = GPB_<G>::v;
    constexpr int NI = R * 2 * G;
    __shared__ float xs[BG_PXC * (BG_FR + 1)];
    __shared__ float ws[BG_PXC * (BG_FR + 1)];
    __shared__ __align__(16) float ps[BG_PXC * R * 3 * GP];
    __shared__ int sides[BG_PXC * R];
    const int chunk = blockIdx.x, c0 = chunk * BG_PXC;
    const int64_t f0 = (int64_t)blockIdx.y * BG_FR;
    bg_stage_chunk_params<R, GP>(leaf, side, D, c0, ps, sides);
    bg_load_tile<HAS_MARG, false>(x, marg, N, D, f0, c0, xs, ws, nullptr);
    bg_cp_async_wait<0>();
    __syncthreads();
    const int f = threadIdx.x;
    float acc[R][2][G];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int g = 0; g < G; ++g) { acc[r][0][g] = 0.f; acc[r][1][g] = 0.f; }
    const int pend = min(BG_PXC, D - c0);
    for (int p = 0; p < pend; ++p) {
        const float xv = xs[p * (BG_FR + 1) + f], wv = ws[p * (BG_FR + 1) + f];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int h = sides[p * R + r];
            float mu[GP], a[GP], b[GP];
            bg_params_smem<G>(ps + (p * R + r) * 3 * GP, mu, a, b);
            if (h) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float d = xv - mu[g];
                    acc[r][1][g] = fmaf(-wv, fmaf(d * d, a[g], b[g]), acc[r][1][g]);
                }
            } else {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float d = xv - mu[g];
                    acc[r][0][g] = fmaf(-wv, fmaf(d * d, a[g], b[g]), acc[r][0][g]);
                }
            }
        }
    }
    const int64_t n = f0 + f;
    if (n < npad) {
        float* dst = part + (int64_t)chunk * NI * npad + n;
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int g = 0; g < G; ++g) dst[(int64_t)((r * 2 + h) * G + g) * npad] = acc[r][h][g];
    }
}

template <int G>
__device__ __forceinline__ float bg_shift_exp(const float* L, float (&e)[G]) {
    float m = L[0];
#pragma unroll
    for (int g = 1; g < G; ++g) m = fmaxf(m, L[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) e[g] = expf(L[g] - m);
    return m;
}

// forward: combine chunks, products and root sum.  CTA = 32 frames x NI/3 thread groups: the
// chunk partials are summed by 384 threads (a frame-per-thread loop over 32 chunks x 36 values was
// latency bound: 38 us for 1536 frames), then one thread per frame does the 3 products + root sum.
#define BG_ROOT_FR 32
template <int R, int G>
__global__ void __launch_bounds__(BG_ROOT_FR * (R * 2 * G / 3)) spn1_fwd_root_kernel(
    int nchunks, int64_t N, int64_t ppad, int64_t npad, const float* __restrict__ part,
    const float* __restrict__ rlin, const float* __restrict__ rlog, float* __restrict__ leaf_val,
    float* __restrict__ out) {
    constexpr int NI = R * 2 * G;
    __shared__ float Ls[NI][BG_ROOT_FR + 1];
    const int fl = threadIdx.x % BG_ROOT_FR, grp = threadIdx.x / BG_ROOT_FR;
    const int64_t n = (int64_t)blockIdx.x * BG_ROOT_FR + fl;
    if (n < N) {
        float acc[3] = {0.f, 0.f, 0.f};
        for (int c = 0; c < nchunks; ++c) {          // fixed order: deterministic
            const float* src = part + ((int64_t)c * NI + grp * 3) * ppad + n;
#pragma unroll
            for (int i = 0; i < 3; ++i) acc[i] += src[(int64_t)i * ppad];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            Ls[grp * 3 + i][fl] = acc[i];
            leaf_val[(int64_t)(grp * 3 + i) * npad + n] = acc[i];
        }
    }
    __syncthreads();
    if (grp != 0 || n >= N) return;
    float L[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) L[i] = Ls[i][fl];
    float vals[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float eA[G], eB[G];
        const float mA = bg_shift_exp<G>(&L[(r * 2) * G], eA), mB = bg_shift_exp<G>(&L[(r * 2 + 1) * G], eB);
        float U = 0.f;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            float inner = 0.f;
#pragma unroll
            for (int i = 0; i < G; ++i) inner = fmaf(eA[i], __ldg(rlin + r * G * G + j * G + i), inner);
            U = fmaf(eB[j], inner, U);
        }
        if (U > LIN_SUM_FLOOR) {
            vals[r] = mA + mB + logf(U);
        } else {
            // exact log-domain value from the shared copy (leaf_val was written by other threads)
            float M = -INFINITY;
            for (int j = 0; j < G; ++j)
                for (int i = 0; i < G; ++i)
                    M = fmaxf(M, Ls[(r * 2) * G + i][fl] + Ls[(r * 2 + 1) * G + j][fl] + rlog[r * G * G + j * G + i]);
            float acc = 0.f;
            if (M > -INFINITY)
                for (int j = 0; j < G; ++j)
                    for (int i = 0; i < G; ++i)
                        acc += expf(Ls[(r * 2) * G + i][fl] + Ls[(r * 2 + 1) * G + j][fl] + rlog[r * G * G + j * G + i] - M);
            vals[r] = (M > -INFINITY) ? M + logf(acc) : M;
        }
    }
    float M = vals[0];
#pragma unroll
    for (int r = 1; r < R; ++r) M = fmaxf(M, vals[r]);
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) acc += expf(vals[r] - M);
    out[n] = (M > -INFINITY) ? M + logf(acc) : M;
}

// ------------------------------------------------------------------------------------
// backward 1/4: root -> leaf-vector gradients; thread = frame
// ------------------------------------------------------------------------------------
template <int R, int G>
__global__ void __launch_bounds__(128) spn1_bwd_root_kernel(
    int64_t N, int64_t npad, const float* __restrict__ rlin, const float* __restrict__ rlog,
    const float* __restrict__ leaf_val, const float* __restrict__ out, const float* __restrict__ g_out,
    float* __restrict__ gleaf, float* __restrict__ aux_root, float* __restrict__ g_rlog) {
    constexpr int NI = R * 2 * G;
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= npad) return;
    if (n >= N) {
        for (int i = 0; i < NI; ++i) gleaf[(int64_t)i * npad + n] = 0.f;
        for (int i = 0; i < R * (1 + 2 * G); ++i) aux_root[(int64_t)i * npad + n] = 0.f;
        return;
    }
    const float go = g_out[n], ov = out[n];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float A[G], B[G], eA[G], eB[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            A[g] = leaf_val[(int64_t)((r * 2) * G + g) * npad + n];
            B[g] = leaf_val[(int64_t)((r * 2 + 1) * G + g) * npad + n];
        }
        const float mA = bg_shift_exp<G>(A, eA), mB = bg_shift_exp<G>(B, eB);
        float colA[G], rowB[G];
#pragma unroll
        for (int i = 0; i < G; ++i) colA[i] = 0.f;
        float U = 0.f;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            float inner = 0.f;
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const float w = __ldg(rlin + r * G * G + j * G + i);
                inner = fmaf(eA[i], w, inner);
                colA[i] = fmaf(eB[j], w, colA[i]);
            }
            rowB[j] = inner;
            U = fmaf(eB[j], inner, U);
        }
        float* ga = gleaf + (int64_t)((r * 2) * G) * npad + n;
        float* gb = gleaf + (int64_t)((r * 2 + 1) * G) * npad + n;
        float* ar = aux_root + (int64_t)r * (1 + 2 * G) * npad + n;
        if (U > LIN_SUM_FLOOR) {
            const float val = mA + mB + logf(U);
            const float c = go * expf(val - ov) / U;
            ar[0] = c;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                ga[(int64_t)g * npad] = c * eA[g] * colA[g];
                gb[(int64_t)g * npad] = c * eB[g] * rowB[g];
                ar[(int64_t)(1 + g) * npad] = eA[g];
                ar[(int64_t)(1 + G + g) * npad] = eB[g];
            }
        } else {
            const float* a0 = leaf_val + (int64_t)((r * 2) * G) * npad + n;
            const float* b0 = leaf_val + (int64_t)((r * 2 + 1) * G) * npad + n;
            const float val = bg_slow_logsumexp(a0, b0, npad, G, rlog + r * G * G);
            const float gr = go * expf(val - ov);
            ar[0] = 0.f;
            for (int g = 0; g < G; ++g) {
                ga[(int64_t)g * npad] = 0.f;
                gb[(int64_t)g * npad] = 0.f;
                ar[(int64_t)(1 + g) * npad] = 0.f;
                ar[(int64_t)(1 + G + g) * npad] = 0.f;
            }
            if (gr != 0.f && val > -INFINITY) {
                for (int j = 0; j < G; ++j)
                    for (int i = 0; i < G; ++i) {
                        const float resp = gr * expf(a0[i * npad] + b0[j * npad] + rlog[r * G * G + j * G + i] - val);
                        ga[(int64_t)i * npad] += resp;
                        gb[(int64_t)j * npad] += resp;
                        atomicAdd(g_rlog + r * G * G + j * G + i, resp);
                    }
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// backward 2/4: gradient w.r.t. the marginalisation mask (and optionally the input);
// same tiling as the forward leaf pass, results written back through the staged tile.
// ------------------------------------------------------------------------------------
template <int R, int G, bool HAS_MARG>
__global__ void __launch_bounds__(BG_FR) spn1_bwd_input_kernel(
    int D, const int32_t* __restrict__ side, int64_t N, int64_t npad, const float* __restrict__ x,
    const float* __restrict__ marg, const float* __restrict__ leaf, const float* __restrict__ gleaf,
    float* __restrict__ g_x, float* __restrict__ g_marg) {
    constexpr int GP = GPB_<G>::v;
    extern __shared__ float smem[];
    float* xs = smem;
    float* ws = xs + BG_PXC * (BG_FR + 1);
    float* ps = ws + BG_PXC * (BG_FR + 1);                  // 16-byte aligned: 2 * 32 * 129 floats precede it
    int* sides = reinterpret_cast<int*>(ps + BG_PXC * R * 3 * GP);
    const int chunk = blockIdx.x, c0 = chunk * BG_PXC;
    const int64_t f0 = (int64_t)blockIdx.y * BG_FR;
    bg_stage_chunk_params<R, GP>(leaf, side, D, c0, ps, sides);
    bg_load_tile<HAS_MARG, true>(x, marg, N, D, f0, c0, xs, ws, nullptr);
    bg_cp_async_wait<0>();
    __syncthreads();
    const int f = threadIdx.x;
    const int64_t n = f0 + f;
    float gl[R][2][G];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int g = 0; g < G; ++g)
                gl[r][h][g] = (n < N) ? gleaf[(int64_t)((r * 2 + h) * G + g) * npad + n] : 0.f;
    const int pend = min(BG_PXC, D - c0);
    for (int p = 0; p < pend; ++p) {
        const float xv = xs[p * (BG_FR + 1) + f], wraw = ws[p * (BG_FR + 1) + f];
        const float wv = fabsf(wraw);
        const bool inside = !signbit(wraw);
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int h = sides[p * R + r];
            float mu[GP], a[GP], b[GP];
            bg_params_smem<G>(ps + (p * R + r) * 3 * GP, mu, a, b);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float d = xv - mu[g];
                const float ad = a[g] * d;
                const float gv = h ? gl[r][1][g] : gl[r][0][g];
                t1 = fmaf(gv, ad, t1);
                t2 = fmaf(gv, fmaf(ad, d, b[g]), t2);
            }
        }
        xs[p * (BG_FR + 1) + f] = -2.f * wv * t1;
        ws[p * (BG_FR + 1) + f] = inside ? t2 : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int px = c0 + lane;
    for (int ff = warp; ff < BG_FR; ff += nwarp) {
        const int64_t nn = f0 + ff;
        if (nn < N && px < D) {
            if (g_x) g_x[nn * D + px] = xs[lane * (BG_FR + 1) + ff];
            if (HAS_MARG && g_marg) g_marg[nn * D + px] = ws[lane * (BG_FR + 1) + ff];
        }
    }
}

// ------------------------------------------------------------------------------------
// backward 3/4: leaf parameter gradients.  thread = pixel (its R*G parameters and their
// accumulators stay in registers), CTA = 128 pixels x one chunk of frames.
// ------------------------------------------------------------------------------------
template <int R, int G, bool HAS_MARG>
__global__ void __launch_bounds__(128) spn1_bwd_leafparam_kernel(
    int D, const int32_t* __restrict__ side, int64_t N, int64_t npad, int chunk,
    const float* __restrict__ x, const float* __restrict__ marg, const float* __restrict__ leaf,
    const float* __restrict__ gleaf, float* __restrict__ g_leaf) {
    constexpr int GP = GPB_<G>::v;
    constexpr int NI = R * 2 * G;
    __shared__ float gls[NI * 33];
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = px < D;
    const int pxc = active ? px : 0;
    float mu[R][G], a[R][G];
    int hoff[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float* lp = leaf + ((int64_t)pxc * R + r) * 3 * GP;
#pragma unroll
        for (int g = 0; g < G; ++g) { mu[r][g] = lp[g]; a[r][g] = lp[GP + g]; }
        hoff[r] = ((r * 2 + side[pxc * R + r]) * G) * 33;
    }
    float s1[R][G], s2[R][G], s3[R][G];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int g = 0; g < G; ++g) { s1[r][g] = 0.f; s2[r][g] = 0.f; s3[r][g] = 0.f; }
    const int64_t c0 = (int64_t)blockIdx.y * chunk;
    const int64_t c1 = min(c0 + (int64_t)chunk, N);
    for (int64_t base = c0; base < c1; base += 32) {
        for (int idx = threadIdx.x; idx < NI * 32; idx += blockDim.x) {
            const int row = idx >> 5, pt = idx & 31;
            const int64_t n = base + pt;
            gls[row * 33 + pt] = (n < c1) ? gleaf[(int64_t)row * npad + n] : 0.f;
        }
        __syncthreads();
        if (active) {
            const int lim = (int)min((int64_t)32, c1 - base);
            for (int pt = 0; pt < lim; ++pt) {
                const int64_t n = base + pt;
                const float xv = __ldg(x + n * D + px);
                const float wv = HAS_MARG ? 1.f - fminf(fmaxf(__ldg(marg + n * D + px), 0.f), 1.f) : 1.f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float d = xv - mu[r][g];
                        const float gw = gls[hoff[r] + g * 33 + pt] * wv;
                        s1[r][g] = fmaf(gw, d, s1[r][g]);
                        s2[r][g] = fmaf(gw * d, d, s2[r][g]);
                        s3[r][g] += gw;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (active) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float* dst = g_leaf + ((int64_t)px * R + r) * 3 * GP;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                atomicAdd(dst + g, 2.f * a[r][g] * s1[r][g]);
                atomicAdd(dst + GP + g, -s2[r][g]);
                atomicAdd(dst + 2 * GP + g, -s3[r][g]);
            }
        }
    }
}

// Same mapping, but the [32 frames][128 pixels] frame / mask tiles and the leaf-vector gradients
// stream through a two-stage cp.async pipeline (needs D % 4 == 0 for 16-byte copies).  The scalar
// version above waits for two dependent global loads per frame with 4 warps per SM (ncu: 6 %
// occupancy, 6.3 stall cycles per instruction on the load scoreboard).
template <int R, int G, bool HAS_MARG>
__global__ void __launch_bounds__(128) spn1_bwd_leafparam_async_kernel(
    int D, const int32_t* __restrict__ side, int64_t N, int64_t npad, int chunk,
    const float* __restrict__ x, const float* __restrict__ marg, const float* __restrict__ leaf,
    const float* __restrict__ gleaf, float* __restrict__ g_leaf) {
    constexpr int GP = GPB_<G>::v;
    constexpr int NI = R * 2 * G;
    constexpr int TILE = 32 * 128;
    // rows of the leaf-vector gradients GT_LD = 36 floats apart: the threads of a warp read two different rows (the
    // two sides of a split) at the same frame index; with a stride of 32 they share a bank (ncu: 47 % of the
    // shared-memory wavefronts were conflicts).  36 keeps cp.async's 16-byte alignment.
    constexpr int GT_LD = 36;
    extern __shared__ __align__(16) float smem[];
    float* xt = smem;                       // [2][32][128]
    float* mt = xt + 2 * TILE;              // [2][32][128]
    float* gt = mt + 2 * TILE;              // [2][NI][GT_LD]
    const int tid = threadIdx.x;
    const int px0 = blockIdx.x * 128, px = px0 + tid;
    const bool active = px < D;
    const int pxc = active ? px : 0;
    float mu[R][G], a[R][G];
    int hoff[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float* lp = leaf + ((int64_t)pxc * R + r) * 3 * GP;
#pragma unroll
        for (int g = 0; g < G; ++g) { mu[r][g] = lp[g]; a[r][g] = lp[GP + g]; }
        hoff[r] = ((r * 2 + side[pxc * R + r]) * G) * GT_LD;
    }
    float s1[R][G], s2[R][G], s3[R][G];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int g = 0; g < G; ++g) { s1[r][g] = 0.f; s2[r][g] = 0.f; s3[r][g] = 0.f; }
    const int64_t c0 = (int64_t)blockIdx.y * chunk;
    const int64_t c1 = min(c0 + (int64_t)chunk, N);
    const int nb = (int)((c1 - c0 + 31) / 32);
    const int cols4 = min(128, D - px0) / 4;          // float4 columns of this CTA's pixel slab
    auto issue = [&](int b) {
        const int st = b & 1;
        const int64_t base = c0 + (int64_t)b * 32;
        const int rows = (int)min((int64_t)32, c1 - base);
        for (int i = tid; i < rows * 32; i += 128) {
            const int pt = i >> 5, c4 = i & 31;
            if (c4 < cols4) {
                bg_cp_async16(xt + st * TILE + pt * 128 + c4 * 4, x + (base + pt) * D + px0 + c4 * 4);
                if (HAS_MARG) bg_cp_async16(mt + st * TILE + pt * 128 + c4 * 4, marg + (base + pt) * D + px0 + c4 * 4);
            }
        }
        // gradients of the leaf vectors: rows of 32 frames (npad is a multiple of 32, so rows past c1 exist)
        for (int i = tid; i < NI * 8; i += 128) {
            const int row = i >> 3, c4 = i & 7;
            bg_cp_async16(gt + st * NI * GT_LD + row * GT_LD + c4 * 4, gleaf + (int64_t)row * npad + base + c4 * 4);
        }
        bg_cp_async_commit();
    };
    if (nb > 0) issue(0);
    for (int b = 0; b < nb; ++b) {
        if (b + 1 < nb) {
            issue(b + 1);
            bg_cp_async_wait<1>();
        } else {
            bg_cp_async_wait<0>();
        }
        __syncthreads();
        if (active) {
            const int st = b & 1;
            const int lim = (int)min((int64_t)32, c1 - (c0 + (int64_t)b * 32));
            const float* xs = xt + st * TILE + tid;
            const float* ms = mt + st * TILE + tid;
            const float* gs = gt + st * NI * GT_LD;
#pragma unroll 2
            for (int pt = 0; pt < lim; ++pt) {
                const float xv = xs[pt * 128];
                const float wv = HAS_MARG ? 1.f - fminf(fmaxf(ms[pt * 128], 0.f), 1.f) : 1.f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float d = xv - mu[r][g];
                        const float gw = gs[hoff[r] + g * GT_LD + pt] * wv;
                        s1[r][g] = fmaf(gw, d, s1[r][g]);
                        s2[r][g] = fmaf(gw * d, d, s2[r][g]);
                        s3[r][g] += gw;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (active) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float* dst = g_leaf + ((int64_t)px * R + r) * 3 * GP;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                atomicAdd(dst + g, 2.f * a[r][g] * s1[r][g]);
                atomicAdd(dst + GP + g, -s2[r][g]);
                atomicAdd(dst + 2 * GP + g, -s3[r][g]);
            }
        }
    }
}

// backward 4/4: root log-weight gradients, G_rlog[r][k] += W[k] * sum_n c eA[i] eB[j]
template <int R, int G>
__global__ void __launch_bounds__(64) spn1_bwd_rootparam_kernel(
    int64_t N, int64_t npad, int chunk, const float* __restrict__ rlin, const float* __restrict__ aux_root,
    float* __restrict__ g_rlog) {
    __shared__ float tile[(1 + 2 * G) * 33];
    const int r = blockIdx.x, tid = threadIdx.x;
    const bool active = tid < G * G;
    const int i = active ? tid % G : 0, j = active ? tid / G : 0;
    const int64_t c0 = (int64_t)blockIdx.y * chunk;
    const int64_t c1 = min(c0 + (int64_t)chunk, N);
    float acc = 0.f;
    for (int64_t base = c0; base < c1; base += 32) {
        for (int idx = tid; idx < (1 + 2 * G) * 32; idx += blockDim.x) {
            const int row = idx >> 5, pt = idx & 31;
            const int64_t n = base + pt;
            tile[row * 33 + pt] = (n < c1) ? aux_root[(int64_t)(r * (1 + 2 * G) + row) * npad + n] : 0.f;
        }
        __syncthreads();
        if (active)
            for (int pt = 0; pt < 32; ++pt)
                acc = fmaf(tile[pt] * tile[(1 + i) * 33 + pt], tile[(1 + G + j) * 33 + pt], acc);
        __syncthreads();
    }
    if (active) atomicAdd(g_rlog + r * G * G + tid, acc * rlin[r * G * G + tid]);
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
static int check1(const stove_spn1_struct* st) {
    STOVE_CHECK_ARG(st && st->side && st->D > 0, "bad structure");
    if (!(st->R == 3 && st->G == 6)) {
        stove_set_error("spn1: (R, G) = (%d, %d) is not instantiated (only (3, 6))", st->R, st->G);
        return STOVE_ERR_UNSUPPORTED;
    }
    return STOVE_OK;
}

static inline int bg_nchunks(int D) { return (D + BG_PXC - 1) / BG_PXC; }

// the chunk partials use stride ppad = round_up(N, BG_FR) so every thread of the leaf
// kernel may store; everything else uses npad = round_up(N, 32).
extern "C" size_t stove_spn1_fwd_workspace(const stove_spn1_struct* st, int64_t N) {
    if (!st || N <= 0) return 0;
    return sizeof(float) * (size_t)bg_nchunks(st->D) * st->R * 2 * st->G * (size_t)round_up64(N, BG_FR);
}

extern "C" int stove_spn1_fwd(const stove_spn1_struct* st, int64_t N, const float* x, const float* marg,
                              const float* leaf, const float* rlin, const float* rlog, float* leaf_val,
                              float* out, void* workspace, void* stream) {
    int rc = check1(st);
    if (rc) return rc;
    STOVE_CHECK_ARG(N >= 0 && x && leaf && rlin && rlog && leaf_val && out && workspace, "null pointer");
    if (N == 0) return STOVE_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t npad = round_up64(N, 32), ppad = round_up64(N, BG_FR);
    const int nch = bg_nchunks(st->D);
    dim3 grid(nch, (unsigned)(ppad / BG_FR));
    float* part = (float*)workspace;
    if (marg)
        STOVE_KERNEL(K_SPN1_FWD_LEAF, s, spn1_fwd_leaf_kernel<3, 6, true><<<grid, BG_FR, 0, s>>>(st->D, st->side, N, ppad, x, marg, leaf, part));
    else
        STOVE_KERNEL(K_SPN1_FWD_LEAF, s, spn1_fwd_leaf_kernel<3, 6, false><<<grid, BG_FR, 0, s>>>(st->D, st->side, N, ppad, x, marg, leaf, part));
    STOVE_LAUNCH_CHECK();
    STOVE_KERNEL(K_SPN1_FWD_ROOT, s, spn1_fwd_root_kernel<3, 6><<<(unsigned)((N + BG_ROOT_FR - 1) / BG_ROOT_FR), BG_ROOT_FR * 12, 0, s>>>(nch, N, ppad, npad, part, rlin, rlog,
                                                                        leaf_val, out));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

struct Spn1Ws {
    float* gleaf;
    float* aux_root;
    size_t bytes;
};

static Spn1Ws spn1_ws_layout(const stove_spn1_struct* st, int64_t N, void* base) {
    const int64_t npad = round_up64(N, 32);
    Spn1Ws w;
    float* p = (float*)base;
    w.gleaf = p; p += (int64_t)st->R * 2 * st->G * npad;
    w.aux_root = p; p += (int64_t)st->R * (1 + 2 * st->G) * npad;
    w.bytes = (size_t)((char*)p - (char*)base);
    return w;
}

// the two parameter-gradient kernels of the backward pass on a workspace filled by the root pass (spn1_bwd_root_kernel
// here, or the fused chain kernel of scene_ll_bwd.cu)
int spn1_param_kernels(const stove_spn1_struct* st, int64_t N, const float* x, const float* marg, const float* leaf,
                       const float* rlin, void* workspace, float* g_leaf, float* g_rlog, cudaStream_t s_leaf,
                       cudaStream_t s_root) {
    const int D = st->D;
    const int64_t npad = round_up64(N, 32);
    Spn1Ws w = spn1_ws_layout(st, N, workspace);
    int chunk = (int)round_up64((N + 23) / 24, 32);
    if (chunk < 32) chunk = 32;
    const int nchunk = (int)((N + chunk - 1) / chunk);
    {
        // leaf-parameter kernel: CTAs of 4 warps holding 66 KB of shared memory each.  With the chunking above the
        // grid is ~150 CTAs = one per SM, 6 % of the warp slots (ncu) -- it runs at latency, not throughput.  A
        // finer frame chunk (two 32-frame tiles per CTA, still pipelined) gives ~3 CTAs per SM.
        int lchunk = chunk;
        const int ctas_px = (D + 127) / 128;
        while (lchunk > 64 && (int64_t)ctas_px * ((N + lchunk - 1) / lchunk) < 3 * 148) lchunk = round_up(lchunk / 2, 32);
        const int lnchunk = (int)((N + lchunk - 1) / lchunk);
        dim3 grid(ctas_px, lnchunk);
        const int chunk_saved = chunk;
        chunk = lchunk;
        if (D % 4 == 0) {
            const size_t smem = sizeof(float) * (4 * 32 * 128 + 2 * 3 * 2 * 6 * 36);
            if (marg) {
                STOVE_CUDA(cudaFuncSetAttribute(spn1_bwd_leafparam_async_kernel<3, 6, true>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                STOVE_KERNEL(K_SPN1_BWD_LEAFPARAM, s_leaf, spn1_bwd_leafparam_async_kernel<3, 6, true><<<grid, 128, smem, s_leaf>>>(D, st->side, N, npad, chunk, x, marg, leaf, w.gleaf, g_leaf));
            } else {
                STOVE_CUDA(cudaFuncSetAttribute(spn1_bwd_leafparam_async_kernel<3, 6, false>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                STOVE_KERNEL(K_SPN1_BWD_LEAFPARAM, s_leaf, spn1_bwd_leafparam_async_kernel<3, 6, false><<<grid, 128, smem, s_leaf>>>(D, st->side, N, npad, chunk, x, marg, leaf, w.gleaf, g_leaf));
            }
        } else if (marg)
            STOVE_KERNEL(K_SPN1_BWD_LEAFPARAM, s_leaf, spn1_bwd_leafparam_kernel<3, 6, true><<<grid, 128, 0, s_leaf>>>(D, st->side, N, npad, chunk, x, marg, leaf, w.gleaf, g_leaf));
        else
            STOVE_KERNEL(K_SPN1_BWD_LEAFPARAM, s_leaf, spn1_bwd_leafparam_kernel<3, 6, false><<<grid, 128, 0, s_leaf>>>(D, st->side, N, npad, chunk, x, marg, leaf, w.gleaf, g_leaf));
        STOVE_LAUNCH_CHECK();
        chunk = chunk_saved;
    }
    {
        dim3 grid(st->R, nchunk);
        STOVE_KERNEL(K_SPN1_BWD_ROOTPARAM, s_root, spn1_bwd_rootparam_kernel<3, 6><<<grid, 64, 0, s_root>>>(N, npad, chunk, rlin, w.aux_root, g_rlog));
        STOVE_LAUNCH_CHECK();
    }
    return STOVE_OK;
}

extern "C" size_t stove_spn1_bwd_workspace(const stove_spn1_struct* st, int64_t N) {
    if (!st || N <= 0) return 0;
    return spn1_ws_layout(st, N, nullptr).bytes;
}

extern "C" int stove_spn1_bwd(const stove_spn1_struct* st, int64_t N, const float* x, const float* marg,
                              const float* leaf, const float* rlin, const float* rlog,
                              const float* leaf_val, const float* out, const float* g_out, float* g_x,
                              float* g_marg, float* g_leaf, float* g_rlog, void* workspace, void* stream,
                              void* join_stream) {
    int rc = check1(st);
    if (rc) return rc;
    STOVE_CHECK_ARG(N >= 0 && x && leaf && rlin && rlog && leaf_val && out && g_out && g_leaf && g_rlog && workspace,
                    "null pointer");
    STOVE_CHECK_ARG(!(g_marg && !marg), "g_marg requested without marg");
    if (N == 0) return STOVE_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int D = st->D;
    const int64_t npad = round_up64(N, 32);
    Spn1Ws w = spn1_ws_layout(st, N, workspace);
    STOVE_KERNEL(K_SPN1_BWD_ROOT, s, spn1_bwd_root_kernel<3, 6><<<(unsigned)((npad + 127) / 128), 128, 0, s>>>(N, npad, rlin, rlog, leaf_val, out,
                                                                           g_out, w.gleaf, w.aux_root, g_rlog));
    STOVE_LAUNCH_CHECK();
    StoveFork* fk = !stove_opt(OPT_FORK) ? nullptr : stove_fork_get(1);
    if (fk && (rc = stove_fork(fk, s, 2))) return rc;
    cudaStream_t s_leaf = fk ? fk->side[0] : s, s_root = fk ? fk->side[1] : s;
    if (g_x || g_marg) {
        const size_t smem = sizeof(float) * (2 * BG_PXC * (BG_FR + 1) + BG_PXC * 3 * 3 * 8 + BG_PXC * 3);
        dim3 grid(bg_nchunks(D), (unsigned)((N + BG_FR - 1) / BG_FR));
        if (marg) {
            STOVE_CUDA(cudaFuncSetAttribute(spn1_bwd_input_kernel<3, 6, true>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            STOVE_KERNEL(K_SPN1_BWD_INPUT, s, spn1_bwd_input_kernel<3, 6, true><<<grid, BG_FR, smem, s>>>(D, st->side, N, npad, x, marg, leaf, w.gleaf, g_x, g_marg));
        } else {
            STOVE_CUDA(cudaFuncSetAttribute(spn1_bwd_input_kernel<3, 6, false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            STOVE_KERNEL(K_SPN1_BWD_INPUT, s, spn1_bwd_input_kernel<3, 6, false><<<grid, BG_FR, smem, s>>>(D, st->side, N, npad, x, marg, leaf, w.gleaf, g_x, g_marg));
        }
        STOVE_LAUNCH_CHECK();
    }
    if ((rc = spn1_param_kernels(st, N, x, marg, leaf, rlin, workspace, g_leaf, g_rlog, s_leaf, s_root))) return rc;
    if (fk && (rc = stove_join(fk, join_stream ? (cudaStream_t)join_stream : s, 2))) return rc;   // see spn_obj.cu
    return STOVE_OK;
}
