// Fused object RAT-SPN ("D2" structure), forward and backward, fp32, sm_100a.
//
// Replaces RatSpn.forward (model/spn/rat_torch.py:333-357) for the structure built by
// probabilistic_models.py:8-22: R root partitions; each is the product (rat_torch.py:147-163)
// of two mid regions; each mid region is a sum vector (:202-222) over the product of two
// Gauss leaf vectors (:83-109).  The reference evaluates ~2000 ATen ops and materialises
// (N, 25, 10) x 24 and (N, 100, 10) x 12 temporaries; here one CTA owns 32 patches, one warp
// per mid region, and nothing but the 100-float patch row is read from HBM/L2.
//
// Mapping: lane = patch, warp = region q (2R warps).  Leaf parameters are read as broadcast
// float4 loads (every lane of a warp needs the same (mu, a, b)); the patch tile lives in
// shared memory transposed ([pixel][33]) so lane-strided reads are conflict free.
// Sum layers use the max-shifted linear-domain form
//     out[s] = m0 + m1 + log sum_{i,j} e0[i] e1[j] W[j*G+i][s],   e = exp(leaf - max)
// (20 exp + 10 log instead of 1000 exp); if the linear sum falls below LIN_SUM_FLOOR the
// value/gradient is recomputed exactly in the log domain (rare slow path).
#include <stdlib.h>
#include "common.cuh"
#include "spn_math.cuh"

template <int G, bool SMEM>
__device__ __forceinline__ void leaf_accumulate(float (&L)[G], const float* __restrict__ leaf_q,
                                                const int32_t* __restrict__ sc, int p_begin, int p_end,
                                                const float* xs, const float* ws, int lane) {
    constexpr int GP = GP_<G>::v;
#pragma unroll 2
    for (int p = p_begin; p < p_end; ++p) {
        const int px = SMEM ? sc[p] : __ldg(sc + p);
        const float xv = xs[px * 33 + lane];
        const float wv = ws[px * 33 + lane];
        float mu[GP], a[GP], b[GP];
        load_leaf_params<G, SMEM>(leaf_q + (int64_t)p * 3 * GP, mu, a, b);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float d = xv - mu[g];
            const float u = fmaf(d * d, a[g], b[g]);
            L[g] = fmaf(-wv, u, L[g]);
        }
    }
}

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
// A CTA owns `tpc` tiles of 32 patches (tpc * 2R warps): with one tile per CTA and the leaf table
// staged (1 CTA per SM) 168 tiles ran as 148 + 20, i.e. two full CTA latencies; two tiles per CTA
// share one staged table and finish in one wave.
template <int G, int S, bool HAS_MARG, bool STAGE>
__global__ void __launch_bounds__(1024) spn2_fwd_kernel(
    Spn2Dev st, int64_t N, int64_t npad, const float* __restrict__ x, const float* __restrict__ marg,
    const float* __restrict__ leaf, const float* __restrict__ wlin, const float* __restrict__ wlog,
    const float* __restrict__ rlin, const float* __restrict__ rlog, float* __restrict__ leaf_val,
    float* __restrict__ sum_val, float* __restrict__ out) {
    constexpr int GP = GP_<G>::v;
    constexpr int SP = GP_<S>::v;
    extern __shared__ __align__(16) float smem[];
    const int D = st.D, R = st.R, Q = 2 * st.R;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tpc = (blockDim.x >> 5) / Q, tile = warp / Q, q = warp - tile * Q;
    const int per_tile = 2 * up4(D * 33) + Q * S * 32 + R * 32;
    float* xs = smem + tile * per_tile;
    float* ws = xs + up4(D * 33);
    float* ss = ws + up4(D * 33);     // [Q*S][32]
    float* vr = ss + Q * S * 32;      // [R][32]
    float* lf = smem + tpc * per_tile;                                // STAGE: [Q][pmax*3*GP] leaf parameters
    int32_t* scs = reinterpret_cast<int32_t*>(lf + Q * st.pmax * 3 * GP);   // STAGE: [Q][pmax] scope indices
    const int64_t base = ((int64_t)blockIdx.x * tpc + tile) * 32;
    const bool tile_ok = base < npad;

    if (STAGE && tile == 0) {
        // each warp of the first tile fetches the parameter block of its region; in flight during the tile load
        const float4* src = reinterpret_cast<const float4*>(leaf + (int64_t)q * st.pmax * 3 * GP);
        float4* dst = reinterpret_cast<float4*>(lf + q * st.pmax * 3 * GP);
        for (int i = lane; i < st.pmax * 3 * GP / 4; i += 32) cp_async16(dst + i, src + i);
        cp_async_commit();
        for (int i = lane; i < st.pmax; i += 32) scs[q * st.pmax + i] = __ldg(st.scope + q * st.pmax + i);
    }
    if (tile_ok) {
        const int ttid = tid - tile * Q * 32, tthreads = Q * 32;
#pragma unroll 4
        for (int idx = ttid; idx < 32 * D; idx += tthreads) {
            const int pt = idx / D, px = idx - pt * D;
            const int64_t n = base + pt;
            float xv = 0.f, wv = 1.f;
            if (n < N) {
                xv = __ldg(x + n * D + px);
                if (HAS_MARG) wv = 1.f - fminf(fmaxf(__ldg(marg + n * D + px), 0.f), 1.f);
            }
            xs[px * 33 + pt] = xv;
            ws[px * 33 + pt] = wv;
        }
    }
    if (STAGE) cp_async_wait<0>();
    __syncthreads();

    const int64_t n = base + lane;   // < npad when tile_ok
    float L0[G], L1[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { L0[g] = 0.f; L1[g] = 0.f; }
    if (tile_ok) {
        const int nq0 = __ldg(st.n0 + q), nq = __ldg(st.nt + q);
        const int32_t* sc = STAGE ? scs + q * st.pmax : st.scope + q * st.pmax;
        const float* leaf_q = STAGE ? lf + q * st.pmax * 3 * GP : leaf + (int64_t)q * st.pmax * 3 * GP;
        leaf_accumulate<G, STAGE>(L0, leaf_q, sc, 0, nq0, xs, ws, lane);
        leaf_accumulate<G, STAGE>(L1, leaf_q, sc, nq0, nq, xs, ws, lane);
    }
    // The sum and root weights take the place of the leaf table, which is dead once every warp has its leaf
    // vectors: 300 + 100 broadcast loads per thread from L1/L2 in the unrolled product loops (ncu: 6.2 stall
    // cycles per issued instruction on the load scoreboard) become shared-memory reads in program order.
    float* wl = lf;                                   // WSTAGE: [Q][G*G*SP] sum weights
    float* rws = wl + Q * G * G * SP;                 // WSTAGE: [R][S*S] root weights
    const bool WSTAGE = STAGE && (Q * G * G * SP + R * S * S <= Q * st.pmax * 3 * GP);
    if (WSTAGE) {
        __syncthreads();
        if (tile == 0) {
            const float4* wsrc = reinterpret_cast<const float4*>(wlin + (int64_t)q * G * G * SP);
            float4* wdst = reinterpret_cast<float4*>(wl + q * G * G * SP);
            for (int i = lane; i < G * G * SP / 4; i += 32) cp_async16(wdst + i, wsrc + i);
            if (q < R)
                for (int i = lane; i < S * S; i += 32) cp_async4(rws + q * S * S + i, rlin + q * S * S + i);
        }
        cp_async_commit();      // wait_group only covers COMMITTED copies (found by compute-sanitizer racecheck, round 2)
        cp_async_wait<0>();
        __syncthreads();
    }
    if (tile_ok) {
        float* lv = leaf_val + (int64_t)(q * 2) * G * npad + n;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            lv[(int64_t)g * npad] = L0[g];
            lv[(int64_t)(G + g) * npad] = L1[g];
        }
        float e0[G], e1[G];
        const float m0 = shift_exp<G>(L0, e0), m1 = shift_exp<G>(L1, e1);
        float T[S];
#pragma unroll
        for (int s = 0; s < S; ++s) T[s] = 0.f;
        const float4* wq = reinterpret_cast<const float4*>(wlin + (int64_t)q * G * G * SP);
#pragma unroll
        for (int j = 0; j < G; ++j) {
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const float pk = e0[i] * e1[j];
                const int k = j * G + i;
#pragma unroll
                for (int v = 0; v < SP / 4; ++v) {
                    const float4 w = WSTAGE ? lds_f4(wl + (q * G * G + k) * SP + 4 * v) : __ldg(wq + k * (SP / 4) + v);
                    if (4 * v + 0 < S) T[4 * v + 0] = fmaf(pk, w.x, T[4 * v + 0]);
                    if (4 * v + 1 < S) T[4 * v + 1] = fmaf(pk, w.y, T[4 * v + 1]);
                    if (4 * v + 2 < S) T[4 * v + 2] = fmaf(pk, w.z, T[4 * v + 2]);
                    if (4 * v + 3 < S) T[4 * v + 3] = fmaf(pk, w.w, T[4 * v + 3]);
                }
            }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            float val;
            if (T[s] > LIN_SUM_FLOOR) {
                val = m0 + m1 + logf(T[s]);
            } else {
                val = slow_logsumexp(lv, lv + (int64_t)G * npad, (int)npad, G,
                                     wlog + (int64_t)q * G * G * SP + s, SP);
            }
            ss[(q * S + s) * 32 + lane] = val;
            sum_val[(int64_t)(q * S + s) * npad + n] = val;
        }
    }
    __syncthreads();
    if (tile_ok && q < R) {
        const int r = q;
        float A[S], B[S], eA[S], eB[S];
#pragma unroll
        for (int i = 0; i < S; ++i) {
            A[i] = ss[((2 * r) * S + i) * 32 + lane];
            B[i] = ss[((2 * r + 1) * S + i) * 32 + lane];
        }
        const float mA = shift_exp<S>(A, eA), mB = shift_exp<S>(B, eB);
        const float* rw = rlin + r * S * S;
        float U = 0.f;
#pragma unroll
        for (int j = 0; j < S; ++j) {
            float inner = 0.f;
#pragma unroll
            for (int i = 0; i < S; ++i)
                inner = fmaf(eA[i], WSTAGE ? lds_f1(rws + r * S * S + j * S + i) : __ldg(rw + j * S + i), inner);
            U = fmaf(eB[j], inner, U);
        }
        float val;
        if (U > LIN_SUM_FLOOR) {
            val = mA + mB + logf(U);
        } else {
            val = slow_logsumexp(ss + (2 * r) * S * 32 + lane, ss + (2 * r + 1) * S * 32 + lane, 32, S,
                                 rlog + r * S * S, 1);
        }
        vr[r * 32 + lane] = val;
    }
    __syncthreads();
    if (tile_ok && q == 0 && n < N) {
        float M = vr[lane];
        for (int r = 1; r < R; ++r) M = fmaxf(M, vr[r * 32 + lane]);
        float acc = 0.f;
        for (int r = 0; r < R; ++r) acc += expf(vr[r * 32 + lane] - M);
        out[n] = (M > -INFINITY) ? M + logf(acc) : M;
    }
}

// ------------------------------------------------------------------------------------
// backward 1/4: node gradients (root -> mid sums -> leaf vectors)
//   writes gleaf [2R*2*G][npad], aux_reg [2R*(2G+S)][npad] = (e0, e1, qv), aux_root
//   [R*(1+2S)][npad] = (c, eA, eB) for the parameter-gradient kernels.
// ------------------------------------------------------------------------------------
template <int G, int S, bool STAGE>
__global__ void __launch_bounds__(512) spn2_bwd_nodes_kernel(
    Spn2Dev st, int64_t N, int64_t npad, const float* __restrict__ wlin, const float* __restrict__ wlog,
    const float* __restrict__ rlin, const float* __restrict__ rlog, const float* __restrict__ leaf_val,
    const float* __restrict__ sum_val, const float* __restrict__ out, const float* __restrict__ g_out,
    float* __restrict__ gleaf, float* __restrict__ aux_reg, float* __restrict__ aux_root,
    float* __restrict__ g_wlog, float* __restrict__ g_rlog) {
    constexpr int SP = GP_<S>::v;
    extern __shared__ float smem[];
    const int R = st.R, Q = 2 * st.R;
    float* ss = smem;                 // [Q*S][32] sum values
    float* gsm = ss + Q * S * 32;     // [Q*S][32] their gradients
    float* wl = gsm + Q * S * 32;     // STAGE: [Q][G*G*SP] sum weights
    float* rws = wl + Q * G * G * SP;  // STAGE: [R][S*S] root weights
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * 32;
    const int64_t n = base + lane;
    if (STAGE && warp < Q) {
        const float4* wsrc = reinterpret_cast<const float4*>(wlin + (int64_t)warp * G * G * SP);
        float4* wdst = reinterpret_cast<float4*>(wl + warp * G * G * SP);
        for (int i = lane; i < G * G * SP / 4; i += 32) cp_async16(wdst + i, wsrc + i);
        if (warp < R)
            for (int i = lane; i < S * S; i += 32) cp_async4(rws + warp * S * S + i, rlin + warp * S * S + i);
        cp_async_commit();
    }

    for (int idx = tid; idx < Q * S * 32; idx += blockDim.x) {
        const int row = idx >> 5, pt = idx & 31;
        ss[idx] = sum_val[(int64_t)row * npad + base + pt];
    }
    if (STAGE) cp_async_wait<0>();
    __syncthreads();
    if (warp < R) {
        const int r = warp;
        const float go = (n < N) ? g_out[n] : 0.f;
        const float ov = (n < N) ? out[n] : 0.f;
        float A[S], B[S], eA[S], eB[S];
#pragma unroll
        for (int i = 0; i < S; ++i) {
            A[i] = ss[((2 * r) * S + i) * 32 + lane];
            B[i] = ss[((2 * r + 1) * S + i) * 32 + lane];
        }
        const float mA = shift_exp<S>(A, eA), mB = shift_exp<S>(B, eB);
        const float* rw = rlin + r * S * S;
        float colA[S], rowB[S];   // colA[i] = sum_j eB[j] w[j,i];  rowB[j] = sum_i eA[i] w[j,i]
#pragma unroll
        for (int i = 0; i < S; ++i) colA[i] = 0.f;
        float U = 0.f;
#pragma unroll
        for (int j = 0; j < S; ++j) {
            float inner = 0.f;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                float w;
                if constexpr (STAGE) w = lds_f1(rws + r * S * S + j * S + i); else w = __ldg(rw + j * S + i);
                inner = fmaf(eA[i], w, inner);
                colA[i] = fmaf(eB[j], w, colA[i]);
            }
            rowB[j] = inner;
            U = fmaf(eB[j], inner, U);
        }
        float* ar = aux_root + (int64_t)r * (1 + 2 * S) * npad + n;
        if (U > LIN_SUM_FLOOR) {
            const float val = mA + mB + logf(U);
            const float c = go * expf(val - ov) / U;
            ar[0] = c;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                gsm[((2 * r) * S + i) * 32 + lane] = c * eA[i] * colA[i];
                gsm[((2 * r + 1) * S + i) * 32 + lane] = c * eB[i] * rowB[i];
                ar[(int64_t)(1 + i) * npad] = eA[i];
                ar[(int64_t)(1 + S + i) * npad] = eB[i];
            }
        } else {
            const float* a0 = ss + (2 * r) * S * 32 + lane;
            const float* b0 = ss + (2 * r + 1) * S * 32 + lane;
            const float val = slow_logsumexp(a0, b0, 32, S, rlog + r * S * S, 1);
            const float gr = go * expf(val - ov);
            ar[0] = 0.f;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                gsm[((2 * r) * S + i) * 32 + lane] = 0.f;
                gsm[((2 * r + 1) * S + i) * 32 + lane] = 0.f;
                ar[(int64_t)(1 + i) * npad] = 0.f;
                ar[(int64_t)(1 + S + i) * npad] = 0.f;
            }
            if (gr != 0.f && val > -INFINITY)
                slow_sum_backward(a0, b0, 32, S, rlog + r * S * S, 1, val, gr,
                                  gsm + (2 * r) * S * 32 + lane, gsm + (2 * r + 1) * S * 32 + lane, 32,
                                  g_rlog + r * S * S);
        }
    }
    __syncthreads();
    if (warp < Q) {
        const int q = warp;
        float L0[G], L1[G], e0[G], e1[G];
        const float* lv = leaf_val + (int64_t)(q * 2) * G * npad + n;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            L0[g] = lv[(int64_t)g * npad];
            L1[g] = lv[(int64_t)(G + g) * npad];
        }
        shift_exp<G>(L0, e0);
        shift_exp<G>(L1, e1);
        float T[S], qv[S];
#pragma unroll
        for (int s = 0; s < S; ++s) T[s] = 0.f;
        const float4* wq = STAGE ? reinterpret_cast<const float4*>(wl + q * G * G * SP)
                                 : reinterpret_cast<const float4*>(wlin + (int64_t)q * G * G * SP);
#pragma unroll
        for (int j = 0; j < G; ++j) {
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const float pk = e0[i] * e1[j];
                const int k = j * G + i;
#pragma unroll
                for (int v = 0; v < SP / 4; ++v) {
                    float4 w;
                    if constexpr (STAGE) w = lds_f4(wl + (q * G * G + k) * SP + 4 * v); else w = __ldg(wq + k * (SP / 4) + v);
                    if (4 * v + 0 < S) T[4 * v + 0] = fmaf(pk, w.x, T[4 * v + 0]);
                    if (4 * v + 1 < S) T[4 * v + 1] = fmaf(pk, w.y, T[4 * v + 1]);
                    if (4 * v + 2 < S) T[4 * v + 2] = fmaf(pk, w.z, T[4 * v + 2]);
                    if (4 * v + 3 < S) T[4 * v + 3] = fmaf(pk, w.w, T[4 * v + 3]);
                }
            }
        }
        unsigned slow_mask = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float gs = gsm[(q * S + s) * 32 + lane];
            if (T[s] > LIN_SUM_FLOOR) {
                qv[s] = gs / T[s];
            } else {
                qv[s] = 0.f;
                if (gs != 0.f) slow_mask |= 1u << s;
            }
        }
        float acc0[G], acc1[G];
#pragma unroll
        for (int g = 0; g < G; ++g) { acc0[g] = 0.f; acc1[g] = 0.f; }
#pragma unroll
        for (int j = 0; j < G; ++j) {
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const int k = j * G + i;
                float v = 0.f;
#pragma unroll
                for (int u = 0; u < SP / 4; ++u) {
                    float4 w;
                    if constexpr (STAGE) w = lds_f4(wl + (q * G * G + k) * SP + 4 * u); else w = __ldg(wq + k * (SP / 4) + u);
                    if (4 * u + 0 < S) v = fmaf(qv[4 * u + 0], w.x, v);
                    if (4 * u + 1 < S) v = fmaf(qv[4 * u + 1], w.y, v);
                    if (4 * u + 2 < S) v = fmaf(qv[4 * u + 2], w.z, v);
                    if (4 * u + 3 < S) v = fmaf(qv[4 * u + 3], w.w, v);
                }
                acc0[i] = fmaf(e1[j], v, acc0[i]);
                acc1[j] = fmaf(e0[i], v, acc1[j]);
            }
        }
        float* gl = gleaf + (int64_t)(q * 2) * G * npad + n;
        float* aq = aux_reg + (int64_t)q * (2 * G + S) * npad + n;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            gl[(int64_t)g * npad] = e0[g] * acc0[g];
            gl[(int64_t)(G + g) * npad] = e1[g] * acc1[g];
            aq[(int64_t)g * npad] = e0[g];
            aq[(int64_t)(G + g) * npad] = e1[g];
        }
#pragma unroll
        for (int s = 0; s < S; ++s) aq[(int64_t)(2 * G + s) * npad] = qv[s];
        if (slow_mask) {
            for (int s = 0; s < S; ++s)
                if (slow_mask & (1u << s)) {
                    const float sumv = ss[(q * S + s) * 32 + lane];
                    if (sumv > -INFINITY)
                        slow_sum_backward(lv, lv + (int64_t)G * npad, (int)npad, G,
                                          wlog + (int64_t)q * G * G * SP + s, SP, sumv,
                                          gsm[(q * S + s) * 32 + lane], gl, gl + (int64_t)G * npad,
                                          (int)npad, g_wlog + (int64_t)q * G * G * SP + s);
                }
        }
    }
}

// ------------------------------------------------------------------------------------
// backward 2/4: input gradients.  lane = patch, each warp walks a strided set of pixels;
// a pixel belongs to exactly one leaf per repetition (pix_slot), so no atomics.
//   L = -w (a d^2 + b):  dL/dx = -2 w a d,  dL/dw = -(a d^2 + b),  w = 1 - clamp(m)
// ------------------------------------------------------------------------------------
// A CTA owns `tpc` tiles of 32 patches (16 warps each) that share one staged leaf table.
template <int G, bool HAS_MARG, bool STAGE>
__global__ void __launch_bounds__(1024) spn2_bwd_input_kernel(
    Spn2Dev st, int64_t N, int64_t npad, const float* __restrict__ x, const float* __restrict__ marg,
    const float* __restrict__ leaf, const float* __restrict__ gleaf, float* __restrict__ g_x,
    float* __restrict__ g_marg) {
    constexpr int GP = GP_<G>::v;
    constexpr int TW = 16;                          // warps per tile
    extern __shared__ __align__(16) float smem[];
    const int D = st.D, R = st.R, Q = 2 * st.R;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tpc = (blockDim.x >> 5) / TW, tile = warp / TW, tw = warp - tile * TW;
    const int per_tile = 2 * up4(D * 33) + Q * 2 * G * 32;
    float* xs = smem + tile * per_tile;
    float* ws = xs + up4(D * 33);     // weight 1 - clamp(mask); sign bit set = mask outside [0, 1]
    float* gl = ws + up4(D * 33);     // [Q*2*G][32]
    float* lf = smem + tpc * per_tile;   // STAGE: the whole leaf table [Q*pmax][3*GP]
    int32_t* slots = reinterpret_cast<int32_t*>(lf + Q * st.pmax * 3 * GP);   // STAGE: [D*R] pixel -> slot
    __shared__ int n0s[16];
    const int64_t base = ((int64_t)blockIdx.x * tpc + tile) * 32;
    const bool tile_ok = base < npad;
    if (tid < Q) n0s[tid] = __ldg(st.n0 + tid);
    if (STAGE) {
        const float4* src = reinterpret_cast<const float4*>(leaf);
        float4* dst = reinterpret_cast<float4*>(lf);
        for (int i = tid; i < Q * st.pmax * 3 * GP / 4; i += blockDim.x) cp_async16(dst + i, src + i);
        cp_async_commit();
        for (int i = tid; i < D * R; i += blockDim.x) slots[i] = __ldg(st.slot + i);
    }
    const int ttid = tid - tile * TW * 32, tthreads = TW * 32;
    if (tile_ok) {
#pragma unroll 4
        for (int idx = ttid; idx < 32 * D; idx += tthreads) {
            const int pt = idx / D, px = idx - pt * D;
            const int64_t n = base + pt;
            float xv = 0.f, mv = 0.f;
            if (n < N) {
                xv = __ldg(x + n * D + px);
                if (HAS_MARG) mv = __ldg(marg + n * D + px);
            }
            const float wv = 1.f - fminf(fmaxf(mv, 0.f), 1.f);
            xs[px * 33 + pt] = xv;
            ws[px * 33 + pt] = (mv >= 0.f && mv <= 1.f) ? wv : -wv;
        }
        for (int idx = ttid; idx < Q * 2 * G * 32; idx += tthreads) {
            const int row = idx >> 5, pt = idx & 31;
            gl[idx] = gleaf[(int64_t)row * npad + base + pt];
        }
    }
    if (STAGE) cp_async_wait<0>();
    __syncthreads();
    if (tile_ok) {
        for (int px = tw; px < D; px += TW) {
            const float xv = xs[px * 33 + lane], wraw = ws[px * 33 + lane];
            const float wv = fabsf(wraw);
            const bool inside = !signbit(wraw);
            float t1 = 0.f, t2 = 0.f;
            for (int r = 0; r < R; ++r) {
                const int slot = STAGE ? slots[px * R + r] : __ldg(st.slot + px * R + r);
                const int q = slot / st.pmax, p = slot - q * st.pmax;
                const int h = (p >= n0s[q]) ? 1 : 0;
                float mu[GP], a[GP], b[GP];
                if (STAGE) load_leaf_params<G, true>(lf + slot * 3 * GP, mu, a, b);
                else load_leaf_params<G, false>(leaf + (int64_t)slot * 3 * GP, mu, a, b);
                const float* glb = gl + ((q * 2 + h) * G) * 32 + lane;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float d = xv - mu[g];
                    const float ad = a[g] * d;
                    const float gv = glb[g * 32];
                    t1 = fmaf(gv, ad, t1);
                    t2 = fmaf(gv, fmaf(ad, d, b[g]), t2);
                }
            }
            // reuse the tile in place: every (px, lane) entry is owned by exactly one thread
            xs[px * 33 + lane] = -2.f * wv * t1;
            ws[px * 33 + lane] = inside ? t2 : 0.f;
        }
    }
    __syncthreads();
    if (tile_ok) {
        for (int idx = ttid; idx < 32 * D; idx += tthreads) {
            const int pt = idx / D, px = idx - pt * D;
            const int64_t n = base + pt;
            if (n < N) {
                if (g_x) g_x[n * D + px] = xs[px * 33 + pt];
                if (HAS_MARG && g_marg) g_marg[n * D + px] = ws[px * 33 + pt];
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// backward 3/4: leaf parameter gradients.  One CTA per (region, patch chunk); thread =
// (pixel of the region, gaussian); reduction over patches runs in registers.
//   dmu = 2a * sum gl w d ;  da = -sum gl w d^2 ;  db = -sum gl w
// ------------------------------------------------------------------------------------
template <int G, bool HAS_MARG>
__global__ void __launch_bounds__(1024) spn2_bwd_leafparam_kernel(
    Spn2Dev st, int64_t N, int64_t npad, int chunk, const float* __restrict__ x,
    const float* __restrict__ marg, const float* __restrict__ leaf, const float* __restrict__ gleaf,
    float* __restrict__ g_leaf) {
    constexpr int GP = GP_<G>::v;
    extern __shared__ float smem[];
    const int D = st.D;
    float* xs = smem;                     // [32][D+1]
    float* ws = xs + 32 * (D + 1);
    float* gls = ws + 32 * (D + 1);       // [2G][33]
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    const int nq0 = st.n0[q], nq = st.nt[q];
    const bool active = tid < nq * G;
    const int p = active ? tid / G : 0, g = active ? tid - (tid / G) * G : 0;
    const int px = st.scope[q * st.pmax + p];
    const int hg = ((p >= nq0) ? G : 0) + g;
    const float* lp = leaf + ((int64_t)q * st.pmax + p) * 3 * GP;
    const float mu = lp[g], a = lp[GP + g];
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const int64_t c0 = (int64_t)blockIdx.y * chunk;
    const int64_t c1 = min(c0 + (int64_t)chunk, N);
    for (int64_t base = c0; base < c1; base += 32) {
        for (int idx = tid; idx < 32 * D; idx += blockDim.x) {
            const int pt = idx / D, pp = idx - pt * D;
            const int64_t n = base + pt;
            float xv = 0.f, wv = 0.f;
            if (n < c1) {
                xv = __ldg(x + n * D + pp);
                wv = HAS_MARG ? 1.f - fminf(fmaxf(__ldg(marg + n * D + pp), 0.f), 1.f) : 1.f;
            }
            xs[pt * (D + 1) + pp] = xv;
            ws[pt * (D + 1) + pp] = wv;
        }
        for (int idx = tid; idx < 2 * G * 32; idx += blockDim.x) {
            const int row = idx >> 5, pt = idx & 31;
            const int64_t n = base + pt;
            gls[row * 33 + pt] = (n < c1) ? gleaf[(int64_t)((q * 2) * G + row) * npad + n] : 0.f;
        }
        __syncthreads();
        if (active) {
#pragma unroll 8
            for (int pt = 0; pt < 32; ++pt) {
                const float d = xs[pt * (D + 1) + px] - mu;
                const float gw = gls[hg * 33 + pt] * ws[pt * (D + 1) + px];
                s1 = fmaf(gw, d, s1);
                s2 = fmaf(gw * d, d, s2);
                s3 += gw;
            }
        }
        __syncthreads();
    }
    if (active) {
        float* dst = g_leaf + ((int64_t)q * st.pmax + p) * 3 * GP;
        atomicAdd(dst + g, 2.f * a * s1);
        atomicAdd(dst + GP + g, -s2);
        atomicAdd(dst + 2 * GP + g, -s3);
    }
}

// Same mapping with the patch / mask tiles and the leaf-vector gradients streamed through a
// two-stage cp.async pipeline (needs D % 4 == 0).
template <int G, bool HAS_MARG>
__global__ void __launch_bounds__(1024) spn2_bwd_leafparam_async_kernel(
    Spn2Dev st, int64_t N, int64_t npad, int chunk, const float* __restrict__ x,
    const float* __restrict__ marg, const float* __restrict__ leaf, const float* __restrict__ gleaf,
    float* __restrict__ g_leaf) {
    constexpr int GP = GP_<G>::v, GT_LD = 36;
    extern __shared__ __align__(16) float smem[];
    const int D = st.D, TILE = 32 * D;
    float* xt = smem;                     // [2][32][D]
    float* mt = xt + 2 * TILE;            // [2][32][D]
    // rows of the leaf-vector gradients are GT_LD = 36 floats apart: threads of a warp read 10-20 different rows
    // at the same patch index, which a stride of 32 puts in ONE bank (10- to 20-way conflict on a third of the
    // shared-memory reads); 36 keeps the 16-byte alignment cp.async needs and leaves at most a 3-way conflict
    float* gt = mt + 2 * TILE;            // [2][2G][GT_LD]
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    const int nq0 = st.n0[q], nq = st.nt[q];
    const bool active = tid < nq * G;
    const int p = active ? tid / G : 0, g = active ? tid - (tid / G) * G : 0;
    const int px = st.scope[q * st.pmax + p];
    const int hg = ((p >= nq0) ? G : 0) + g;
    const float* lp = leaf + ((int64_t)q * st.pmax + p) * 3 * GP;
    const float mu = lp[g], a = lp[GP + g];
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const int64_t c0 = (int64_t)blockIdx.y * chunk;
    const int64_t c1 = min(c0 + (int64_t)chunk, N);
    const int nb = (int)((c1 - c0 + 31) / 32);
    auto issue = [&](int b) {
        const int stg = b & 1;
        const int64_t base = c0 + (int64_t)b * 32;
        const int rows = (int)min((int64_t)32, c1 - base);
        // rows of a tile are contiguous in x / marg: one linear copy of rows * D floats
        const float4* xsrc = reinterpret_cast<const float4*>(x + base * D);
        const float4* msrc = reinterpret_cast<const float4*>(marg + base * D);
        float4* xdst = reinterpret_cast<float4*>(xt + stg * TILE);
        float4* mdst = reinterpret_cast<float4*>(mt + stg * TILE);
        for (int i = tid; i < rows * D / 4; i += blockDim.x) {
            cp_async16(xdst + i, xsrc + i);
            if (HAS_MARG) cp_async16(mdst + i, msrc + i);
        }
        for (int i = tid; i < 2 * G * 8; i += blockDim.x) {
            const int row = i >> 3, c4 = i & 7;
            cp_async16(gt + stg * 2 * G * GT_LD + row * GT_LD + c4 * 4,
                       gleaf + (int64_t)((q * 2) * G + row) * npad + base + c4 * 4);
        }
        cp_async_commit();
    };
    if (nb > 0) issue(0);
    for (int b = 0; b < nb; ++b) {
        if (b + 1 < nb) {
            issue(b + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (active) {
            const int stg = b & 1;
            const int lim = (int)min((int64_t)32, c1 - (c0 + (int64_t)b * 32));
            const float* xs = xt + stg * TILE + px;
            const float* ms = mt + stg * TILE + px;
            const float* gs = gt + stg * 2 * G * GT_LD + hg * GT_LD;
#pragma unroll 8
            for (int pt = 0; pt < lim; ++pt) {
                const float d = xs[pt * D] - mu;
                const float wv = HAS_MARG ? 1.f - fminf(fmaxf(ms[pt * D], 0.f), 1.f) : 1.f;
                const float gw = gs[pt] * wv;
                s1 = fmaf(gw, d, s1);
                s2 = fmaf(gw * d, d, s2);
                s3 += gw;
            }
        }
        __syncthreads();
    }
    if (active) {
        float* dst = g_leaf + ((int64_t)q * st.pmax + p) * 3 * GP;
        atomicAdd(dst + g, 2.f * a * s1);
        atomicAdd(dst + GP + g, -s2);
        atomicAdd(dst + 2 * GP + g, -s3);
    }
}

// ------------------------------------------------------------------------------------
// backward 4/4: sum-weight gradients (w.r.t. the LOG weights).
//   region q:  G_wlog[k][s] += W[k][s] * sum_n qv[n][s] e0[n][i] e1[n][j]
//   root r  :  G_rlog[k]    += W[k]    * sum_n c[n] eA[n][i] eB[n][j]
// ------------------------------------------------------------------------------------
template <int G, int S>
__global__ void __launch_bounds__(256) spn2_bwd_sumparam_kernel(
    Spn2Dev st, int64_t N, int64_t npad, int chunk, const float* __restrict__ wlin,
    const float* __restrict__ rlin, const float* __restrict__ aux_reg, const float* __restrict__ aux_root,
    float* __restrict__ g_wlog, float* __restrict__ g_rlog) {
    constexpr int SP = GP_<S>::v;
    constexpr int ROWS = (2 * G + S) > (1 + 2 * S) ? (2 * G + S) : (1 + 2 * S);
    __shared__ float tile[ROWS * 33];
    const int Q = 2 * st.R;
    const int tid = threadIdx.x;
    const bool is_region = (int)blockIdx.x < Q;
    const int64_t c0 = (int64_t)blockIdx.y * chunk;
    const int64_t c1 = min(c0 + (int64_t)chunk, N);
    if (is_region) {
        const int q = blockIdx.x;
        const bool active = tid < G * G;
        const int i = active ? tid % G : 0, j = active ? tid / G : 0;
        float acc[S];
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = 0.f;
        for (int64_t base = c0; base < c1; base += 32) {
            for (int idx = tid; idx < (2 * G + S) * 32; idx += blockDim.x) {
                const int row = idx >> 5, pt = idx & 31;
                const int64_t n = base + pt;
                tile[row * 33 + pt] = (n < c1) ? aux_reg[(int64_t)(q * (2 * G + S) + row) * npad + n] : 0.f;
            }
            __syncthreads();
            if (active) {
                for (int pt = 0; pt < 32; ++pt) {
                    const float pk = tile[i * 33 + pt] * tile[(G + j) * 33 + pt];
#pragma unroll
                    for (int s = 0; s < S; ++s) acc[s] = fmaf(pk, tile[(2 * G + s) * 33 + pt], acc[s]);
                }
            }
            __syncthreads();
        }
        if (active) {
            const int64_t o = ((int64_t)q * G * G + tid) * SP;
#pragma unroll
            for (int s = 0; s < S; ++s) atomicAdd(g_wlog + o + s, acc[s] * wlin[o + s]);
        }
    } else {
        const int r = blockIdx.x - Q;
        const bool active = tid < S * S;
        const int i = active ? tid % S : 0, j = active ? tid / S : 0;
        float acc = 0.f;
        for (int64_t base = c0; base < c1; base += 32) {
            for (int idx = tid; idx < (1 + 2 * S) * 32; idx += blockDim.x) {
                const int row = idx >> 5, pt = idx & 31;
                const int64_t n = base + pt;
                tile[row * 33 + pt] = (n < c1) ? aux_root[(int64_t)(r * (1 + 2 * S) + row) * npad + n] : 0.f;
            }
            __syncthreads();
            if (active) {
                for (int pt = 0; pt < 32; ++pt)
                    acc = fmaf(tile[pt] * tile[(1 + i) * 33 + pt], tile[(1 + S + j) * 33 + pt], acc);
            }
            __syncthreads();
        }
        if (active) atomicAdd(g_rlog + r * S * S + tid, acc * rlin[r * S * S + tid]);
    }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
static int check_struct(const stove_spn2_struct* st) {
    STOVE_CHECK_ARG(st, "null struct");
    STOVE_CHECK_ARG(st->D > 0 && st->R > 0 && st->R <= 8 && st->pmax > 0, "bad structure sizes");
    STOVE_CHECK_ARG(st->region_scope && st->region_n0 && st->region_n && st->pix_slot, "null structure table");
    return STOVE_OK;
}

static Spn2Dev to_dev(const stove_spn2_struct* st) {
    Spn2Dev d;
    d.D = st->D; d.R = st->R; d.pmax = st->pmax;
    d.scope = st->region_scope; d.n0 = st->region_n0; d.nt = st->region_n; d.slot = st->pix_slot;
    return d;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 227 * 1024) {
        stove_set_error("kernel needs %zu B of shared memory (> 227 KB)", bytes);
        return STOVE_ERR_UNSUPPORTED;
    }
    if (bytes > 48 * 1024)
        STOVE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return STOVE_OK;
}

template <int G, int S>
static int spn2_fwd_launch(const stove_spn2_struct* st, int64_t N, const float* x, const float* marg,
                           const float* leaf, const float* wlin, const float* wlog, const float* rlin,
                           const float* rlog, float* leaf_val, float* sum_val, float* out, cudaStream_t s) {
    const int Q = 2 * st->R;
    const int64_t npad = round_up64(N, 32);
    constexpr int GP = GP_<G>::v;
    const int ntiles = (int)(npad / 32);
    const size_t per_tile = sizeof(float) * ((size_t)2 * up4(st->D * 33) + (size_t)Q * S * 32 + (size_t)st->R * 32);
    const size_t stage_extra = sizeof(float) * ((size_t)Q * st->pmax * 3 * GP + (size_t)Q * st->pmax);
    // two tiles per CTA once there are more tiles than SMs (one wave instead of 1 + a tail)
    int tpc = (ntiles > 148 && 2 * Q * 32 <= 1024) ? 2 : 1;
    if (tpc * per_tile + stage_extra > 227 * 1024) tpc = 1;
    const bool stage = tpc * per_tile + stage_extra <= 227 * 1024;   // else: parameters straight from L2
    const size_t smem = tpc * per_tile + (stage ? stage_extra : 0);
    const int blocks = (ntiles + tpc - 1) / tpc, threads = 32 * Q * tpc;
    Spn2Dev d = to_dev(st);
    int rc;
#define SPN2_FWD_LAUNCH(M_, ST_)                                                                               \
    do {                                                                                                       \
        if ((rc = set_smem(spn2_fwd_kernel<G, S, M_, ST_>, smem))) return rc;                                  \
        STOVE_KERNEL(K_SPN2_FWD, s, spn2_fwd_kernel<G, S, M_, ST_><<<blocks, threads, smem, s>>>(              \
            d, N, npad, x, marg, leaf, wlin, wlog, rlin, rlog, leaf_val, sum_val, out));                       \
    } while (0)
    if (marg) { if (stage) SPN2_FWD_LAUNCH(true, true); else SPN2_FWD_LAUNCH(true, false); }
    else { if (stage) SPN2_FWD_LAUNCH(false, true); else SPN2_FWD_LAUNCH(false, false); }
#undef SPN2_FWD_LAUNCH
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

struct Spn2Ws {
    float* gleaf;
    float* aux_reg;
    float* aux_root;
    size_t bytes;
};

static Spn2Ws spn2_ws_layout(const stove_spn2_struct* st, int64_t N, void* base) {
    const int64_t npad = round_up64(N, 32);
    const int Q = 2 * st->R, G = st->G, S = st->S;
    Spn2Ws w;
    float* p = (float*)base;
    w.gleaf = p; p += (int64_t)Q * 2 * G * npad;
    w.aux_reg = p; p += (int64_t)Q * (2 * G + S) * npad;
    w.aux_root = p; p += (int64_t)st->R * (1 + 2 * S) * npad;
    w.bytes = (size_t)((char*)p - (char*)base);
    return w;
}

// the two parameter-gradient kernels of the backward pass: they read the workspace written by the node pass
// (spn2_bwd_nodes_kernel here, or the fused chain kernel of scene_ll_bwd.cu)
template <int G, int S>
static int spn2_param_launch(const stove_spn2_struct* st, int64_t N, const float* x, const float* marg,
                             const float* leaf, const float* wlin, const float* rlin, void* workspace,
                             float* g_leaf, float* g_wlog, float* g_rlog, cudaStream_t s_leaf, cudaStream_t s_sum) {
    const int Q = 2 * st->R, D = st->D;
    const int64_t npad = round_up64(N, 32);
    Spn2Dev d = to_dev(st);
    Spn2Ws w = spn2_ws_layout(st, N, workspace);
    int rc;
    // chunk the patch axis so that the grid has a few hundred CTAs
    int chunk = (int)round_up64((N + 23) / 24, 32);
    if (chunk < 32) chunk = 32;
    const int nchunk = (int)((N + chunk - 1) / chunk);
    {
        const int threads = round_up(st->pmax * G, 32);
        STOVE_CHECK_ARG(threads <= 1024, "region too large for the leaf-gradient kernel");
        dim3 grid(Q, nchunk);
        const size_t smem_async = sizeof(float) * ((size_t)4 * 32 * D + (size_t)2 * 2 * G * 36);
        if (D % 4 == 0 && smem_async <= 227 * 1024) {
            if (marg) {
                if ((rc = set_smem(spn2_bwd_leafparam_async_kernel<G, true>, smem_async))) return rc;
                STOVE_KERNEL(K_SPN2_BWD_LEAFPARAM, s_leaf, spn2_bwd_leafparam_async_kernel<G, true><<<grid, threads, smem_async, s_leaf>>>(d, N, npad, chunk, x, marg, leaf, w.gleaf, g_leaf));
            } else {
                if ((rc = set_smem(spn2_bwd_leafparam_async_kernel<G, false>, smem_async))) return rc;
                STOVE_KERNEL(K_SPN2_BWD_LEAFPARAM, s_leaf, spn2_bwd_leafparam_async_kernel<G, false><<<grid, threads, smem_async, s_leaf>>>(d, N, npad, chunk, x, marg, leaf, w.gleaf, g_leaf));
            }
        } else {
            const size_t smem = sizeof(float) * ((size_t)2 * 32 * (D + 1) + (size_t)2 * G * 33);
            if (marg) {
                if ((rc = set_smem(spn2_bwd_leafparam_kernel<G, true>, smem))) return rc;
                STOVE_KERNEL(K_SPN2_BWD_LEAFPARAM, s_leaf, spn2_bwd_leafparam_kernel<G, true><<<grid, threads, smem, s_leaf>>>(d, N, npad, chunk, x, marg, leaf, w.gleaf, g_leaf));
            } else {
                if ((rc = set_smem(spn2_bwd_leafparam_kernel<G, false>, smem))) return rc;
                STOVE_KERNEL(K_SPN2_BWD_LEAFPARAM, s_leaf, spn2_bwd_leafparam_kernel<G, false><<<grid, threads, smem, s_leaf>>>(d, N, npad, chunk, x, marg, leaf, w.gleaf, g_leaf));
            }
        }
        STOVE_LAUNCH_CHECK();
    }
    {
        const int need = (G * G > S * S) ? G * G : S * S;
        STOVE_CHECK_ARG(need <= 256, "G*G or S*S > 256");
        dim3 grid(Q + st->R, nchunk);
        STOVE_KERNEL(K_SPN2_BWD_SUMPARAM, s_sum, spn2_bwd_sumparam_kernel<G, S><<<grid, 256, 0, s_sum>>>(d, N, npad, chunk, wlin, rlin, w.aux_reg, w.aux_root, g_wlog, g_rlog));
        STOVE_LAUNCH_CHECK();
    }
    return STOVE_OK;
}

int spn2_param_kernels(const stove_spn2_struct* st, int64_t N, const float* x, const float* marg, const float* leaf,
                       const float* wlin, const float* rlin, void* workspace, float* g_leaf, float* g_wlog,
                       float* g_rlog, cudaStream_t s_leaf, cudaStream_t s_sum) {
    if (st->G == 10 && st->S == 10)
        return spn2_param_launch<10, 10>(st, N, x, marg, leaf, wlin, rlin, workspace, g_leaf, g_wlog, g_rlog, s_leaf, s_sum);
    stove_set_error("spn2_param_kernels: (num_gauss, num_sums) = (%d, %d) is not instantiated", st->G, st->S);
    return STOVE_ERR_UNSUPPORTED;
}

template <int G, int S>
static int spn2_bwd_launch(const stove_spn2_struct* st, int64_t N, const float* x, const float* marg,
                           const float* leaf, const float* wlin, const float* wlog, const float* rlin,
                           const float* rlog, const float* leaf_val, const float* sum_val,
                           const float* out, const float* g_out, float* g_x, float* g_marg, float* g_leaf,
                           float* g_wlog, float* g_rlog, void* workspace, cudaStream_t s, cudaStream_t join_s) {
    const int Q = 2 * st->R, D = st->D;
    const int64_t npad = round_up64(N, 32);
    Spn2Dev d = to_dev(st);
    Spn2Ws w = spn2_ws_layout(st, N, workspace);
    const int blocks = (int)(npad / 32);
    int rc;
    {
        constexpr int SP = GP_<S>::v;
        const size_t base_smem = sizeof(float) * (size_t)2 * Q * S * 32;
        const size_t stage_smem = base_smem + sizeof(float) * ((size_t)Q * G * G * SP + (size_t)st->R * S * S);
        // sum and root weights staged in shared memory (cp.async, read back with pinned-order loads): every
        // (pair, sum) step otherwise waits for its own broadcast load from L1/L2 (ncu: 7.7 stall cycles per issued
        // instruction on the load scoreboard); 47 -> 37 us
        if (stage_smem <= 227 * 1024 && stove_opt(OPT_SPN2_NODES_STAGE)) {
            if ((rc = set_smem(spn2_bwd_nodes_kernel<G, S, true>, stage_smem))) return rc;
            STOVE_KERNEL(K_SPN2_BWD_NODES, s, spn2_bwd_nodes_kernel<G, S, true><<<blocks, 32 * Q, stage_smem, s>>>(
                d, N, npad, wlin, wlog, rlin, rlog, leaf_val, sum_val, out, g_out, w.gleaf, w.aux_reg, w.aux_root,
                g_wlog, g_rlog));
        } else {
            if ((rc = set_smem(spn2_bwd_nodes_kernel<G, S, false>, base_smem))) return rc;
            STOVE_KERNEL(K_SPN2_BWD_NODES, s, spn2_bwd_nodes_kernel<G, S, false><<<blocks, 32 * Q, base_smem, s>>>(
                d, N, npad, wlin, wlog, rlin, rlog, leaf_val, sum_val, out, g_out, w.gleaf, w.aux_reg, w.aux_root,
                g_wlog, g_rlog));
        }
        STOVE_LAUNCH_CHECK();
    }
    // the three remaining kernels only depend on the node pass: input gradients stay on the caller's
    // stream, the two parameter-gradient kernels run on side streams
    StoveFork* fk = !stove_opt(OPT_FORK) ? nullptr : stove_fork_get(0);
    if (fk && (rc = stove_fork(fk, s, 2))) return rc;
    cudaStream_t s_leaf = fk ? fk->side[0] : s, s_sum = fk ? fk->side[1] : s;
    if (g_x || g_marg) {
        constexpr int GP = GP_<G>::v;
        const size_t per_tile = sizeof(float) * ((size_t)2 * up4(D * 33) + (size_t)Q * 2 * G * 32);
        const size_t stage_extra = sizeof(float) * ((size_t)Q * st->pmax * 3 * GP + (size_t)D * st->R);
        int tpc = blocks > 148 ? 2 : 1;
        if (tpc * per_tile + stage_extra > 227 * 1024) tpc = 1;
        const bool stage = tpc * per_tile + stage_extra <= 227 * 1024;
        const size_t smem = tpc * per_tile + (stage ? stage_extra : 0);
        const int in_blocks = (blocks + tpc - 1) / tpc, in_threads = 512 * tpc;
#define SPN2_BWD_INPUT_LAUNCH(M_, ST_)                                                                      \
    do {                                                                                                    \
        if ((rc = set_smem(spn2_bwd_input_kernel<G, M_, ST_>, smem))) return rc;                            \
        STOVE_KERNEL(K_SPN2_BWD_INPUT, s, spn2_bwd_input_kernel<G, M_, ST_><<<in_blocks, in_threads, smem, s>>>( \
            d, N, npad, x, marg, leaf, w.gleaf, g_x, g_marg));                                              \
    } while (0)
        if (marg) { if (stage) SPN2_BWD_INPUT_LAUNCH(true, true); else SPN2_BWD_INPUT_LAUNCH(true, false); }
        else { if (stage) SPN2_BWD_INPUT_LAUNCH(false, true); else SPN2_BWD_INPUT_LAUNCH(false, false); }
#undef SPN2_BWD_INPUT_LAUNCH
        STOVE_LAUNCH_CHECK();
    }
    if ((rc = spn2_param_launch<G, S>(st, N, x, marg, leaf, wlin, rlin, workspace, g_leaf, g_wlog, g_rlog, s_leaf, s_sum))) return rc;
    // the parameter gradients are consumed by the backward of the parameter packing: when the caller names the
    // stream that runs on, the side streams join THERE and the caller's stream continues right after the
    // input-gradient kernel (what the rest of the backward chain waits for)
    if (fk && (rc = stove_join(fk, join_s ? join_s : s, 2))) return rc;
    return STOVE_OK;
}

#define SPN2_DISPATCH(CALL)                                                                   \
    if (st->G == 10 && st->S == 10) return CALL(10, 10);                                      \
    if (st->G == 4 && st->S == 4) return CALL(4, 4);                                          \
    if (st->G == 8 && st->S == 8) return CALL(8, 8);                                          \
    if (st->G == 6 && st->S == 3) return CALL(6, 3);                                          \
    stove_set_error("spn2: (num_gauss, num_sums) = (%d, %d) is not instantiated", st->G, st->S); \
    return STOVE_ERR_UNSUPPORTED;

extern "C" int stove_spn2_fwd(const stove_spn2_struct* st, int64_t N, const float* x, const float* marg,
                              const float* leaf, const float* wlin, const float* wlog, const float* rlin,
                              const float* rlog, float* leaf_val, float* sum_val, float* out, void* stream) {
    int rc = check_struct(st);
    if (rc) return rc;
    STOVE_CHECK_ARG(N >= 0 && x && leaf && wlin && wlog && rlin && rlog && leaf_val && sum_val && out, "null pointer");
    if (N == 0) return STOVE_OK;
    cudaStream_t s = (cudaStream_t)stream;
#define CALL(G_, S_) spn2_fwd_launch<G_, S_>(st, N, x, marg, leaf, wlin, wlog, rlin, rlog, leaf_val, sum_val, out, s)
    SPN2_DISPATCH(CALL)
#undef CALL
}

extern "C" size_t stove_spn2_bwd_workspace(const stove_spn2_struct* st, int64_t N) {
    if (!st || N <= 0) return 0;
    return spn2_ws_layout(st, N, nullptr).bytes;
}

extern "C" int stove_spn2_bwd(const stove_spn2_struct* st, int64_t N, const float* x, const float* marg,
                              const float* leaf, const float* wlin, const float* wlog, const float* rlin,
                              const float* rlog, const float* leaf_val, const float* sum_val,
                              const float* out, const float* g_out, float* g_x, float* g_marg,
                              float* g_leaf, float* g_wlog, float* g_rlog, void* workspace, void* stream,
                              void* join_stream) {
    int rc = check_struct(st);
    if (rc) return rc;
    STOVE_CHECK_ARG(N >= 0 && x && leaf && wlin && wlog && rlin && rlog && leaf_val && sum_val && out && g_out &&
                        g_leaf && g_wlog && g_rlog && workspace, "null pointer");
    STOVE_CHECK_ARG(!(g_marg && !marg), "g_marg requested without marg");
    if (N == 0) return STOVE_OK;
    cudaStream_t s = (cudaStream_t)stream;
#define CALL(G_, S_) spn2_bwd_launch<G_, S_>(st, N, x, marg, leaf, wlin, wlog, rlin, rlog, leaf_val, sum_val, \
                                             out, g_out, g_x, g_marg, g_leaf, g_wlog, g_rlog, workspace, s, \
                                             (cudaStream_t)join_stream)
    SPN2_DISPATCH(CALL)
#undef CALL
}
