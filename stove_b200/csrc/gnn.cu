// Fused graph-network dynamics step (forward, backward) and persistent rollout, fp32, sm_100a.
//
// Replaces Dynamics.forward + core (model/video_prediction/dynamics.py:181-265) for
// core_idx 0 -- ~20 addmm + ~350 elementwise launches per step in the reference -- and the
// time loop of Stove.rollout (model/video_prediction/stove.py:777-861).
//
// One CTA owns SEQ sequences: all O^2 pair rows and O object rows of those sequences live in
// shared memory, feature-major ([feature][row]) so that a thread computing 4 output features
// of one row reads its activations conflict-free and the weights as float4.  Weights are staged
// once per CTA in shared memory (~92 KB); in the rollout kernel they stay there for all time
// steps, and the state never leaves the chip between steps.
// The backward kernel recomputes the forward on-chip (no activation tape in HBM), then walks
// the layers in reverse; weight gradients go to a per-CTA slab in global memory (plain stores,
// no atomics) that a second kernel reduces.
#include "gnn_common.cuh"





// row stride of a feature-major buffer: multiple of 4, congruent 4 mod 8 (see dense())
__host__ __device__ static inline int pad_ld(int rows) {
    int ld = (rows + 3) & ~3;
    if ((ld & 7) == 0) ld += 4;
    return ld;
}

// shared-memory activation buffers for SEQ sequences (float offsets)
struct GnnBuf {
    int ldo, ldp, lds;     // row strides: object rows, pair rows, sequence rows
    int sin, emb, s, h, selfd, comb, ra0, r1, a1, rel, att, d, f1, f2, cat, o1, out;
    int rh0, rh1, rsum, r2, r3, rew;
    // backward only
    int g_out, g_o1, g_cat, g_f2, g_f1, g_d, g_r1, g_a1, g_att, g_self, g_h, g_s, g_sin, g_emb;
    int g_rew, g_r3, g_r2, g_rsum, g_rh1, g_rh0;
    int total;
};

__host__ __device__ static inline GnnBuf gnn_buffers(const stove_gnn_cfg& c, int in_dim, int seq, bool bwd) {
    GnnBuf b;
    const int cl = c.cl, O = c.num_obj;
    b.ldo = pad_ld(seq * O);
    b.ldp = pad_ld(seq * O * O);
    b.lds = pad_ld(seq);
    int at = 0;
    auto take = [&](int& off, int feats, int ld) { off = at; at += feats * ld; };
    take(b.sin, in_dim, b.ldo);
    take(b.emb, c.action_dim > 0 ? O * 4 : 0, b.lds);
    take(b.s, cl, b.ldo);
    take(b.h, cl, b.ldo);
    take(b.selfd, cl, b.ldo);
    take(b.comb, 2 * cl + 1, b.ldp);
    take(b.ra0, 4 * cl, b.ldp);
    take(b.r1, cl, b.ldp);
    take(b.a1, cl, b.ldp);
    take(b.rel, cl, b.ldp);
    take(b.att, 1, b.ldp);
    take(b.d, cl, b.ldo);
    take(b.f1, cl, b.ldo);
    take(b.f2, cl, b.ldo);
    take(b.cat, 2 * cl, b.ldo);
    take(b.o1, cl, b.ldo);
    take(b.out, cl, b.ldo);
    const int rw = c.reward ? 1 : 0;
    take(b.rh0, rw * cl, b.ldo);
    take(b.rh1, rw * cl, b.ldo);
    take(b.rsum, rw * cl, b.lds);
    take(b.r2, rw * cl / 2, b.lds);
    take(b.r3, rw * cl / 4, b.lds);
    take(b.rew, rw, b.lds);
    if (bwd) {
        take(b.g_out, cl, b.ldo);
        take(b.g_o1, cl, b.ldo);
        take(b.g_cat, 2 * cl, b.ldo);
        take(b.g_f2, cl, b.ldo);
        take(b.g_f1, cl, b.ldo);
        take(b.g_d, cl, b.ldo);
        take(b.g_r1, cl, b.ldp);
        take(b.g_a1, cl, b.ldp);
        take(b.g_att, 1, b.ldp);
        take(b.g_self, cl, b.ldo);
        take(b.g_h, cl, b.ldo);
        take(b.g_s, cl, b.ldo);
        take(b.g_sin, in_dim, b.ldo);
        take(b.g_emb, c.action_dim > 0 ? O * 4 : 0, b.lds);
        take(b.g_rew, rw, b.lds);
        take(b.g_r3, rw * cl / 4, b.lds);
        take(b.g_r2, rw * cl / 2, b.lds);
        take(b.g_rsum, rw * cl, b.lds);
        take(b.g_rh1, rw * cl, b.ldo);
        take(b.g_rh0, rw * cl, b.ldo);
    }
    b.total = at;
    return b;
}


// ------------------------------------------------------------------------------------
// Dense layers on feature-major shared-memory activations ([feature][row], row stride ld).
// ld is a multiple of 4 with ld = 4 (mod 8) (pad_ld): groups of 4 rows are float4-aligned and
// lanes that walk the feature index hit distinct bank groups.  Register blocking: one item =
// 4 rows x 4 output features (16 FMA per 2 LDS.128); small layers fall back to 1 x 4 so that
// more threads take part (these launches are latency bound, not throughput bound).
// Rows beyond `rows` inside the last group hold garbage; rows never mix, and every reduction
// over rows below stops at `rows`.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float4 act4(float4 v, int act, int nonlin) {
    v.x = apply_act(v.x, act, nonlin);
    v.y = apply_act(v.y, act, nonlin);
    v.z = apply_act(v.z, act, nonlin);
    v.w = apply_act(v.w, act, nonlin);
    return v;
}

// out[n][row] = act(b[n] + sum_k W[k][n] in[k][row]) (+ res[n][row]);  W is [K][N]
__device__ __noinline__ void dense(const float* __restrict__ W, const float* __restrict__ bias, int K, int N,
                      const float* in, int ldi, float* out, int ldo, int rows, int act, int nonlin,
                      const float* res, int ldr) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if ((N & 3) == 0) {
        const int n4 = N >> 2, ngrp = (rows + 3) >> 2;
        const float4* W4 = reinterpret_cast<const float4*>(W);
        const float4* b4 = reinterpret_cast<const float4*>(bias);
        if (ngrp * n4 * 2 >= nt) {
            for (int it = tid; it < ngrp * n4; it += nt) {
                const int rg = it / n4, c4 = it - rg * n4;
                const float4 bv = b4[c4];
                float4 a0 = make_float4(bv.x, bv.x, bv.x, bv.x), a1 = make_float4(bv.y, bv.y, bv.y, bv.y),
                       a2 = make_float4(bv.z, bv.z, bv.z, bv.z), a3 = make_float4(bv.w, bv.w, bv.w, bv.w);
                const float* ip = in + rg * 4;
#pragma unroll 2
                for (int k = 0; k < K; ++k) {
                    const float4 x = *reinterpret_cast<const float4*>(ip + k * ldi);
                    const float4 w = W4[k * n4 + c4];
                    a0.x = fmaf(x.x, w.x, a0.x); a0.y = fmaf(x.y, w.x, a0.y); a0.z = fmaf(x.z, w.x, a0.z); a0.w = fmaf(x.w, w.x, a0.w);
                    a1.x = fmaf(x.x, w.y, a1.x); a1.y = fmaf(x.y, w.y, a1.y); a1.z = fmaf(x.z, w.y, a1.z); a1.w = fmaf(x.w, w.y, a1.w);
                    a2.x = fmaf(x.x, w.z, a2.x); a2.y = fmaf(x.y, w.z, a2.y); a2.z = fmaf(x.z, w.z, a2.z); a2.w = fmaf(x.w, w.z, a2.w);
                    a3.x = fmaf(x.x, w.w, a3.x); a3.y = fmaf(x.y, w.w, a3.y); a3.z = fmaf(x.z, w.w, a3.z); a3.w = fmaf(x.w, w.w, a3.w);
                }
                float4 y[4] = {act4(a0, act, nonlin), act4(a1, act, nonlin), act4(a2, act, nonlin), act4(a3, act, nonlin)};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (res) {
                        const float4 r = *reinterpret_cast<const float4*>(res + (c4 * 4 + e) * ldr + rg * 4);
                        y[e].x += r.x; y[e].y += r.y; y[e].z += r.z; y[e].w += r.w;
                    }
                    *reinterpret_cast<float4*>(out + (c4 * 4 + e) * ldo + rg * 4) = y[e];
                }
            }
        } else {
            for (int it = tid; it < rows * n4; it += nt) {
                const int row = it / n4, c4 = it - row * n4;
                float4 acc = b4[c4];
                const float* ip = in + row;
#pragma unroll 4
                for (int k = 0; k < K; ++k) {
                    const float a = ip[k * ldi];
                    const float4 w = W4[k * n4 + c4];
                    acc.x = fmaf(a, w.x, acc.x);
                    acc.y = fmaf(a, w.y, acc.y);
                    acc.z = fmaf(a, w.z, acc.z);
                    acc.w = fmaf(a, w.w, acc.w);
                }
                float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float y = apply_act(v[e], act, nonlin);
                    if (res) y += res[(c4 * 4 + e) * ldr + row];
                    out[(c4 * 4 + e) * ldo + row] = y;
                }
            }
        }
    } else {
        for (int it = tid; it < rows * N; it += nt) {
            const int n = it / rows, row = it - n * rows;
            float acc = bias[n];
            for (int k = 0; k < K; ++k) acc = fmaf(in[k * ldi + row], W[k * N + n], acc);
            float y = apply_act(acc, act, nonlin);
            if (res) y += res[n * ldr + row];
            out[n * ldo + row] = y;
        }
    }
}

// gin[k][row] (=, +=) sum_n W[k][n] gp[n][row], optionally scaled by act'(ymul[k][row]);
// gin may alias ymul (each element is read and written by the same thread)
__device__ __noinline__ void dense_bwd_input(const float* __restrict__ W, int K, int N, const float* gp, int ldg,
                                float* gin, int ldi, int rows, bool accumulate,
                                const float* ymul = nullptr, int ldy = 0, int act = ACT_NONE, int nonlin = 0) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if ((N & 3) == 0) {
        const int ngrp = (rows + 3) >> 2, n4 = N >> 2;
        for (int it = tid; it < K * ngrp; it += nt) {
            const int k = it / ngrp, rg = it - k * ngrp;
            const float4* w4 = reinterpret_cast<const float4*>(W + k * N);
            const float* g = gp + rg * 4;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
            for (int q = 0; q < n4; ++q) {
                const float4 w = w4[q];
                const float4 g0 = *reinterpret_cast<const float4*>(g + (4 * q + 0) * ldg);
                const float4 g1 = *reinterpret_cast<const float4*>(g + (4 * q + 1) * ldg);
                const float4 g2 = *reinterpret_cast<const float4*>(g + (4 * q + 2) * ldg);
                const float4 g3 = *reinterpret_cast<const float4*>(g + (4 * q + 3) * ldg);
                acc.x = fmaf(w.x, g0.x, acc.x); acc.y = fmaf(w.x, g0.y, acc.y); acc.z = fmaf(w.x, g0.z, acc.z); acc.w = fmaf(w.x, g0.w, acc.w);
                acc.x = fmaf(w.y, g1.x, acc.x); acc.y = fmaf(w.y, g1.y, acc.y); acc.z = fmaf(w.y, g1.z, acc.z); acc.w = fmaf(w.y, g1.w, acc.w);
                acc.x = fmaf(w.z, g2.x, acc.x); acc.y = fmaf(w.z, g2.y, acc.y); acc.z = fmaf(w.z, g2.z, acc.z); acc.w = fmaf(w.z, g2.w, acc.w);
                acc.x = fmaf(w.w, g3.x, acc.x); acc.y = fmaf(w.w, g3.y, acc.y); acc.z = fmaf(w.w, g3.z, acc.z); acc.w = fmaf(w.w, g3.w, acc.w);
            }
            float* dst = gin + k * ldi + rg * 4;
            if (ymul) {
                const float4 y = *reinterpret_cast<const float4*>(ymul + k * ldy + rg * 4);
                acc.x *= act_grad(y.x, act, nonlin); acc.y *= act_grad(y.y, act, nonlin);
                acc.z *= act_grad(y.z, act, nonlin); acc.w *= act_grad(y.w, act, nonlin);
            }
            if (accumulate) {
                const float4 o = *reinterpret_cast<const float4*>(dst);
                acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            }
            *reinterpret_cast<float4*>(dst) = acc;
        }
    } else {
        for (int it = tid; it < rows * K; it += nt) {
            const int k = it / rows, row = it - k * rows;
            float acc = 0.f;
            const float* w = W + k * N;
            for (int n = 0; n < N; ++n) acc = fmaf(w[n], gp[n * ldg + row], acc);
            if (ymul) acc *= act_grad(ymul[k * ldy + row], act, nonlin);
            if (accumulate) gin[k * ldi + row] += acc;
            else gin[k * ldi + row] = acc;
        }
    }
}

// slab_w[k][n] (=, +=) sum_row x[k][row] gp[n][row];  slab_b[n] (=, +=) sum_row gp[n][row]
__device__ __noinline__ void dense_bwd_weight(float* __restrict__ slab_w, float* __restrict__ slab_b, int K, int N,
                                 const float* x, int ldx, const float* gp, int ldg, int rows, bool accumulate) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int full = rows >> 2;
    if ((N & 3) == 0) {
        const int n4 = N >> 2;
        for (int it = tid; it < K * n4; it += nt) {
            const int c4 = it / K, k = it - c4 * K;      // k fastest: x reads hit distinct bank groups
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            const float* xp = x + k * ldx;
            const float* g = gp + (c4 * 4) * ldg;
            for (int rg = 0; rg < full; ++rg) {
                const float4 xv = *reinterpret_cast<const float4*>(xp + rg * 4);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 gv = *reinterpret_cast<const float4*>(g + e * ldg + rg * 4);
                    acc[e] = fmaf(xv.x, gv.x, acc[e]);
                    acc[e] = fmaf(xv.y, gv.y, acc[e]);
                    acc[e] = fmaf(xv.z, gv.z, acc[e]);
                    acc[e] = fmaf(xv.w, gv.w, acc[e]);
                }
            }
            for (int row = full * 4; row < rows; ++row) {
                const float xv = xp[row];
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[e] = fmaf(xv, g[e * ldg + row], acc[e]);
            }
            float4* dst = reinterpret_cast<float4*>(slab_w + k * N + c4 * 4);
            float4 v = make_float4(acc[0], acc[1], acc[2], acc[3]);
            if (accumulate) {
                const float4 o = *dst;
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *dst = v;
        }
    } else {
        for (int it = tid; it < K * N; it += nt) {
            const int n = it / K, k = it - n * K;
            float acc = 0.f;
            for (int row = 0; row < rows; ++row) acc = fmaf(x[k * ldx + row], gp[n * ldg + row], acc);
            if (accumulate) slab_w[k * N + n] += acc;
            else slab_w[k * N + n] = acc;
        }
    }
    for (int n = tid; n < N; n += nt) {
        float acc = 0.f;
        for (int row = 0; row < rows; ++row) acc += gp[n * ldg + row];
        if (accumulate) slab_b[n] += acc;
        else slab_b[n] = acc;
    }
}

// gp[n][row] *= act'(y[n][row])
__device__ void scale_by_act_grad(float* gp, int ldg, const float* y, int ldy, int feats, int rows, int act,
                                  int nonlin) {
    for (int it = threadIdx.x; it < feats * rows; it += blockDim.x) {
        const int n = it / rows, row = it - n * rows;
        gp[n * ldg + row] *= act_grad(y[n * ldy + row], act, nonlin);
    }
}

// ------------------------------------------------------------------------------------
// forward core on shared-memory buffers: expects sm[b.sin] (state part, rows = nseq*O) and,
// if action conditioned, act rows in sm[b.emb] are computed here from `actions`.
// ------------------------------------------------------------------------------------
__device__ void gnn_forward_core(const stove_gnn_cfg& c, const GnnLayout& L, const GnnBuf& b,
                                 const float* __restrict__ W, float* sm, int nseq,
                                 const float* __restrict__ act_rows /* [nseq] rows of A floats, global */,
                                 int64_t act_stride) {
    const int cl = c.cl, O = c.num_obj, nl = c.nonlin;
    const int RO = nseq * O, RP = nseq * O * O;
    const int tid = threadIdx.x, nt = blockDim.x;
    if (c.action_dim > 0) {
        // emb[o*4+e][seq] = act_W^T a + b   (dynamics.py:238-244)
        for (int it = tid; it < nseq * O * 4; it += nt) {
            const int sq = it / (O * 4), n = it - sq * (O * 4);
            float acc = W[L.act_b + n];
            const float* a = act_rows + sq * act_stride;
            for (int k = 0; k < c.action_dim; ++k) acc = fmaf(__ldg(a + k), W[L.act_w + k * O * 4 + n], acc);
            sm[b.emb + n * b.lds + sq] = acc;
        }
        __syncthreads();
        for (int it = tid; it < nseq * O * 4; it += nt) {
            const int sq = it / (O * 4), n = it - sq * (O * 4);
            const int o = n >> 2, e = n & 3;
            sm[b.sin + (gnn_sdim(c) + e) * b.ldo + sq * O + o] = sm[b.emb + n * b.lds + sq];
        }
        __syncthreads();
    }
    // state encoder with raw pass-through of the first lim_enc dims (dynamics.py:250)
    dense(W + L.enc_w, W + L.enc_b, L.in_dim, cl, sm + b.sin, b.ldo, sm + b.s, b.ldo, RO, ACT_NONE, nl, nullptr, 0);
    __syncthreads();
    for (int it = tid; it < c.lim_enc * RO; it += nt) {
        const int k = it / RO, row = it - k * RO;
        sm[b.s + k * b.ldo + row] = sm[b.sin + k * b.ldo + row];
    }
    __syncthreads();
    // pair inputs [s_i, s_j, |p_i - p_j|^2]  (dynamics.py:186-193)
    for (int it = tid; it < (2 * cl + 1) * RP; it += nt) {
        const int k = it / RP, p = it - k * RP;
        const int sq = p / (O * O), ij = p - sq * O * O, i = ij / O, j = ij - i * O;
        float v;
        if (k < cl) v = sm[b.s + k * b.ldo + sq * O + i];
        else if (k < 2 * cl) v = sm[b.s + (k - cl) * b.ldo + sq * O + j];
        else {
            const float dx = sm[b.s + sq * O + i] - sm[b.s + sq * O + j];
            const float dy = sm[b.s + b.ldo + sq * O + i] - sm[b.s + b.ldo + sq * O + j];
            v = dx * dx + dy * dy;
        }
        sm[b.comb + k * b.ldp + p] = v;
    }
    dense(W + L.self0_w, W + L.self0_b, cl, cl, sm + b.s, b.ldo, sm + b.h, b.ldo, RO, ACT_NL, nl, nullptr, 0);
    __syncthreads();
    dense(W + L.self1_w, W + L.self1_b, cl, cl, sm + b.h, b.ldo, sm + b.selfd, b.ldo, RO, ACT_NONE, nl, sm + b.h, b.ldo);
    dense(W + L.ra0_w, W + L.ra0_b, 2 * cl + 1, 4 * cl, sm + b.comb, b.ldp, sm + b.ra0, b.ldp, RP, ACT_NL, nl, nullptr, 0);
    __syncthreads();
    dense(W + L.rel1_w, W + L.rel1_b, 2 * cl, cl, sm + b.ra0, b.ldp, sm + b.r1, b.ldp, RP, ACT_NL, nl, nullptr, 0);
    dense(W + L.att1_w, W + L.att1_b, 2 * cl, cl, sm + b.ra0 + 2 * cl * b.ldp, b.ldp, sm + b.a1, b.ldp, RP, ACT_NL, nl, nullptr, 0);
    __syncthreads();
    dense(W + L.rel2_w, W + L.rel2_b, cl, cl, sm + b.r1, b.ldp, sm + b.rel, b.ldp, RP, ACT_NONE, nl, sm + b.r1, b.ldp);
    dense(W + L.att2_w, W + L.att2_b, cl, 1, sm + b.a1, b.ldp, sm + b.att, b.ldp, RP, ACT_EXP, nl, nullptr, 0);
    __syncthreads();
    // d_i = self_i + sum_{j != i} rel_ij * att_ij   (dynamics.py:203-208; the diagonal is
    // multiplied by the zero mask exactly as the reference does, so inf * 0 = nan is preserved)
    for (int it = tid; it < cl * RO; it += nt) {
        const int k = it / RO, row = it - k * RO;
        const int sq = row / O, i = row - sq * O;
        float acc = 0.f;
        for (int j = 0; j < O; ++j) {
            const int p = sq * O * O + i * O + j;
            acc += sm[b.rel + k * b.ldp + p] * (i == j ? 0.f : 1.f) * sm[b.att + p];
        }
        sm[b.d + k * b.ldo + row] = sm[b.selfd + k * b.ldo + row] + acc;
    }
    __syncthreads();
    dense(W + L.aff0_w, W + L.aff0_b, cl, cl, sm + b.d, b.ldo, sm + b.f1, b.ldo, RO, ACT_TANH, nl, nullptr, 0);
    if (c.reward)
        dense(W + L.rew00_w, W + L.rew00_b, cl, cl, sm + b.d, b.ldo, sm + b.rh0, b.ldo, RO, ACT_RELU, nl, nullptr, 0);
    __syncthreads();
    dense(W + L.aff1_w, W + L.aff1_b, cl, cl, sm + b.f1, b.ldo, sm + b.f2, b.ldo, RO, ACT_TANH, nl, sm + b.f1, b.ldo);
    if (c.reward)
        dense(W + L.rew02_w, W + L.rew02_b, cl, cl, sm + b.rh0, b.ldo, sm + b.rh1, b.ldo, RO, ACT_NONE, nl, nullptr, 0);
    __syncthreads();
    // cat = [aff3, s]
    dense(W + L.aff2_w, W + L.aff2_b, cl, cl, sm + b.f2, b.ldo, sm + b.cat, b.ldo, RO, ACT_NONE, nl, nullptr, 0);
    for (int it = tid; it < cl * RO; it += nt) {
        const int k = it / RO, row = it - k * RO;
        sm[b.cat + (cl + k) * b.ldo + row] = sm[b.s + k * b.ldo + row];
    }
    if (c.reward) {
        for (int it = tid; it < cl * nseq; it += nt) {
            const int k = it / nseq, sq = it - k * nseq;
            float acc = 0.f;
            for (int o = 0; o < O; ++o) acc += sm[b.rh1 + k * b.ldo + sq * O + o];
            sm[b.rsum + k * b.lds + sq] = acc;
        }
    }
    __syncthreads();
    dense(W + L.out0_w, W + L.out0_b, 2 * cl, cl, sm + b.cat, b.ldo, sm + b.o1, b.ldo, RO, ACT_TANH, nl, nullptr, 0);
    if (c.reward)
        dense(W + L.rew10_w, W + L.rew10_b, cl, cl / 2, sm + b.rsum, b.lds, sm + b.r2, b.lds, nseq, ACT_RELU, nl, nullptr, 0);
    __syncthreads();
    dense(W + L.out1_w, W + L.out1_b, cl, cl, sm + b.o1, b.ldo, sm + b.out, b.ldo, RO, ACT_NONE, nl, sm + b.o1, b.ldo);
    if (c.reward)
        dense(W + L.rew12_w, W + L.rew12_b, cl / 2, cl / 4, sm + b.r2, b.lds, sm + b.r3, b.lds, nseq, ACT_RELU, nl, nullptr, 0);
    __syncthreads();
    if (c.reward) {
        dense(W + L.rew14_w, W + L.rew14_b, cl / 4, 1, sm + b.r3, b.lds, sm + b.rew, b.lds, nseq, ACT_SIGMOID, nl, nullptr, 0);
        __syncthreads();
    }
}

__device__ __forceinline__ void stage_weights(const float* __restrict__ weights, float* Ws, int total) {
    const float4* src = reinterpret_cast<const float4*>(weights);
    float4* dst = reinterpret_cast<float4*>(Ws);
    for (int i = threadIdx.x; i < total / 4; i += blockDim.x) dst[i] = __ldg(src + i);
}

// load the state part of s_in: s [n][O][sdim] (+ appearance [n][O][app_dim])
__device__ __forceinline__ void load_inputs(const stove_gnn_cfg& c, const GnnBuf& b, float* sm, int64_t seq0,
                                            int nseq, const float* __restrict__ s, int sdim, int soff,
                                            const float* __restrict__ app) {
    const int O = c.num_obj, half = gnn_sdim(c);
    for (int it = threadIdx.x; it < nseq * O * half; it += blockDim.x) {
        const int row = it / half, k = it - row * half;
        sm[b.sin + k * b.ldo + row] = __ldg(s + (seq0 * O + row) * sdim + soff + k);
    }
    if (c.app_dim > 0) {
        const int a0 = half + (c.action_dim > 0 ? 4 : 0);
        for (int it = threadIdx.x; it < nseq * O * c.app_dim; it += blockDim.x) {
            const int row = it / c.app_dim, k = it - row * c.app_dim;
            sm[b.sin + (a0 + k) * b.ldo + row] = __ldg(app + (seq0 * O + row) * c.app_dim + k);
        }
    }
}

// ------------------------------------------------------------------------------------
// forward kernel (one dynamics step)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) gnn_fwd_kernel(stove_gnn_cfg c, GnnLayout L, int seq, int64_t n,
                                                      const float* __restrict__ s,
                                                      const float* __restrict__ actions,
                                                      const float* __restrict__ app,
                                                      const float* __restrict__ weights,
                                                      float* __restrict__ out, float* __restrict__ reward) {
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;
    float* sm = smem + L.total;
    const GnnBuf b = gnn_buffers(c, L.in_dim, seq, false);
    stage_weights(weights, Ws, L.total);
    for (int i = threadIdx.x; i < b.total; i += blockDim.x) sm[i] = 0.f;
    const int cl = c.cl, O = c.num_obj;
    const int64_t ngroups = (n + seq - 1) / seq;
    for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int64_t seq0 = grp * seq;
        const int nseq = (int)min((int64_t)seq, n - seq0);
        __syncthreads();
        load_inputs(c, b, sm, seq0, nseq, s, gnn_sdim(c), 0, app);
        __syncthreads();
        gnn_forward_core(c, L, b, Ws, sm, nseq, actions ? actions + seq0 * c.action_dim : nullptr, c.action_dim);
        for (int it = threadIdx.x; it < nseq * O * cl; it += blockDim.x) {
            const int row = it / cl, k = it - row * cl;
            out[(seq0 * O + row) * cl + k] = sm[b.out + k * b.ldo + row];
        }
        if (c.reward && reward)
            for (int sq = threadIdx.x; sq < nseq; sq += blockDim.x) reward[seq0 + sq] = sm[b.rew + sq];
    }
}

// ------------------------------------------------------------------------------------
// rollout kernel: `num` dynamics steps with the state resident in shared memory
//   (stove.py:823-846 + dynamics.py:147-179 constrain_z_dyn)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) gnn_rollout_kernel(
    stove_gnn_cfg c, GnnLayout L, int seq, int64_t n, int num, const float* __restrict__ z_last,
    const float* __restrict__ actions, int action_len, const float* __restrict__ app,
    const float* __restrict__ weights, const float* __restrict__ noise, float pos_var, float vel_std,
    float latent_std, float* __restrict__ z_out, float* __restrict__ std_out, float* __restrict__ logq_out,
    float* __restrict__ rewards) {
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;
    float* sm = smem + L.total;
    const GnnBuf b = gnn_buffers(c, L.in_dim, seq, false);
    stage_weights(weights, Ws, L.total);
    for (int i = threadIdx.x; i < b.total; i += blockDim.x) sm[i] = 0.f;
    const int cl = c.cl, O = c.num_obj, half = cl / 2, zd = half + 2;
    const int64_t ngroups = (n + seq - 1) / seq;
    for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int64_t seq0 = grp * seq;
        const int nseq = (int)min((int64_t)seq, n - seq0);
        const int RO = nseq * O;
        __syncthreads();
        load_inputs(c, b, sm, seq0, nseq, z_last, zd, 2, app);
        __syncthreads();
        for (int t = 0; t < num; ++t) {
            const float* arow = nullptr;
            if (actions) arow = actions + (seq0 * action_len + (t % action_len)) * c.action_dim;
            gnn_forward_core(c, L, b, Ws, sm, nseq, arow, (int64_t)action_len * c.action_dim);
            // constrain + integrate positions, write the new state back into s_in
            for (int it = threadIdx.x; it < RO * half; it += blockDim.x) {
                const int row = it / half, k = it - row * half;
                const int64_t gr = seq0 * O + row;
                float m = 2.f * sigmoidf_(sm[b.out + k * b.ldo + row]) - 1.f;
                if (k < 2) m += sm[b.sin + k * b.ldo + row];
                float val = m;
                const int64_t o16 = ((gr / O) * num + t) * O * half + (gr % O) * half + k;
                if (noise || std_out) {
                    const float sraw = sigmoidf_(sm[b.out + (half + k) * b.ldo + row]);
                    const float sd = (k < 2 ? pos_var : (k < 4 ? vel_std : latent_std)) * sraw;
                    if (std_out) std_out[o16] = sd;
                    if (noise) {
                        const float e = __ldg(noise + o16);
                        val = m + sd * e;
                        if (logq_out) logq_out[o16] = -0.5f * e * e - logf(sd) - HALF_LOG_2PI;
                    }
                }
                // stash in `out` (no longer needed) so every thread reads the OLD s_in above
                sm[b.out + k * b.ldo + row] = val;
                const int64_t oz = ((gr / O) * num + t) * O * zd + (gr % O) * zd;
                z_out[oz + 2 + k] = val;
                if (k < 2) z_out[oz + k] = __ldg(z_last + gr * zd + k);
            }
            if (c.reward && rewards)
                for (int sq = threadIdx.x; sq < nseq; sq += blockDim.x)
                    rewards[(seq0 + sq) * num + t] = sm[b.rew + sq];
            __syncthreads();
            for (int it = threadIdx.x; it < RO * half; it += blockDim.x) {
                const int row = it / half, k = it - row * half;
                sm[b.sin + k * b.ldo + row] = sm[b.out + k * b.ldo + row];
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------
// Backward core: expects the forward activations (gnn_forward_core) and the upstream gradient
// sm[b.g_out] in shared memory; leaves d/d(s_in) in sm[b.g_sin] (first cl/2 features, raw
// pass-through already added) and writes this CTA's weight gradients into `slab`.
// ------------------------------------------------------------------------------------
__device__ void gnn_backward_core(const stove_gnn_cfg& c, const GnnLayout& L, const GnnBuf& b,
                                  const float* __restrict__ W, float* sm, float* __restrict__ slab, bool accum,
                                  int nseq, const float* __restrict__ actions, int64_t act_stride,
                                  const float* __restrict__ g_reward, int64_t g_reward_stride) {
    const int cl = c.cl, O = c.num_obj, nl = c.nonlin, half = cl / 2;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int RO = nseq * O, RP = nseq * O * O;
    // ---- out1: result = W o1 + b + o1
    dense_bwd_weight(slab + L.out1_w, slab + L.out1_b, cl, cl, sm + b.o1, b.ldo, sm + b.g_out, b.ldo, RO, accum);
    dense_bwd_input(W + L.out1_w, cl, cl, sm + b.g_out, b.ldo, sm + b.g_o1, b.ldo, RO, false);
    __syncthreads();
    for (int it = tid; it < cl * RO; it += nt) {
        const int k = it / RO, row = it - k * RO;
        const float y = sm[b.o1 + k * b.ldo + row];
        sm[b.g_o1 + k * b.ldo + row] = (sm[b.g_o1 + k * b.ldo + row] + sm[b.g_out + k * b.ldo + row]) * (1.f - y * y);
    }
    __syncthreads();
    // ---- out0: o1 = tanh(W cat + b); g_o1 now holds the pre-activation gradient
    dense_bwd_weight(slab + L.out0_w, slab + L.out0_b, 2 * cl, cl, sm + b.cat, b.ldo, sm + b.g_o1, b.ldo, RO, accum);
    dense_bwd_input(W + L.out0_w, 2 * cl, cl, sm + b.g_o1, b.ldo, sm + b.g_cat, b.ldo, RO, false);
    __syncthreads();
    // g_cat[0:cl] = g_f3 ; g_cat[cl:2cl] -> g_s
    for (int it = tid; it < cl * RO; it += nt) {
        const int k = it / RO, row = it - k * RO;
        sm[b.g_s + k * b.ldo + row] = sm[b.g_cat + (cl + k) * b.ldo + row];
    }
    // ---- aff2: f3 = W f2 + b
    dense_bwd_weight(slab + L.aff2_w, slab + L.aff2_b, cl, cl, sm + b.f2, b.ldo, sm + b.g_cat, b.ldo, RO, accum);
    dense_bwd_input(W + L.aff2_w, cl, cl, sm + b.g_cat, b.ldo, sm + b.g_f2, b.ldo, RO, false);
    __syncthreads();
    // ---- aff1: f2 = tanh(W f1 + b) + f1 ; tanh output = f2 - f1
    for (int it = tid; it < cl * RO; it += nt) {
        const int k = it / RO, row = it - k * RO;
        const float th = sm[b.f2 + k * b.ldo + row] - sm[b.f1 + k * b.ldo + row];
        const float g = sm[b.g_f2 + k * b.ldo + row];
        sm[b.g_f1 + k * b.ldo + row] = g;                       // residual path
        sm[b.g_f2 + k * b.ldo + row] = g * (1.f - th * th);     // pre-activation gradient
    }
    __syncthreads();
    dense_bwd_weight(slab + L.aff1_w, slab + L.aff1_b, cl, cl, sm + b.f1, b.ldo, sm + b.g_f2, b.ldo, RO, accum);
    dense_bwd_input(W + L.aff1_w, cl, cl, sm + b.g_f2, b.ldo, sm + b.g_f1, b.ldo, RO, true);
    __syncthreads();
    // ---- aff0: f1 = tanh(W d + b)
    scale_by_act_grad(sm + b.g_f1, b.ldo, sm + b.f1, b.ldo, cl, RO, ACT_TANH, nl);
    __syncthreads();
    dense_bwd_weight(slab + L.aff0_w, slab + L.aff0_b, cl, cl, sm + b.d, b.ldo, sm + b.g_f1, b.ldo, RO, accum);
    dense_bwd_input(W + L.aff0_w, cl, cl, sm + b.g_f1, b.ldo, sm + b.g_d, b.ldo, RO, false);
    __syncthreads();
    // ---- reward head (dynamics.py:254-263)
    if (c.reward) {
        for (int sq = tid; sq < nseq; sq += nt) {
            const float r = sm[b.rew + sq];
            const float g = g_reward ? __ldg(g_reward + sq * g_reward_stride) : 0.f;
            sm[b.g_rew + sq] = g * r * (1.f - r);
        }
        __syncthreads();
        dense_bwd_weight(slab + L.rew14_w, slab + L.rew14_b, cl / 4, 1, sm + b.r3, b.lds, sm + b.g_rew, b.lds, nseq, accum);
        dense_bwd_input(W + L.rew14_w, cl / 4, 1, sm + b.g_rew, b.lds, sm + b.g_r3, b.lds, nseq, false);
        __syncthreads();
        scale_by_act_grad(sm + b.g_r3, b.lds, sm + b.r3, b.lds, cl / 4, nseq, ACT_RELU, nl);
        __syncthreads();
        dense_bwd_weight(slab + L.rew12_w, slab + L.rew12_b, cl / 2, cl / 4, sm + b.r2, b.lds, sm + b.g_r3, b.lds, nseq, accum);
        dense_bwd_input(W + L.rew12_w, cl / 2, cl / 4, sm + b.g_r3, b.lds, sm + b.g_r2, b.lds, nseq, false);
        __syncthreads();
        scale_by_act_grad(sm + b.g_r2, b.lds, sm + b.r2, b.lds, cl / 2, nseq, ACT_RELU, nl);
        __syncthreads();
        dense_bwd_weight(slab + L.rew10_w, slab + L.rew10_b, cl, cl / 2, sm + b.rsum, b.lds, sm + b.g_r2, b.lds, nseq, accum);
        dense_bwd_input(W + L.rew10_w, cl, cl / 2, sm + b.g_r2, b.lds, sm + b.g_rsum, b.lds, nseq, false);
        __syncthreads();
        for (int it = tid; it < cl * RO; it += nt) {
            const int k = it / RO, row = it - k * RO;
            sm[b.g_rh1 + k * b.ldo + row] = sm[b.g_rsum + k * b.lds + row / O];
        }
        __syncthreads();
        dense_bwd_weight(slab + L.rew02_w, slab + L.rew02_b, cl, cl, sm + b.rh0, b.ldo, sm + b.g_rh1, b.ldo, RO, accum);
        dense_bwd_input(W + L.rew02_w, cl, cl, sm + b.g_rh1, b.ldo, sm + b.g_rh0, b.ldo, RO, false);
        __syncthreads();
        scale_by_act_grad(sm + b.g_rh0, b.ldo, sm + b.rh0, b.ldo, cl, RO, ACT_RELU, nl);
        __syncthreads();
        dense_bwd_weight(slab + L.rew00_w, slab + L.rew00_b, cl, cl, sm + b.d, b.ldo, sm + b.g_rh0, b.ldo, RO, accum);
        dense_bwd_input(W + L.rew00_w, cl, cl, sm + b.g_rh0, b.ldo, sm + b.g_d, b.ldo, RO, true);
        __syncthreads();
    }
    // ---- aggregation: d_i = self_i + sum_j mask rel_ij att_ij
    for (int p = tid; p < RP; p += nt) {
        const int sq = p / (O * O), ij = p - sq * O * O, i = ij / O, j = ij - i * O;
        const float mask = (i == j) ? 0.f : 1.f;
        float acc = 0.f;
        for (int k = 0; k < cl; ++k) acc = fmaf(sm[b.g_d + k * b.ldo + sq * O + i], sm[b.rel + k * b.ldp + p], acc);
        // att = exp(lin): d att / d lin = att
        sm[b.g_att + p] = acc * mask * sm[b.att + p];
    }
    __syncthreads();
    for (int it = tid; it < cl * RP; it += nt) {
        const int k = it / RP, p = it - k * RP;
        const int sq = p / (O * O), ij = p - sq * O * O, i = ij / O, j = ij - i * O;
        const float mask = (i == j) ? 0.f : 1.f;
        // overwrite rel with its gradient (rel itself is no longer needed)
        sm[b.rel + k * b.ldp + p] = sm[b.g_d + k * b.ldo + sq * O + i] * mask * sm[b.att + p];
    }
    for (int it = tid; it < cl * RO; it += nt) {
        const int k = it / RO, row = it - k * RO;
        sm[b.g_self + k * b.ldo + row] = sm[b.g_d + k * b.ldo + row];
    }
    __syncthreads();
    // ---- att2 (cl -> 1, exp) and rel2 (rel = W r1 + b + r1)
    dense_bwd_weight(slab + L.att2_w, slab + L.att2_b, cl, 1, sm + b.a1, b.ldp, sm + b.g_att, b.ldp, RP, accum);
    dense_bwd_input(W + L.att2_w, cl, 1, sm + b.g_att, b.ldp, sm + b.g_a1, b.ldp, RP, false);
    dense_bwd_weight(slab + L.rel2_w, slab + L.rel2_b, cl, cl, sm + b.r1, b.ldp, sm + b.rel, b.ldp, RP, accum);
    dense_bwd_input(W + L.rel2_w, cl, cl, sm + b.rel, b.ldp, sm + b.g_r1, b.ldp, RP, false);
    __syncthreads();
    for (int it = tid; it < cl * RP; it += nt) {
        const int k = it / RP, p = it - k * RP;
        sm[b.g_r1 + k * b.ldp + p] = (sm[b.g_r1 + k * b.ldp + p] + sm[b.rel + k * b.ldp + p]) *
                                     act_grad(sm[b.r1 + k * b.ldp + p], ACT_NL, nl);
        sm[b.g_a1 + k * b.ldp + p] *= act_grad(sm[b.a1 + k * b.ldp + p], ACT_NL, nl);
    }
    __syncthreads();
    // ---- rel1 / att1 (2cl -> cl); input gradients overwrite... need r0/a0 for the weight
    // gradient first, so compute weight gradients, then input gradients into comb-sized scratch
    dense_bwd_weight(slab + L.rel1_w, slab + L.rel1_b, 2 * cl, cl, sm + b.ra0, b.ldp, sm + b.g_r1, b.ldp, RP, accum);
    dense_bwd_weight(slab + L.att1_w, slab + L.att1_b, 2 * cl, cl, sm + b.ra0 + 2 * cl * b.ldp, b.ldp, sm + b.g_a1, b.ldp, RP, accum);
    __syncthreads();
    // in place: ra0 <- (W^T g) * act'(ra0)
    dense_bwd_input(W + L.rel1_w, 2 * cl, cl, sm + b.g_r1, b.ldp, sm + b.ra0, b.ldp, RP, false,
                    sm + b.ra0, b.ldp, ACT_NL, nl);
    dense_bwd_input(W + L.att1_w, 2 * cl, cl, sm + b.g_a1, b.ldp, sm + b.ra0 + 2 * cl * b.ldp, b.ldp, RP, false,
                    sm + b.ra0 + 2 * cl * b.ldp, b.ldp, ACT_NL, nl);
    __syncthreads();
    // ---- rel0|att0 (2cl+1 -> 4cl)
    dense_bwd_weight(slab + L.ra0_w, slab + L.ra0_b, 2 * cl + 1, 4 * cl, sm + b.comb, b.ldp, sm + b.ra0, b.ldp, RP, accum);
    __syncthreads();
    dense_bwd_input(W + L.ra0_w, 2 * cl + 1, 4 * cl, sm + b.ra0, b.ldp, sm + b.comb, b.ldp, RP, false);
    __syncthreads();
    // ---- scatter pair-input gradients to the objects (comb now holds g_comb)
    for (int it = tid; it < cl * RO; it += nt) {
        const int k = it / RO, row = it - k * RO;
        const int sq = row / O, i = row - sq * O;
        float acc = 0.f;
        for (int j = 0; j < O; ++j) {
            acc += sm[b.comb + k * b.ldp + sq * O * O + i * O + j];            // as first argument
            acc += sm[b.comb + (cl + k) * b.ldp + sq * O * O + j * O + i];     // as second argument
        }
        if (k < 2) {
            // dist_ij = (x_i-x_j)^2 + (y_i-y_j)^2
            const float xi = sm[b.s + k * b.ldo + row];
            for (int j = 0; j < O; ++j) {
                const float xj = sm[b.s + k * b.ldo + sq * O + j];
                acc += 2.f * (xi - xj) * (sm[b.comb + 2 * cl * b.ldp + sq * O * O + i * O + j] +
                                          sm[b.comb + 2 * cl * b.ldp + sq * O * O + j * O + i]);
            }
        }
        sm[b.g_s + k * b.ldo + row] += acc;
    }
    // ---- self1: self = W h + b + h
    dense_bwd_weight(slab + L.self1_w, slab + L.self1_b, cl, cl, sm + b.h, b.ldo, sm + b.g_self, b.ldo, RO, accum);
    dense_bwd_input(W + L.self1_w, cl, cl, sm + b.g_self, b.ldo, sm + b.g_h, b.ldo, RO, false);
    __syncthreads();
    for (int it = tid; it < cl * RO; it += nt) {
        const int k = it / RO, row = it - k * RO;
        sm[b.g_h + k * b.ldo + row] = (sm[b.g_h + k * b.ldo + row] + sm[b.g_self + k * b.ldo + row]) *
                                      act_grad(sm[b.h + k * b.ldo + row], ACT_NL, nl);
    }
    __syncthreads();
    // ---- self0: h = phi(W s + b)
    dense_bwd_weight(slab + L.self0_w, slab + L.self0_b, cl, cl, sm + b.s, b.ldo, sm + b.g_h, b.ldo, RO, accum);
    dense_bwd_input(W + L.self0_w, cl, cl, sm + b.g_h, b.ldo, sm + b.g_s, b.ldo, RO, true);
    __syncthreads();
    // ---- encoder: s = [s_in[:lim], enc(s_in)[lim:]]
    // g_enc_out = g_s with the first lim rows zeroed (kept aside in g_h, free now)
    for (int it = tid; it < cl * RO; it += nt) {
        const int k = it / RO, row = it - k * RO;
        sm[b.g_h + k * b.ldo + row] = (k < c.lim_enc) ? 0.f : sm[b.g_s + k * b.ldo + row];
    }
    __syncthreads();
    dense_bwd_weight(slab + L.enc_w, slab + L.enc_b, L.in_dim, cl, sm + b.sin, b.ldo, sm + b.g_h, b.ldo, RO, accum);
    dense_bwd_input(W + L.enc_w, L.in_dim, cl, sm + b.g_h, b.ldo, sm + b.g_sin, b.ldo, RO, false);
    __syncthreads();
    // raw pass-through of the first lim_enc features
    for (int it = tid; it < c.lim_enc * RO; it += nt) {
        const int k = it / RO, row = it - k * RO;
        sm[b.g_sin + k * b.ldo + row] += sm[b.g_s + k * b.ldo + row];
    }
    if (c.action_dim > 0) {
        // emb = act_W^T a + b ; g_emb[o*4+e][seq] = g_sin[half+e][seq*O+o]
        for (int it = tid; it < nseq * O * 4; it += nt) {
            const int sq = it / (O * 4), nn = it - sq * (O * 4);
            sm[b.g_emb + nn * b.lds + sq] = sm[b.g_sin + (gnn_sdim(c) + (nn & 3)) * b.ldo + sq * O + (nn >> 2)];
        }
        __syncthreads();
        const int NA = O * 4;
        for (int it = tid; it < c.action_dim * NA; it += nt) {
            const int k = it / NA, nn = it - k * NA;
            float acc = 0.f;
            for (int sq = 0; sq < nseq; ++sq)
                acc = fmaf(__ldg(actions + sq * act_stride + k), sm[b.g_emb + nn * b.lds + sq], acc);
            if (accum) slab[L.act_w + k * NA + nn] += acc;
            else slab[L.act_w + k * NA + nn] = acc;
        }
        for (int nn = tid; nn < NA; nn += nt) {
            float acc = 0.f;
            for (int sq = 0; sq < nseq; ++sq) acc += sm[b.g_emb + nn * b.lds + sq];
            if (accum) slab[L.act_b + nn] += acc;
            else slab[L.act_b + nn] = acc;
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------
// backward kernel: recompute forward on chip, then reverse.  slab = per-CTA weight gradients.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) gnn_bwd_kernel(stove_gnn_cfg c, GnnLayout L, int seq, int64_t n,
                                                      int stage_w, const float* __restrict__ s,
                                                      const float* __restrict__ actions,
                                                      const float* __restrict__ app,
                                                      const float* __restrict__ weights,
                                                      const float* __restrict__ g_out,
                                                      const float* __restrict__ g_reward,
                                                      float* __restrict__ g_s, float* __restrict__ slabs) {
    extern __shared__ __align__(16) float smem[];
    const float* W = weights;
    float* sm = smem;
    if (stage_w) {
        stage_weights(weights, smem, L.total);
        W = smem;
        sm = smem + L.total;
    }
    const GnnBuf b = gnn_buffers(c, L.in_dim, seq, true);
    for (int i = threadIdx.x; i < b.total; i += blockDim.x) sm[i] = 0.f;
    float* slab = slabs + (int64_t)blockIdx.x * L.total;
    const int cl = c.cl, O = c.num_obj, nl = c.nonlin, half = gnn_sdim(c);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int64_t ngroups = (n + seq - 1) / seq;
    bool accum = false;
    for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x, accum = true) {
        const int64_t seq0 = grp * seq;
        const int nseq = (int)min((int64_t)seq, n - seq0);
        const int RO = nseq * O, RP = nseq * O * O;
        __syncthreads();
        load_inputs(c, b, sm, seq0, nseq, s, half, 0, app);
        __syncthreads();
        gnn_forward_core(c, L, b, W, sm, nseq, actions ? actions + seq0 * c.action_dim : nullptr, c.action_dim);
        // upstream gradient, feature-major
        for (int it = tid; it < RO * cl; it += nt) {
            const int row = it / cl, k = it - row * cl;
            sm[b.g_out + k * b.ldo + row] = __ldg(g_out + (seq0 * O + row) * cl + k);
        }
        __syncthreads();
        gnn_backward_core(c, L, b, W, sm, slab, accum, nseq, actions ? actions + seq0 * c.action_dim : nullptr,
                          c.action_dim, g_reward ? g_reward + seq0 : nullptr, 1);
        for (int it = tid; it < RO * half; it += nt) {
            const int row = it / half, k = it - row * half;
            g_s[(seq0 * O + row) * half + k] = sm[b.g_sin + k * b.ldo + row];
        }
    }
}

// ------------------------------------------------------------------------------------
// Fused dynamics-loop step of Stove.stove_forward (stove.py:696-713): GNN core, then
// constrain_z_dyn (dynamics.py:147-179), position integration, Gaussian fusion with the SuPAIR
// state + reparametrised sample + log q (full_state, stove.py:103-170) and the transition
// log-likelihood of the sample (transition_lik, stove.py:172-198) -- ~45 tensor ops per time
// step in the reference -- as the epilogue of the forward kernel and the prologue of the
// backward kernel.  All tensors are addressed as base + sequence * stride so the caller can
// pass time slices of (n, T, ...) tensors without copies.
// ------------------------------------------------------------------------------------

__device__ __forceinline__ void load_inputs_strided(const stove_gnn_cfg& c, const GnnBuf& b, float* sm, int nseq,
                                                    const float* __restrict__ z_prev, int64_t zss,
                                                    const float* __restrict__ app, int64_t ass) {
    const int cl = c.cl, O = c.num_obj, half = cl / 2, zd = half + 2;
    for (int it = threadIdx.x; it < nseq * O * half; it += blockDim.x) {
        const int row = it / half, k = it - row * half;
        const int sq = row / O, o = row - sq * O;
        sm[b.sin + k * b.ldo + row] = __ldg(z_prev + sq * zss + o * zd + 2 + k);
    }
    if (c.app_dim > 0) {
        const int a0 = half + (c.action_dim > 0 ? 4 : 0);
        for (int it = threadIdx.x; it < nseq * O * c.app_dim; it += blockDim.x) {
            const int row = it / c.app_dim, k = it - row * c.app_dim;
            const int sq = row / O, o = row - sq * O;
            sm[b.sin + (a0 + k) * b.ldo + row] = __ldg(app + sq * ass + o * c.app_dim + k);
        }
    }
}

// forward quantities of one (row, feature j) of the fused epilogue
struct FuseVal {
    float zd, sd, zdyn, mean, std, m_sup, s_sup, scale;
};
__device__ __forceinline__ FuseVal fuse_forward(const stove_gnn_cfg& c, const FuseCfg& f, const GnnBuf& b,
                                                const float* sm, int row, int j, const float* sup6,
                                                const float* sstd6) {
    FuseVal v;
    const int half = c.cl / 2;
    if (j < 2) {
        v.zd = v.sd = v.zdyn = 0.f; v.scale = 1.f; v.m_sup = v.s_sup = 0.f;
        v.mean = __ldg(sup6 + j);
        v.std = __ldg(sstd6 + j);
        return v;
    }
    const int i = j - 2;
    v.scale = f.scale[i < 2 ? 0 : (i < 4 ? 1 : 2)];
    v.zd = 2.f * sigmoidf_(sm[b.out + i * b.ldo + row]) - 1.f;
    v.sd = v.scale * sigmoidf_(sm[b.out + (half + i) * b.ldo + row]);
    v.zdyn = v.zd + (i < 2 ? sm[b.sin + i * b.ldo + row] : 0.f);
    if (i < 4) {
        v.m_sup = __ldg(sup6 + 2 + i);
        v.s_sup = __ldg(sstd6 + 2 + i);
        const float A = v.s_sup * v.s_sup, B = v.sd * v.sd, D = A + B;
        v.mean = (A * v.zdyn + B * v.m_sup) / D;
        v.std = v.sd * v.s_sup / sqrtf(D);
    } else {
        v.m_sup = v.s_sup = 0.f;
        v.mean = v.zdyn;
        v.std = v.sd;
    }
    return v;
}

__global__ void __launch_bounds__(512) dynstep_fwd_kernel(stove_gnn_cfg c, GnnLayout L, FuseCfg f, int seq,
                                                          int64_t n, stove_dynstep_io io,
                                                          const float* __restrict__ weights) {
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;
    float* sm = smem + L.total;
    const GnnBuf b = gnn_buffers(c, L.in_dim, seq, false);
    stage_weights(weights, Ws, L.total);
    for (int i = threadIdx.x; i < b.total; i += blockDim.x) sm[i] = 0.f;
    const int cl = c.cl, O = c.num_obj, half = cl / 2, zdim = half + 2;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int64_t ngroups = (n + seq - 1) / seq;
    for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int64_t seq0 = grp * seq;
        const int nseq = (int)min((int64_t)seq, n - seq0);
        const int RO = nseq * O;
        __syncthreads();
        load_inputs_strided(c, b, sm, nseq, io.z_prev + seq0 * io.z_prev_ss, io.z_prev_ss,
                            io.app ? io.app + seq0 * io.app_ss : nullptr, io.app_ss);
        __syncthreads();
        gnn_forward_core(c, L, b, Ws, sm, nseq, io.actions ? io.actions + seq0 * io.act_ss : nullptr, io.act_ss);
        // epilogue; per-item log q / transition terms go to scratch (the dead `cat` buffer) and are
        // summed per sequence in a fixed order (deterministic)
        float* s_lq = sm + b.cat;
        float* s_tr = s_lq + RO * zdim;
        for (int it = tid; it < RO * zdim; it += nt) {
            const int row = it / zdim, j = it - row * zdim;
            const int sq = row / O, o = row - sq * O;
            const int64_t gs = seq0 + sq;
            const FuseVal v = fuse_forward(c, f, b, sm, row, j, io.sup + gs * io.sup_ss + o * 6,
                                           io.sup_std + gs * io.sup_ss + o * 6);
            const float e = __ldg(io.eps + gs * io.eps_ss + o * zdim + j);
            const float z = v.mean + v.std * e;
            io.z_out[gs * io.z_out_ss + o * zdim + j] = z;
            if (io.z_std) io.z_std[gs * io.z_std_ss + o * zdim + j] = v.std;
            s_lq[it] = -0.5f * e * e - logf(v.std) - HALF_LOG_2PI;
            float tr = 0.f;
            if (j >= 2) {
                const int i = j - 2;
                io.z_dyn[gs * io.zdyn_ss + o * half + i] = v.zdyn;
                io.z_dyn_std[gs * io.zdyn_ss + o * half + i] = v.sd;
                const float st = f.trans_std[i], d = z - v.zdyn;
                tr = -(d * d) / (2.f * st * st) - logf(st) - HALF_LOG_2PI;
            }
            s_tr[it] = tr;
        }
        __syncthreads();
        for (int sq = tid; sq < nseq; sq += nt) {
            float a = 0.f, t = 0.f;
            for (int q = sq * O * zdim; q < (sq + 1) * O * zdim; ++q) { a += s_lq[q]; t += s_tr[q]; }
            io.logq[(seq0 + sq) * io.sc_ss] = a;
            io.trans[(seq0 + sq) * io.sc_ss] = t;
            if (c.reward && io.reward) io.reward[(seq0 + sq) * io.sc_ss] = sm[b.rew + sq];
        }
    }
}

__global__ void __launch_bounds__(512) dynstep_bwd_kernel(stove_gnn_cfg c, GnnLayout L, FuseCfg f, int seq,
                                                          int64_t n, int stage_w, int slab_accumulate,
                                                          stove_dynstep_io io,
                                                          const float* __restrict__ weights,
                                                          float* __restrict__ slabs) {
    extern __shared__ __align__(16) float smem[];
    const float* W = weights;
    float* sm = smem;
    if (stage_w) {
        stage_weights(weights, smem, L.total);
        W = smem;
        sm = smem + L.total;
    }
    const GnnBuf b = gnn_buffers(c, L.in_dim, seq, true);
    for (int i = threadIdx.x; i < b.total; i += blockDim.x) sm[i] = 0.f;
    float* slab = slabs + (int64_t)blockIdx.x * L.total;
    const int cl = c.cl, O = c.num_obj, half = cl / 2, zdim = half + 2;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int64_t ngroups = (n + seq - 1) / seq;
    bool accum = slab_accumulate != 0;
    for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x, accum = true) {
        const int64_t seq0 = grp * seq;
        const int nseq = (int)min((int64_t)seq, n - seq0);
        const int RO = nseq * O;
        __syncthreads();
        load_inputs_strided(c, b, sm, nseq, io.z_prev + seq0 * io.z_prev_ss, io.z_prev_ss,
                            io.app ? io.app + seq0 * io.app_ss : nullptr, io.app_ss);
        __syncthreads();
        gnn_forward_core(c, L, b, W, sm, nseq, io.actions ? io.actions + seq0 * io.act_ss : nullptr, io.act_ss);
        // prologue: gradients of (sample, log q, transition lik) -> raw network output, SuPAIR inputs
        for (int it = tid; it < RO * zdim; it += nt) {
            const int row = it / zdim, j = it - row * zdim;
            const int sq = row / O, o = row - sq * O;
            const int64_t gs = seq0 + sq;
            const FuseVal v = fuse_forward(c, f, b, sm, row, j, io.sup + gs * io.sup_ss + o * 6,
                                           io.sup_std + gs * io.sup_ss + o * 6);
            const float e = __ldg(io.eps + gs * io.eps_ss + o * zdim + j);
            const float z = v.mean + v.std * e;
            float gz = 0.f;
            if (io.g_z_a) gz += __ldg(io.g_z_a + gs * io.g_z_a_ss + o * zdim + j);
            if (io.g_z_b) gz += __ldg(io.g_z_b + gs * io.g_z_b_ss + o * zdim + j);
            const float glq = io.g_logq ? __ldg(io.g_logq + gs * io.g_sc_ss) : 0.f;
            const float gtr = io.g_trans ? __ldg(io.g_trans + gs * io.g_sc_ss) : 0.f;
            float g_zdyn = 0.f;
            if (j >= 2) {
                const float st = f.trans_std[j - 2];
                const float t = (z - v.zdyn) / (st * st) * gtr;
                gz -= t;
                g_zdyn = t;
            }
            const float g_mean = gz, g_std = gz * e - glq / v.std;
            float* gsup = io.g_sup + gs * io.g_sup_ss + o * 6;
            float* gsst = io.g_sup_std + gs * io.g_sup_ss + o * 6;
            float* gzp = io.g_z_prev + gs * io.g_z_prev_ss + o * zdim;
            if (j < 2) {
                gsup[j] = g_mean;
                gsst[j] = g_std;
                gzp[j] = 0.f;
                continue;
            }
            const int i = j - 2;
            float g_sd;
            if (i < 4) {
                const float A = v.s_sup * v.s_sup, B = v.sd * v.sd, D = A + B, rD = 1.f / D, rD15 = rD / sqrtf(D);
                g_zdyn += g_mean * A * rD;
                gsup[2 + i] = g_mean * B * rD;
                const float gA = g_mean * (v.zdyn - v.mean) * rD, gB = g_mean * (v.m_sup - v.mean) * rD;
                g_sd = g_std * v.s_sup * A * rD15 + gB * 2.f * v.sd;
                gsst[2 + i] = g_std * v.sd * B * rD15 + gA * 2.f * v.s_sup;
            } else {
                g_zdyn += g_mean;
                g_sd = g_std;
            }
            sm[b.g_out + i * b.ldo + row] = g_zdyn * (1.f - v.zd * v.zd) * 0.5f;
            sm[b.g_out + (half + i) * b.ldo + row] = g_sd * v.sd * (1.f - v.sd / v.scale);
            if (i < 2) gzp[2 + i] = g_zdyn;          // direct path pos_t = pos_{t-1} + delta
        }
        __syncthreads();
        gnn_backward_core(c, L, b, W, sm, slab, accum, nseq, io.actions ? io.actions + seq0 * io.act_ss : nullptr,
                          io.act_ss, io.g_reward ? io.g_reward + seq0 * io.g_sc_ss : nullptr, io.g_sc_ss);
        for (int it = tid; it < RO * half; it += nt) {
            const int row = it / half, k = it - row * half;
            const int sq = row / O, o = row - sq * O;
            float* gzp = io.g_z_prev + (seq0 + sq) * io.g_z_prev_ss + o * zdim + 2 + k;
            const float g = sm[b.g_sin + k * b.ldo + row];
            *gzp = (k < 2) ? *gzp + g : g;
        }
    }
}

int stove_team_rollout(const stove_gnn_cfg* cfg, const GnnLayout& L, int64_t n, int num, const float* z_last,
                       const float* actions, int action_len, const float* app, const float* weights,
                       const float* noise, float pos_var, float vel_std, float latent_std, float* z_out,
                       float* std_out, float* logq_out, float* rewards, cudaStream_t st);

// g_w[i] (=, +=) sum_b slabs[b][i]
__global__ void gnn_reduce_slabs_acc_kernel(const float* __restrict__ slabs, int nslab, int total,
                                            float* __restrict__ g_w, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float acc = accumulate ? g_w[i] : 0.f;
    for (int s = 0; s < nslab; ++s) acc += slabs[(int64_t)s * total + i];
    g_w[i] = acc;
}

// g_w[i] = sum_b slabs[b][i]
__global__ void gnn_reduce_slabs_kernel(const float* __restrict__ slabs, int nslab, int total,
                                        float* __restrict__ g_w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float acc = 0.f;
    for (int s = 0; s < nslab; ++s) acc += slabs[(int64_t)s * total + i];
    g_w[i] = acc;
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------


// largest number of sequences per CTA that fits (optionally with staged weights), <= want

// threads per CTA of the step kernels: the device code only uses blockDim.x.  One CTA per SM (shared memory), so more
// warps only hide latency: 512 threads pay from 5 objects on (multiball, 25+ pair rows per sequence: dynstep_bwd
// 0.85 -> 0.81 ms per step at 6 objects, 2.52 -> 2.31 ms at 9); option gnn_threads overrides (256 / 384 / 512)
static int gnn_threads(const stove_gnn_cfg* c) {
    const int t = stove_opt(OPT_GNN_THREADS);
    if (t == 256 || t == 384 || t == 512) return t;
    return c->num_obj >= 5 ? 512 : 256;
}

static int pick_seq(const stove_gnn_cfg* c, const GnnLayout& L, bool bwd, bool stage, int want) {
    // tuning override (sequences per CTA): options gnn_seq_fwd / gnn_seq_bwd
    if (stove_opt(bwd ? OPT_GNN_SEQ_BWD : OPT_GNN_SEQ_FWD) > 0) want = stove_opt(bwd ? OPT_GNN_SEQ_BWD : OPT_GNN_SEQ_FWD);
    for (int seq = want; seq >= 1; --seq) {
        GnnBuf b = gnn_buffers(*c, L.in_dim, seq, bwd);
        size_t bytes = sizeof(float) * ((size_t)b.total + (stage ? L.total : 0));
        if (bytes <= kMaxSmem) return seq;
    }
    return 0;
}

extern "C" int64_t stove_gnn_weight_count(const stove_gnn_cfg* cfg) {
    if (gnn_check(cfg)) return -1;
    return gnn_layout(cfg).total;
}

// exported for the Python packer: float offsets of every segment, in the order of GnnLayout
extern "C" int stove_gnn_weight_offsets(const stove_gnn_cfg* cfg, int32_t* out, int max_out) {
    int rc = gnn_check(cfg);
    if (rc) return rc;
    GnnLayout L = gnn_layout(cfg);
    const int32_t* src = &L.act_w;
    const int count = (int)(&L.total - &L.act_w) + 1;
    STOVE_CHECK_ARG(out && max_out >= count, "offset buffer too small");
    for (int i = 0; i < count; ++i) out[i] = src[i];
    return count;
}

static int gnn_target_ctas(int64_t n, int seq) {
    int64_t groups = (n + seq - 1) / seq;
    int64_t cap = 148 * 2;
    return (int)(groups < cap ? groups : cap);
}

extern "C" int stove_gnn_fwd(const stove_gnn_cfg* cfg, int64_t n, const float* s, const float* actions,
                             const float* app, const float* weights, float* out, float* reward, void* stream) {
    int rc = gnn_check(cfg);
    if (rc) return rc;
    STOVE_CHECK_ARG(n >= 0 && s && weights && out, "null pointer");
    STOVE_CHECK_ARG(((uintptr_t)weights & 15) == 0, "weights must be 16-byte aligned");
    STOVE_CHECK_ARG((cfg->action_dim > 0) == (actions != nullptr), "actions do not match cfg.action_dim");
    STOVE_CHECK_ARG((cfg->app_dim > 0) == (app != nullptr), "appearances do not match cfg.app_dim");
    if (n == 0) return STOVE_OK;
    GnnLayout L = gnn_layout(cfg);
    // few sequences per CTA: this launch is latency bound, spread it over the chip
    int want = (int)((n + 147) / 148);
    if (want < 1) want = 1;
    if (want > 8) want = 8;
    const int seq = pick_seq(cfg, L, false, true, want);
    if (seq == 0) { stove_set_error("stove_gnn_fwd: configuration does not fit in shared memory"); return STOVE_ERR_UNSUPPORTED; }
    GnnBuf b = gnn_buffers(*cfg, L.in_dim, seq, false);
    const size_t smem = sizeof(float) * ((size_t)b.total + L.total);
    STOVE_CUDA(cudaFuncSetAttribute(gnn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    STOVE_KERNEL(K_GNN_FWD, (cudaStream_t)stream, gnn_fwd_kernel<<<gnn_target_ctas(n, seq), gnn_threads(cfg), smem, (cudaStream_t)stream>>>(*cfg, L, seq, n, s, actions, app,
                                                                                   weights, out, reward));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

struct GnnBwdPlan {
    int seq, stage, ctas;
    size_t smem;
};

static GnnBwdPlan gnn_bwd_plan(const stove_gnn_cfg* cfg, const GnnLayout& L, int64_t n) {
    GnnBwdPlan p;
    int want = (int)((n + 147) / 148);
    if (want < 1) want = 1;
    if (want > 4) want = 4;
    p.stage = 1;
    p.seq = pick_seq(cfg, L, true, true, want);
    if (p.seq == 0) {
        p.stage = 0;
        p.seq = pick_seq(cfg, L, true, false, want);
    }
    if (p.seq == 0) { p.ctas = 0; p.smem = 0; return p; }
    GnnBuf b = gnn_buffers(*cfg, L.in_dim, p.seq, true);
    p.smem = sizeof(float) * ((size_t)b.total + (p.stage ? L.total : 0));
    p.ctas = gnn_target_ctas(n, p.seq);
    return p;
}

extern "C" size_t stove_gnn_bwd_workspace(const stove_gnn_cfg* cfg, int64_t n) {
    if (gnn_check(cfg) || n <= 0) return 0;
    GnnLayout L = gnn_layout(cfg);
    GnnBwdPlan p = gnn_bwd_plan(cfg, L, n);
    return sizeof(float) * (size_t)p.ctas * L.total;
}

extern "C" int stove_gnn_bwd(const stove_gnn_cfg* cfg, int64_t n, const float* s, const float* actions,
                             const float* app, const float* weights, const float* g_out,
                             const float* g_reward, float* g_s, float* g_weights, void* workspace,
                             void* stream) {
    int rc = gnn_check(cfg);
    if (rc) return rc;
    STOVE_CHECK_ARG(n >= 0 && s && weights && g_out && g_s && g_weights && workspace, "null pointer");
    STOVE_CHECK_ARG(((uintptr_t)weights & 15) == 0, "weights must be 16-byte aligned");
    STOVE_CHECK_ARG((cfg->action_dim > 0) == (actions != nullptr), "actions do not match cfg.action_dim");
    STOVE_CHECK_ARG((cfg->app_dim > 0) == (app != nullptr), "appearances do not match cfg.app_dim");
    GnnLayout L = gnn_layout(cfg);
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        STOVE_CUDA(cudaMemsetAsync(g_weights, 0, sizeof(float) * L.total, st));
        return STOVE_OK;
    }
    GnnBwdPlan p = gnn_bwd_plan(cfg, L, n);
    if (p.seq == 0) { stove_set_error("stove_gnn_bwd: configuration does not fit in shared memory"); return STOVE_ERR_UNSUPPORTED; }
    STOVE_CUDA(cudaFuncSetAttribute(gnn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    // padding floats of the slabs are never written: clear them once so the reduction is clean
    STOVE_CUDA(cudaMemsetAsync(workspace, 0, sizeof(float) * (size_t)p.ctas * L.total, st));
    STOVE_KERNEL(K_GNN_BWD, st, gnn_bwd_kernel<<<p.ctas, gnn_threads(cfg), p.smem, st>>>(*cfg, L, p.seq, n, p.stage, s, actions, app, weights, g_out, g_reward,
                                                g_s, (float*)workspace));
    STOVE_LAUNCH_CHECK();
    STOVE_KERNEL(K_GNN_REDUCE_SLABS, st, gnn_reduce_slabs_kernel<<<(L.total + 255) / 256, 256, 0, st>>>((const float*)workspace, p.ctas, L.total, g_weights));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}


extern "C" int stove_dynstep_fwd(const stove_gnn_cfg* cfg, const stove_fuse_cfg* fuse, int64_t n,
                                 const stove_dynstep_io* io, const float* weights, void* stream) {
    int rc = gnn_check(cfg);
    if (rc) return rc;
    STOVE_CHECK_ARG(gnn_default_state(cfg), "state_dim != cl/2 is served by stove_gnn_fwd / stove_gnn_bwd only");
    STOVE_CHECK_ARG(fuse && io && weights && n >= 0, "null pointer");
    STOVE_CHECK_ARG(io->z_prev && io->sup && io->sup_std && io->eps && io->z_out && io->z_dyn && io->z_dyn_std &&
                        io->logq && io->trans, "null tensor in stove_dynstep_io");
    STOVE_CHECK_ARG((cfg->action_dim > 0) == (io->actions != nullptr), "actions do not match cfg.action_dim");
    STOVE_CHECK_ARG((cfg->app_dim > 0) == (io->app != nullptr), "appearances do not match cfg.app_dim");
    STOVE_CHECK_ARG(cfg->cl <= 64, "cl too large");
    if (n == 0) return STOVE_OK;
    GnnLayout L = gnn_layout(cfg);
    int want = (int)((n + 147) / 148);
    if (want < 1) want = 1;
    if (want > 8) want = 8;
    const int seq = pick_seq(cfg, L, false, true, want);
    if (seq == 0) { stove_set_error("stove_dynstep_fwd: configuration does not fit in shared memory"); return STOVE_ERR_UNSUPPORTED; }
    GnnBuf b = gnn_buffers(*cfg, L.in_dim, seq, false);
    const size_t smem = sizeof(float) * ((size_t)b.total + L.total);
    STOVE_CUDA(cudaFuncSetAttribute(dynstep_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaStream_t st = (cudaStream_t)stream;
    STOVE_KERNEL(K_DYNSTEP_FWD, st, dynstep_fwd_kernel<<<gnn_target_ctas(n, seq), gnn_threads(cfg), smem, st>>>(
        *cfg, L, make_fuse(cfg, fuse), seq, n, *io, weights));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" int stove_dynstep_bwd(const stove_gnn_cfg* cfg, const stove_fuse_cfg* fuse, int64_t n,
                                 const stove_dynstep_io* io, const float* weights, float* g_weights,
                                 int first, int last, void* workspace, void* stream) {
    int rc = gnn_check(cfg);
    if (rc) return rc;
    STOVE_CHECK_ARG(gnn_default_state(cfg), "state_dim != cl/2 is served by stove_gnn_fwd / stove_gnn_bwd only");
    STOVE_CHECK_ARG(fuse && io && weights && g_weights && workspace && n >= 0, "null pointer");
    STOVE_CHECK_ARG(io->z_prev && io->sup && io->sup_std && io->eps && io->g_z_prev && io->g_sup && io->g_sup_std,
                    "null tensor in stove_dynstep_io");
    STOVE_CHECK_ARG((cfg->action_dim > 0) == (io->actions != nullptr), "actions do not match cfg.action_dim");
    STOVE_CHECK_ARG((cfg->app_dim > 0) == (io->app != nullptr), "appearances do not match cfg.app_dim");
    if (n == 0) return STOVE_OK;
    GnnLayout L = gnn_layout(cfg);
    GnnBwdPlan p = gnn_bwd_plan(cfg, L, n);
    if (p.seq == 0) { stove_set_error("stove_dynstep_bwd: configuration does not fit in shared memory"); return STOVE_ERR_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    STOVE_CUDA(cudaFuncSetAttribute(dynstep_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    // the per-CTA slabs accumulate over the time steps of one backward pass (same n => same grid):
    // cleared before the first call, reduced into g_weights after the last
    if (first) STOVE_CUDA(cudaMemsetAsync(workspace, 0, sizeof(float) * (size_t)p.ctas * L.total, st));
    STOVE_KERNEL(K_DYNSTEP_BWD, st, dynstep_bwd_kernel<<<p.ctas, gnn_threads(cfg), p.smem, st>>>(
        *cfg, L, make_fuse(cfg, fuse), p.seq, n, p.stage, first ? 0 : 1, *io, weights, (float*)workspace));
    STOVE_LAUNCH_CHECK();
    if (last) {
        STOVE_KERNEL(K_GNN_REDUCE_SLABS, st, gnn_reduce_slabs_acc_kernel<<<(L.total + 255) / 256, 256, 0, st>>>(
            (const float*)workspace, p.ctas, L.total, g_weights, 0));
        STOVE_LAUNCH_CHECK();
    }
    return STOVE_OK;
}

extern "C" int stove_gnn_rollout(const stove_gnn_cfg* cfg, int64_t n, int num, const float* z_last,
                                 const float* actions, int action_len, const float* app,
                                 const float* weights, const float* noise, float pos_var, float vel_std,
                                 float latent_std, float* z_out, float* std_out, float* logq_out,
                                 float* rewards, void* stream) {
    int rc = gnn_check(cfg);
    if (rc) return rc;
    STOVE_CHECK_ARG(gnn_default_state(cfg), "state_dim != cl/2 is served by stove_gnn_fwd / stove_gnn_bwd only");
    STOVE_CHECK_ARG(n >= 0 && num >= 0 && z_last && weights && z_out, "null pointer");
    STOVE_CHECK_ARG(((uintptr_t)weights & 15) == 0, "weights must be 16-byte aligned");
    STOVE_CHECK_ARG((cfg->action_dim > 0) == (actions != nullptr), "actions do not match cfg.action_dim");
    STOVE_CHECK_ARG(!actions || action_len > 0, "action_len must be positive");
    STOVE_CHECK_ARG((cfg->app_dim > 0) == (app != nullptr), "appearances do not match cfg.app_dim");
    STOVE_CHECK_ARG(!(logq_out && !noise), "logq_out requires noise");
    if (n == 0 || num == 0) return STOVE_OK;
    GnnLayout L = gnn_layout(cfg);
    {   // fast path: warp teams (dynloop.cu); returns 1 when the shape is not covered
        const int frc = stove_team_rollout(cfg, L, n, num, z_last, actions, action_len, app, weights, noise, pos_var,
                                           vel_std, latent_std, z_out, std_out, logq_out, rewards, (cudaStream_t)stream);
        if (frc != 1) return frc;
    }
    // one persistent CTA per SM: spread the sequences evenly over 148 CTAs
    int want = (int)((n + 147) / 148);
    if (want < 1) want = 1;
    if (want > 16) want = 16;
    const int seq = pick_seq(cfg, L, false, true, want);
    if (seq == 0) { stove_set_error("stove_gnn_rollout: configuration does not fit in shared memory"); return STOVE_ERR_UNSUPPORTED; }
    GnnBuf b = gnn_buffers(*cfg, L.in_dim, seq, false);
    const size_t smem = sizeof(float) * ((size_t)b.total + L.total);
    STOVE_CUDA(cudaFuncSetAttribute(gnn_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t groups = (n + seq - 1) / seq;
    const int ctas = (int)(groups < 148 * 2 ? groups : 148 * 2);
    STOVE_KERNEL(K_GNN_ROLLOUT, (cudaStream_t)stream, gnn_rollout_kernel<<<ctas, 512, smem, (cudaStream_t)stream>>>(*cfg, L, seq, n, num, z_last, actions, action_len,
                                                                  app, weights, noise, pos_var, vel_std, latent_std,
                                                                  z_out, std_out, logq_out, rewards));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}
