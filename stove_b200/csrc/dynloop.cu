// Warp-team kernels for the graph-network dynamics (O = 3 objects, cl = 32): the whole
// dynamics loop of Stove.stove_forward (model/video_prediction/stove.py:696-713) in ONE launch
// (forward) and three launches (backward), and the persistent rollout of Stove.rollout
// (stove.py:777-861).  The arithmetic is that of Dynamics.forward / core / constrain_z_dyn
// (dynamics.py:147-265), Stove.full_state (stove.py:103-170) and transition_lik (:172-198).
//
// Why warp teams: one sequence has 9 pair rows and 3 object rows -- far too little for a CTA,
// and ncu on the CTA-wide kernels (gnn.cu) showed 10x more issued instructions than useful FMAs
// (runtime index arithmetic, padded rows, idle lanes) plus a CTA barrier per layer.  Here a TEAM
// of NW warps (1 or 2) owns one sequence for all its time steps:
//   * activations live in the team's private slice of shared memory, feature-major, rows padded
//     to 4 (objects) / 12 (pairs) so that a whole row group is one or three LDS.128 broadcasts;
//   * lane = output feature in the forward pass, lane = input feature in the backward pass; the
//     weights are staged once per CTA with row stride N + 1, which makes BOTH access patterns
//     bank-conflict free (W[k][lane] and W[lane][n]);
//   * layers are separated by __syncwarp (NW = 1) or a 64-thread named barrier (NW = 2, where the
//     rel / att branches of the relation network run on different warps);
//   * the first pair layer is factorised: W^T [s_i, s_j, d_ij] = Wa^T s_i + Wb^T s_j + w_d d_ij,
//     i.e. 2 x 3 object-row products instead of 9 pair-row products (2.8x fewer FMAs in the largest
//     layer, forward and backward).  Same math, different summation order.
//   * the state never leaves the chip between time steps.
// Backward: the chain kernel recomputes each step's forward from the saved z_{t-1}, propagates the
// state gradient to the previous step on chip, and writes every layer's (input, pre-activation
// gradient) pair as one contiguous record per (sequence, step).  Weight gradients are NOT on the
// sequential chain: a second, fully parallel kernel turns all records into per-CTA gradient slabs
// held in registers (one pass, deterministic), a third sums the slabs.
#include <string.h>
#include "gnn_common.cuh"

namespace tk {
constexpr int O = 3, P = 9, PR = 12, ORW = 4, CL = 32, HALF = 16, ZD = 18, IN_MAX = 24, A_MAX = 16;
constexpr int LD = CL + 1;            // shared-memory row stride of the N = 32 weight matrices
constexpr int LD_RA0 = 4 * CL + 1;    // ... of rel0|att0 (N = 128)
constexpr int OB = CL * ORW;          // floats of one object-row buffer  [32][4]
constexpr int PB = CL * PR;           // floats of one pair-row buffer    [32][12]

// offsets (floats) of the weight segments inside the staged shared-memory copy
struct TW {
    int in_dim;
    int act_w, act_b, enc_w, enc_b, self0_w, self0_b, self1_w, self1_b, ra0_w, ra0_b, rel1_w, rel1_b,
        att1_w, att1_b, rel2_w, rel2_b, att2_w, att2_b, aff0_w, aff0_b, aff1_w, aff1_b, aff2_w, aff2_b,
        out0_w, out0_b, out1_w, out1_b, rew00_w, rew00_b, rew02_w, rew02_b, rew10_w, rew10_b, rew12_w,
        rew12_b, rew14_w, rew14_b;
    int total;
};
struct Seg { int src, dst, K, N, ld; };
constexpr int MAX_SEG = 40;
struct StageTable { int count; Seg seg[MAX_SEG]; };

// compact per-team layout of the forward-only kernels (dead buffers are reused)
struct FwdLay {
    static constexpr int SIN = 0, S = SIN + IN_MAX * ORW, H = S + OB, SELFD = H + OB, D = SELFD + OB, F1 = D + OB,
                         F2 = F1 + OB, CAT = F2 + OB, O1 = CAT + 2 * OB, OUT = O1 + OB, RH0 = OUT + OB,
                         RH1 = RH0 + OB, SMALL = RH1 + OB, ACT = SMALL + 64, DIST = ACT + A_MAX,
                         RA0 = DIST + 16, R1 = RA0 + 4 * PB, A1 = R1 + PB, REL = RA0, ATT = RA0 + PB,
                         TOTAL = A1 + PB;
};
// backward layout: nothing is overwritten.  The first XREC floats are what the forward pass leaves
// behind for the backward pass (layer inputs, network output, relation / attention values): in
// training the forward kernel stores them per (sequence, step) and the backward chain reloads them
// instead of recomputing the step.  [XREC, REC) are the pre-activation gradients; together they
// are the record the weight-gradient kernel consumes.
struct BwdLay {
    static constexpr int SIN = 0, S = SIN + IN_MAX * ORW, H = S + OB, D = H + OB, F1 = D + OB, F2 = F1 + OB,
                         CAT = F2 + OB, O1 = CAT + 2 * OB, RH0 = O1 + OB, SMALL = RH0 + OB, ACT = SMALL + 64,
                         DIST = ACT + A_MAX, RA0 = DIST + 16, R1 = RA0 + 4 * PB, A1 = R1 + PB, OUT = A1 + PB,
                         REL = OUT + OB, ATT = REL + PB;
    static constexpr int XREC = ATT + 16;
    static constexpr int G_OUT = XREC, G_O1P = G_OUT + OB, G_F3 = G_O1P + OB, G_F2P = G_F3 + OB,
                         G_F1P = G_F2P + OB, G_D = G_F1P + OB, G_HP = G_D + OB, G_ENC = G_HP + OB,
                         G_RH1 = G_ENC + OB, G_RH0P = G_RH1 + OB, G_SMALL = G_RH0P + OB, G_REL = G_SMALL + 64,
                         G_ATT = G_REL + PB, G_R1P = G_ATT + 16, G_A1P = G_R1P + PB, G_RA0P = G_A1P + PB;
    static constexpr int REC = G_RA0P + 4 * PB, GREC = REC - XREC;
    static constexpr int SELFD = REC, RH1 = SELFD + OB, GUV = RH1 + OB, GS_A = GUV + 4 * CL * 8, GS_B = GS_A + OB,
                         GDIST = GS_B + OB, GZ = GDIST + 32, TOTAL = GZ + 64;
};
// SMALL: RSUM [32] | R2 [16] | R3 [8] | REW [1];  G_SMALL: g_r2pre [16] | g_r3pre [8] | g_rewpre [1] | pad | g_emb [12]
constexpr int SM_RSUM = 0, SM_R2 = 32, SM_R3 = 48, SM_REW = 56;
constexpr int GSM_R2 = 0, GSM_R3 = 16, GSM_REW = 24, GSM_EMB = 32;
static_assert(FwdLay::TOTAL % 4 == 0 && BwdLay::XREC % 4 == 0 && BwdLay::REC % 4 == 0 && BwdLay::TOTAL % 4 == 0, "alignment");

template <int NW>
__device__ __forceinline__ void team_sync(int bar) {
    if (NW == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(NW * 32) : "memory");
}

__device__ __forceinline__ void load9(const float* src, float* v) {
    const float4 a = reinterpret_cast<const float4*>(src)[0], b = reinterpret_cast<const float4*>(src)[1],
                 c = reinterpret_cast<const float4*>(src)[2];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; v[8] = c.x;
}
__device__ __forceinline__ void store9(float* dst, const float* v) {
    reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
    reinterpret_cast<float4*>(dst)[2] = make_float4(v[8], 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void store3(float* dst, float a, float b, float c) {
    *reinterpret_cast<float4*>(dst) = make_float4(a, b, c, 0.f);
}

// weights: flat [K][N] in global memory -> row stride ld in shared memory.  4-byte cp.async: no
// register staging and every copy of the CTA is in flight at once (this runs once per launch, on
// the critical path of a kernel that only lives for a few time steps).
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void stage_weights(const StageTable& st, const float* __restrict__ weights, float* Ws) {
    for (int s = 0; s < st.count; ++s) {
        const Seg sg = st.seg[s];
        const int total = sg.K * sg.N;
        if ((sg.N & (sg.N - 1)) == 0) {
            const int sh = 31 - __clz(sg.N);
            for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
                const int k = idx >> sh, n = idx & (sg.N - 1);
                cp_async4(Ws + sg.dst + k * sg.ld + n, weights + sg.src + idx);
            }
        } else {
            for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
                const int k = idx / sg.N, n = idx - k * sg.N;
                cp_async4(Ws + sg.dst + k * sg.ld + n, weights + sg.src + idx);
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
// completes stage_weights (copies overlap whatever the caller does in between)
__device__ __forceinline__ void stage_weights_wait() {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
}

// ---- forward layers -------------------------------------------------------------------------
// object rows: out[n][r] = act(b[n] + sum_k W[k][n] in[k][r]) (+ res[n][r]); n = lane, r < 3
template <int ACT>
__device__ __forceinline__ void obj_fwd(const float* __restrict__ Wm, const float* __restrict__ bias, int K,
                                        const float* in, float* out, const float* res, int nl, int lane) {
    float a0 = bias[lane], a1 = a0, a2 = a0;
    const float4* in4 = reinterpret_cast<const float4*>(in);
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
        const float w = Wm[k * LD + lane];
        const float4 x = in4[k];
        a0 = fmaf(x.x, w, a0);
        a1 = fmaf(x.y, w, a1);
        a2 = fmaf(x.z, w, a2);
    }
    a0 = apply_act(a0, ACT, nl); a1 = apply_act(a1, ACT, nl); a2 = apply_act(a2, ACT, nl);
    if (res) {
        const float4 r = reinterpret_cast<const float4*>(res)[lane];
        a0 += r.x; a1 += r.y; a2 += r.z;
    }
    store3(out + lane * ORW, a0, a1, a2);
}

// pair rows, 32 outputs: n = lane, 9 rows
template <int K, int ACT>
__device__ __forceinline__ void pair_fwd(const float* __restrict__ Wm, const float* __restrict__ bias,
                                         const float* in, float* out, const float* res, int nl, int lane) {
    float acc[P];
    const float bv = bias[lane];
#pragma unroll
    for (int r = 0; r < P; ++r) acc[r] = bv;
    const float4* in4 = reinterpret_cast<const float4*>(in);
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float w = Wm[k * LD + lane];
        const float4 x0 = in4[k * 3], x1 = in4[k * 3 + 1], x2 = in4[k * 3 + 2];
        acc[0] = fmaf(x0.x, w, acc[0]); acc[1] = fmaf(x0.y, w, acc[1]); acc[2] = fmaf(x0.z, w, acc[2]);
        acc[3] = fmaf(x0.w, w, acc[3]); acc[4] = fmaf(x1.x, w, acc[4]); acc[5] = fmaf(x1.y, w, acc[5]);
        acc[6] = fmaf(x1.z, w, acc[6]); acc[7] = fmaf(x1.w, w, acc[7]); acc[8] = fmaf(x2.x, w, acc[8]);
    }
    float rv[P];
    if (res) load9(res + lane * PR, rv);
#pragma unroll
    for (int r = 0; r < P; ++r) {
        acc[r] = apply_act(acc[r], ACT, nl);
        if (res) acc[r] += rv[r];
    }
    store9(out + lane * PR, acc);
}

// rel0|att0 (dynamics.py:186-197), factorised; this call covers features lane + 32 (e0 + e), e < NE
template <int NE>
__device__ __forceinline__ void ra0_fwd(const float* __restrict__ Wm, const float* __restrict__ bias, const float* S,
                                        float* RA0, float* dist_out, int e0, int nl, int lane) {
    float U[NE][3], V[NE][3];
#pragma unroll
    for (int e = 0; e < NE; ++e)
#pragma unroll
        for (int i = 0; i < 3; ++i) U[e][i] = V[e][i] = 0.f;
    const float4* S4 = reinterpret_cast<const float4*>(S);
#pragma unroll 2
    for (int k = 0; k < CL; ++k) {
        const float4 x = S4[k];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int n = lane + 32 * (e0 + e);
            const float wa = Wm[k * LD_RA0 + n], wb = Wm[(CL + k) * LD_RA0 + n];
            U[e][0] = fmaf(x.x, wa, U[e][0]); U[e][1] = fmaf(x.y, wa, U[e][1]); U[e][2] = fmaf(x.z, wa, U[e][2]);
            V[e][0] = fmaf(x.x, wb, V[e][0]); V[e][1] = fmaf(x.y, wb, V[e][1]); V[e][2] = fmaf(x.z, wb, V[e][2]);
        }
    }
    const float4 px = S4[0], py = S4[1];
    const float xs[3] = {px.x, px.y, px.z}, ys[3] = {py.x, py.y, py.z};
    float dist[P];
#pragma unroll
    for (int i = 0; i < O; ++i)
#pragma unroll
        for (int j = 0; j < O; ++j) {
            const float dx = xs[i] - xs[j], dy = ys[i] - ys[j];
            dist[i * O + j] = dx * dx + dy * dy;
        }
    if (dist_out) {
#pragma unroll
        for (int p = 0; p < P; ++p)
            if (lane == p) dist_out[p] = dist[p];
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const int n = lane + 32 * (e0 + e);
        const float b = bias[n], wd = Wm[2 * CL * LD_RA0 + n];
        float v[P];
#pragma unroll
        for (int i = 0; i < O; ++i)
#pragma unroll
            for (int j = 0; j < O; ++j)
                v[i * O + j] = apply_act(fmaf(wd, dist[i * O + j], b + U[e][i] + V[e][j]), ACT_NL, nl);
        store9(RA0 + n * PR, v);
    }
}

// one dynamics step of one sequence by a team of NW warps; `a` = the team's activation slice.
// Expects a[SIN] (state, appearance) and, if action conditioned, act_row (global).  Ends with a
// team barrier: afterwards a[OUT] (and a[SMALL + SM_REW]) are visible to the whole team.
template <class Lay, int NW>
__device__ __forceinline__ void forward_step(const stove_gnn_cfg& c, const TW& w, const float* __restrict__ W, float* a,
                                             const float* __restrict__ act_row, int lane, int part, int bar) {
    const int nl = c.nonlin;
    const bool p0 = (NW == 1) || part == 0, p1 = (NW == 1) || part == 1;
    if (c.action_dim > 0) {
        // action embedding (dynamics.py:238-244): emb[o*4+e] -> s_in[cl/2 + e] of object o
        if (p0) {
            if (lane < c.action_dim) a[Lay::ACT + lane] = __ldg(act_row + lane);
            if (lane < O * 4) {
                float acc = W[w.act_b + lane];
                for (int k = 0; k < c.action_dim; ++k)
                    acc = fmaf(__ldg(act_row + k), W[w.act_w + k * (O * 4) + lane], acc);
                a[Lay::SIN + (HALF + (lane & 3)) * ORW + (lane >> 2)] = acc;
            }
        }
        team_sync<NW>(bar);
    }
    if (p0) {
        // state encoder with raw pass-through of the first lim_enc dims (dynamics.py:250)
        float a0 = W[w.enc_b + lane], a1 = a0, a2 = a0;
        const float4* in4 = reinterpret_cast<const float4*>(a + Lay::SIN);
#pragma unroll 4
        for (int k = 0; k < w.in_dim; ++k) {
            const float wv = W[w.enc_w + k * LD + lane];
            const float4 x = in4[k];
            a0 = fmaf(x.x, wv, a0); a1 = fmaf(x.y, wv, a1); a2 = fmaf(x.z, wv, a2);
        }
        if (lane < c.lim_enc) {
            const float4 x = in4[lane];
            a0 = x.x; a1 = x.y; a2 = x.z;
        }
        store3(a + Lay::S + lane * ORW, a0, a1, a2);
    }
    team_sync<NW>(bar);
    // rel0|att0: rel features (0..63) on part 0, att features (64..127) on part 1
    if (NW == 1) {
        ra0_fwd<4>(W + w.ra0_w, W + w.ra0_b, a + Lay::S, a + Lay::RA0, a + Lay::DIST, 0, nl, lane);
    } else {
        ra0_fwd<2>(W + w.ra0_w, W + w.ra0_b, a + Lay::S, a + Lay::RA0, part == 0 ? a + Lay::DIST : nullptr,
                   2 * part, nl, lane);
    }
    team_sync<NW>(bar);
    if (p0) pair_fwd<2 * CL, ACT_NL>(W + w.rel1_w, W + w.rel1_b, a + Lay::RA0, a + Lay::R1, nullptr, nl, lane);
    if (p1) pair_fwd<2 * CL, ACT_NL>(W + w.att1_w, W + w.att1_b, a + Lay::RA0 + 2 * PB, a + Lay::A1, nullptr, nl, lane);
    team_sync<NW>(bar);
    // part 0: rel2 (residual);  part 1: att2 (32 -> 1, exp) and the self-dynamics branch
    if (p0) pair_fwd<CL, ACT_NONE>(W + w.rel2_w, W + w.rel2_b, a + Lay::R1, a + Lay::REL, a + Lay::R1, nl, lane);
    if (p1) {
        if (lane < P) {
            float acc = W[w.att2_b];
#pragma unroll 8
            for (int k = 0; k < CL; ++k) acc = fmaf(a[Lay::A1 + k * PR + lane], W[w.att2_w + k], acc);
            a[Lay::ATT + lane] = expf(acc);
        }
        obj_fwd<ACT_NL>(W + w.self0_w, W + w.self0_b, CL, a + Lay::S, a + Lay::H, nullptr, nl, lane);
        __syncwarp();
        obj_fwd<ACT_NONE>(W + w.self1_w, W + w.self1_b, CL, a + Lay::H, a + Lay::SELFD, a + Lay::H, nl, lane);
    }
    team_sync<NW>(bar);
    if (p0) {
        // d_i = self_i + sum_j rel_ij * mask_ij * att_ij  (dynamics.py:203-208; the zero mask stays a
        // multiplication so that inf * 0 = nan propagates exactly as in the reference)
        float rel[P], att[P];
        load9(a + Lay::REL + lane * PR, rel);
        load9(a + Lay::ATT, att);
        const float4 sd = reinterpret_cast<const float4*>(a + Lay::SELFD)[lane];
        const float sv[3] = {sd.x, sd.y, sd.z};
        float d[3];
#pragma unroll
        for (int i = 0; i < O; ++i) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < O; ++j) acc += rel[i * O + j] * (i == j ? 0.f : 1.f) * att[i * O + j];
            d[i] = sv[i] + acc;
        }
        store3(a + Lay::D + lane * ORW, d[0], d[1], d[2]);
    }
    team_sync<NW>(bar);
    if (p0) obj_fwd<ACT_TANH>(W + w.aff0_w, W + w.aff0_b, CL, a + Lay::D, a + Lay::F1, nullptr, nl, lane);
    if (p1 && c.reward) obj_fwd<ACT_RELU>(W + w.rew00_w, W + w.rew00_b, CL, a + Lay::D, a + Lay::RH0, nullptr, nl, lane);
    team_sync<NW>(bar);
    if (p0) obj_fwd<ACT_TANH>(W + w.aff1_w, W + w.aff1_b, CL, a + Lay::F1, a + Lay::F2, a + Lay::F1, nl, lane);
    if (p1) {
        // cat = [aff3, s]: second half
        reinterpret_cast<float4*>(a + Lay::CAT + OB)[lane] = reinterpret_cast<const float4*>(a + Lay::S)[lane];
        if (c.reward) {
            // reward head (dynamics.py:254-263): rew02, sum over objects, 32 -> 16 -> 8 -> 1
            obj_fwd<ACT_NONE>(W + w.rew02_w, W + w.rew02_b, CL, a + Lay::RH0, a + Lay::RH1, nullptr, nl, lane);
            __syncwarp();
            const float4 r = reinterpret_cast<const float4*>(a + Lay::RH1)[lane];
            a[Lay::SMALL + SM_RSUM + lane] = r.x + r.y + r.z;
            __syncwarp();
            if (lane < CL / 2) {
                float acc = W[w.rew10_b + lane];
                for (int k = 0; k < CL; ++k)
                    acc = fmaf(a[Lay::SMALL + SM_RSUM + k], W[w.rew10_w + k * (CL / 2 + 1) + lane], acc);
                a[Lay::SMALL + SM_R2 + lane] = fmaxf(acc, 0.f);
            }
            __syncwarp();
            if (lane < CL / 4) {
                float acc = W[w.rew12_b + lane];
                for (int k = 0; k < CL / 2; ++k)
                    acc = fmaf(a[Lay::SMALL + SM_R2 + k], W[w.rew12_w + k * (CL / 4 + 1) + lane], acc);
                a[Lay::SMALL + SM_R3 + lane] = fmaxf(acc, 0.f);
            }
            __syncwarp();
            if (lane == 0) {
                float acc = W[w.rew14_b];
                for (int k = 0; k < CL / 4; ++k) acc = fmaf(a[Lay::SMALL + SM_R3 + k], W[w.rew14_w + k], acc);
                a[Lay::SMALL + SM_REW] = sigmoidf_(acc);
            }
        }
    }
    team_sync<NW>(bar);
    if (p0) obj_fwd<ACT_NONE>(W + w.aff2_w, W + w.aff2_b, CL, a + Lay::F2, a + Lay::CAT, nullptr, nl, lane);
    team_sync<NW>(bar);
    if (p0) obj_fwd<ACT_TANH>(W + w.out0_w, W + w.out0_b, 2 * CL, a + Lay::CAT, a + Lay::O1, nullptr, nl, lane);
    team_sync<NW>(bar);
    if (p0) obj_fwd<ACT_NONE>(W + w.out1_w, W + w.out1_b, CL, a + Lay::O1, a + Lay::OUT, a + Lay::O1, nl, lane);
    team_sync<NW>(bar);
}

// ---- constrain_z_dyn + Gaussian fusion with the SuPAIR state (one (object, feature) item) ----
struct FuseVal {
    float zd, sd, zdyn, mean, std, m_sup, s_sup, scale;
};
template <class Lay>
__device__ __forceinline__ FuseVal fuse_forward(const FuseCfg& f, const float* a, int o, int j,
                                                const float* __restrict__ sup6, const float* __restrict__ sstd6) {
    FuseVal v;
    if (j < 2) {
        v.zd = v.sd = v.zdyn = 0.f; v.scale = 1.f; v.m_sup = v.s_sup = 0.f;
        v.mean = __ldg(sup6 + j);
        v.std = __ldg(sstd6 + j);
        return v;
    }
    const int i = j - 2;
    v.scale = f.scale[i < 2 ? 0 : (i < 4 ? 1 : 2)];
    v.zd = 2.f * sigmoidf_(a[Lay::OUT + i * ORW + o]) - 1.f;
    v.sd = v.scale * sigmoidf_(a[Lay::OUT + (HALF + i) * ORW + o]);
    v.zdyn = v.zd + (i < 2 ? a[Lay::SIN + i * ORW + o] : 0.f);
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

struct LoopIO {
    int T, skip;
    const float *z_init, *sup, *sup_std, *eps, *actions, *app;
    float *z, *z_dyn, *z_dyn_std, *z_std, *logq, *trans, *reward;
    const float *g_z, *g_logq, *g_trans, *g_reward;
    float *g_z_init, *g_sup, *g_sup_std;
    float* xrec;        // [n][S][XREC]: forward activations kept for the backward pass (may be NULL)
};

template <class Lay>
__device__ __forceinline__ void load_state(float* a, const float* __restrict__ zsrc, int lane) {
    for (int e = lane; e < O * HALF; e += 32) {
        const int o = e / HALF, k = e - o * HALF;
        a[Lay::SIN + k * ORW + o] = __ldg(zsrc + o * ZD + 2 + k);
    }
}
template <class Lay>
__device__ __forceinline__ void load_app(const stove_gnn_cfg& c, float* a, const float* __restrict__ asrc, int lane) {
    const int a0 = HALF + (c.action_dim > 0 ? 4 : 0);
    for (int e = lane; e < O * c.app_dim; e += 32) {
        const int o = e / c.app_dim, k = e - o * c.app_dim;
        a[Lay::SIN + (a0 + k) * ORW + o] = __ldg(asrc + o * c.app_dim + k);
    }
}

// ---------------------------------------------------------------------------------------------
// forward: the whole dynamics loop, one team per sequence
// ---------------------------------------------------------------------------------------------
template <bool B, class X, class Y> struct pick_ { using type = X; };
template <class X, class Y> struct pick_<false, X, Y> { using type = Y; };

template <int NW, bool SAVE>
__global__ void __launch_bounds__(448, 1) dynloop_fwd_kernel(stove_gnn_cfg c, TW w, StageTable st, FuseCfg f,
                                                             int64_t n, LoopIO io, const float* __restrict__ weights) {
    using Lay = typename pick_<SAVE, BwdLay, FwdLay>::type;
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;
    stage_weights(st, weights, Ws);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int team = warp / NW, part = warp % NW, tpc = (blockDim.x >> 5) / NW, bar = 1 + team;
    float* a = smem + w.total + team * Lay::TOTAL;
    for (int i = lane + 32 * part; i < Lay::TOTAL; i += 32 * NW) a[i] = 0.f;
    stage_weights_wait();
    const int T = io.T, S = io.T - io.skip;
    for (int64_t sq = (int64_t)blockIdx.x * tpc + team; sq < n; sq += (int64_t)gridDim.x * tpc) {
        if (part == 0) load_state<Lay>(a, io.z_init + sq * O * ZD, lane);
        for (int k = 0; k < S; ++k) {
            const int t = io.skip + k;
            if (part == 0 && c.app_dim > 0) load_app<Lay>(c, a, io.app + ((sq * T + t - 1) * O) * c.app_dim, lane);
            team_sync<NW>(bar);
            forward_step<Lay, NW>(c, w, Ws, a, io.actions ? io.actions + (sq * T + t - 1) * c.action_dim : nullptr,
                                  lane, part, bar);
            if (SAVE) {
                // keep this step's activations for the backward pass (before the state is advanced)
                float4* dst = reinterpret_cast<float4*>(io.xrec + (sq * S + k) * (int64_t)BwdLay::XREC);
                const float4* src = reinterpret_cast<const float4*>(a);
                for (int i = lane + 32 * part; i < BwdLay::XREC / 4; i += 32 * NW) dst[i] = src[i];
                team_sync<NW>(bar);
            }
            if (part == 0) {
                // sample z_t ~ q, log q, transition likelihood; the sample is the next step's state
                float lq = 0.f, tr = 0.f;
                const int64_t zo = (sq * S + k) * O;
                for (int it = lane; it < O * ZD; it += 32) {
                    const int o = it / ZD, j = it - o * ZD;
                    const FuseVal v = fuse_forward<Lay>(f, a, o, j, io.sup + ((sq * T + t) * O + o) * 6,
                                                        io.sup_std + ((sq * T + t) * O + o) * 6);
                    const float e = __ldg(io.eps + (((int64_t)k * n + sq) * O + o) * ZD + j);
                    const float z = v.mean + v.std * e;
                    io.z[(zo + o) * ZD + j] = z;
                    if (io.z_std) io.z_std[(zo + o) * ZD + j] = v.std;
                    lq += -0.5f * e * e - logf(v.std) - HALF_LOG_2PI;
                    if (j >= 2) {
                        const int i = j - 2;
                        io.z_dyn[(zo + o) * HALF + i] = v.zdyn;
                        io.z_dyn_std[(zo + o) * HALF + i] = v.sd;
                        const float sd = f.trans_std[i], d = z - v.zdyn;
                        tr += -(d * d) / (2.f * sd * sd) - logf(sd) - HALF_LOG_2PI;
                        a[Lay::SIN + i * ORW + o] = z;
                    }
                }
                lq = warp_sum(lq);
                tr = warp_sum(tr);
                if (lane == 0) {
                    io.logq[sq * S + k] = lq;
                    io.trans[sq * S + k] = tr;
                    if (c.reward && io.reward) io.reward[sq * S + k] = a[Lay::SMALL + SM_REW];
                }
            }
        }
        team_sync<NW>(bar);
    }
}

// ---------------------------------------------------------------------------------------------
// rollout: `num` steps, state resident in shared memory (stove.py:823-846 + dynamics.py:147-179)
// ---------------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(448, 1) team_rollout_kernel(
    stove_gnn_cfg c, TW w, StageTable st, int64_t n, int num, const float* __restrict__ z_last,
    const float* __restrict__ actions, int action_len, const float* __restrict__ app,
    const float* __restrict__ weights, const float* __restrict__ noise, float pos_var, float vel_std,
    float latent_std, float* __restrict__ z_out, float* __restrict__ std_out, float* __restrict__ logq_out,
    float* __restrict__ rewards) {
    using Lay = FwdLay;
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;
    stage_weights(st, weights, Ws);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int team = warp / NW, part = warp % NW, tpc = (blockDim.x >> 5) / NW, bar = 1 + team;
    float* a = smem + w.total + team * Lay::TOTAL;
    for (int i = lane + 32 * part; i < Lay::TOTAL; i += 32 * NW) a[i] = 0.f;
    stage_weights_wait();
    for (int64_t sq = (int64_t)blockIdx.x * tpc + team; sq < n; sq += (int64_t)gridDim.x * tpc) {
        if (part == 0) {
            load_state<Lay>(a, z_last + sq * O * ZD, lane);
            if (c.app_dim > 0) load_app<Lay>(c, a, app + sq * O * c.app_dim, lane);
        }
        team_sync<NW>(bar);
        for (int t = 0; t < num; ++t) {
            const float* arow = actions ? actions + (sq * action_len + (t % action_len)) * c.action_dim : nullptr;
            forward_step<Lay, NW>(c, w, Ws, a, arow, lane, part, bar);
            if (part == 0) {
                // constrain, integrate positions, optional sampling; every item reads and writes only its
                // own state element, so the update is in place
                for (int e = lane; e < O * HALF; e += 32) {
                    const int o = e / HALF, k = e - o * HALF;
                    float m = 2.f * sigmoidf_(a[Lay::OUT + k * ORW + o]) - 1.f;
                    if (k < 2) m += a[Lay::SIN + k * ORW + o];
                    float val = m;
                    const int64_t o16 = ((sq * num + t) * O + o) * HALF + k;
                    if (noise || std_out) {
                        const float sd = (k < 2 ? pos_var : (k < 4 ? vel_std : latent_std)) *
                                         sigmoidf_(a[Lay::OUT + (HALF + k) * ORW + o]);
                        if (std_out) std_out[o16] = sd;
                        if (noise) {
                            const float e_ = __ldg(noise + o16);
                            val = m + sd * e_;
                            if (logq_out) logq_out[o16] = -0.5f * e_ * e_ - logf(sd) - HALF_LOG_2PI;
                        }
                    }
                    a[Lay::SIN + k * ORW + o] = val;
                    const int64_t oz = ((sq * num + t) * O + o) * ZD;
                    z_out[oz + 2 + k] = val;
                    if (k < 2) z_out[oz + k] = __ldg(z_last + (sq * O + o) * ZD + k);
                }
                if (c.reward && rewards && lane == 0) rewards[sq * num + t] = a[Lay::SMALL + SM_REW];
            }
            team_sync<NW>(bar);
        }
    }
}

// ---- backward helpers -----------------------------------------------------------------------
// r[i] += sum_n Wrow[n] g[n][i]   (object rows; Wrow = this lane's row of W)
__device__ __forceinline__ void obj_bwd_in(const float* __restrict__ Wrow, const float* g, float* r) {
    const float4* g4 = reinterpret_cast<const float4*>(g);
#pragma unroll 8
    for (int n = 0; n < CL; ++n) {
        const float wv = Wrow[n];
        const float4 gv = g4[n];
        r[0] = fmaf(wv, gv.x, r[0]); r[1] = fmaf(wv, gv.y, r[1]); r[2] = fmaf(wv, gv.z, r[2]);
    }
}
// acc[p] += sum_n Wrow[n] g[n][p]  (pair rows)
__device__ __forceinline__ void pair_bwd_in(const float* __restrict__ Wrow, const float* g, float* acc) {
    const float4* g4 = reinterpret_cast<const float4*>(g);
#pragma unroll 4
    for (int n = 0; n < CL; ++n) {
        const float wv = Wrow[n];
        const float4 g0 = g4[n * 3], g1 = g4[n * 3 + 1], g2 = g4[n * 3 + 2];
        acc[0] = fmaf(wv, g0.x, acc[0]); acc[1] = fmaf(wv, g0.y, acc[1]); acc[2] = fmaf(wv, g0.z, acc[2]);
        acc[3] = fmaf(wv, g0.w, acc[3]); acc[4] = fmaf(wv, g1.x, acc[4]); acc[5] = fmaf(wv, g1.y, acc[5]);
        acc[6] = fmaf(wv, g1.z, acc[6]); acc[7] = fmaf(wv, g1.w, acc[7]); acc[8] = fmaf(wv, g2.x, acc[8]);
    }
}
// two input rows at once (shares the gradient loads)
__device__ __forceinline__ void pair_bwd_in2(const float* __restrict__ Wr0, const float* __restrict__ Wr1,
                                             const float* g, float* acc0, float* acc1) {
    const float4* g4 = reinterpret_cast<const float4*>(g);
#pragma unroll 2
    for (int n = 0; n < CL; ++n) {
        const float w0 = Wr0[n], w1 = Wr1[n];
        const float4 g0 = g4[n * 3], g1 = g4[n * 3 + 1], g2 = g4[n * 3 + 2];
        const float gv[P] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x};
#pragma unroll
        for (int p = 0; p < P; ++p) {
            acc0[p] = fmaf(w0, gv[p], acc0[p]);
            acc1[p] = fmaf(w1, gv[p], acc1[p]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward chain: steps in reverse, state gradient carried on chip, one record per (sequence, step)
// ---------------------------------------------------------------------------------------------
template <int NW, bool SAVED>
__global__ void __launch_bounds__(192, 1) dynloop_bwd_kernel(stove_gnn_cfg c, TW w, StageTable st, FuseCfg f,
                                                             int64_t n, LoopIO io, const float* __restrict__ weights,
                                                             float* __restrict__ xrec, float* __restrict__ grec) {
    using Lay = BwdLay;
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;
    const float* W = Ws;
    stage_weights(st, weights, Ws);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int team = warp / NW, part = warp % NW, tpc = (blockDim.x >> 5) / NW, bar = 1 + team;
    const bool p0 = (NW == 1) || part == 0, p1 = (NW == 1) || part == 1;
    float* a = smem + w.total + team * Lay::TOTAL;
    for (int i = lane + 32 * part; i < Lay::TOTAL; i += 32 * NW) a[i] = 0.f;
    stage_weights_wait();
    const int T = io.T, S = io.T - io.skip, nl = c.nonlin;
    for (int64_t sq = (int64_t)blockIdx.x * tpc + team; sq < n; sq += (int64_t)gridDim.x * tpc) {
        if (p0)
            for (int e = lane; e < 64; e += 32) a[Lay::GZ + e] = 0.f;
        for (int k = S - 1; k >= 0; --k) {
            const int t = io.skip + k;
            if (SAVED) {
                // the forward kernel kept this step's activations: one asynchronous block copy
                const float4* src = reinterpret_cast<const float4*>(xrec + (sq * S + k) * (int64_t)Lay::XREC);
                float4* dst = reinterpret_cast<float4*>(a);
                for (int i = lane + 32 * part; i < Lay::XREC / 4; i += 32 * NW) cp_async16(dst + i, src + i);
                cp_async_commit();
                cp_async_wait<0>();
                team_sync<NW>(bar);
            } else {
                if (p0) {
                    load_state<Lay>(a, k == 0 ? io.z_init + sq * O * ZD : io.z + ((sq * S + k - 1) * O) * ZD, lane);
                    if (c.app_dim > 0) load_app<Lay>(c, a, io.app + ((sq * T + t - 1) * O) * c.app_dim, lane);
                }
                team_sync<NW>(bar);
                forward_step<Lay, NW>(c, w, W, a, io.actions ? io.actions + (sq * T + t - 1) * c.action_dim : nullptr,
                                      lane, part, bar);
            }
            // ---- prologue: d(sample, log q, transition lik) -> raw network output, SuPAIR inputs
            if (p0) {
                const float glq = io.g_logq ? __ldg(io.g_logq + sq * S + k) : 0.f;
                const float gtr = io.g_trans ? __ldg(io.g_trans + sq * S + k) : 0.f;
                for (int it = lane; it < O * ZD; it += 32) {
                    const int o = it / ZD, j = it - o * ZD;
                    const FuseVal v = fuse_forward<Lay>(f, a, o, j, io.sup + ((sq * T + t) * O + o) * 6,
                                                        io.sup_std + ((sq * T + t) * O + o) * 6);
                    const float e = __ldg(io.eps + (((int64_t)k * n + sq) * O + o) * ZD + j);
                    const float z = v.mean + v.std * e;
                    float gz = a[Lay::GZ + o * ZD + j];
                    if (io.g_z) gz += __ldg(io.g_z + ((sq * S + k) * O + o) * ZD + j);
                    float g_zdyn = 0.f;
                    if (j >= 2) {
                        const float sd = f.trans_std[j - 2];
                        const float tt = (z - v.zdyn) / (sd * sd) * gtr;
                        gz -= tt;
                        g_zdyn = tt;
                    }
                    const float g_mean = gz, g_std = gz * e - glq / v.std;
                    float* gsup = io.g_sup + ((sq * T + t) * O + o) * 6;
                    float* gsst = io.g_sup_std + ((sq * T + t) * O + o) * 6;
                    if (j < 2) {
                        gsup[j] = g_mean;
                        gsst[j] = g_std;
                        a[Lay::GZ + o * ZD + j] = 0.f;
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
                    a[Lay::G_OUT + i * ORW + o] = g_zdyn * (1.f - v.zd * v.zd) * 0.5f;
                    a[Lay::G_OUT + (HALF + i) * ORW + o] = g_sd * v.sd * (1.f - v.sd / v.scale);
                    // direct path pos_t = pos_{t-1} + delta; the network path is added at the end of the step
                    a[Lay::GZ + o * ZD + j] = (i < 2) ? g_zdyn : 0.f;
                }
            }
            team_sync<NW>(bar);
            float gs[3] = {0.f, 0.f, 0.f};      // d/ds accumulated by part 1 (out0 second half, self branch, att half of ra0)
            float gF2[3] = {0.f, 0.f, 0.f};
            // ---- B1: out1 (result = W o1 + b + o1, o1 = tanh(.));  reward head on part 1
            if (p0) {
                float r[3] = {0.f, 0.f, 0.f};
                obj_bwd_in(W + w.out1_w + lane * LD, a + Lay::G_OUT, r);
                const float4 go = reinterpret_cast<const float4*>(a + Lay::G_OUT)[lane];
                const float4 o1 = reinterpret_cast<const float4*>(a + Lay::O1)[lane];
                store3(a + Lay::G_O1P + lane * ORW, (r[0] + go.x) * (1.f - o1.x * o1.x),
                       (r[1] + go.y) * (1.f - o1.y * o1.y), (r[2] + go.z) * (1.f - o1.z * o1.z));
            }
            if (p1 && c.reward) {
                if (lane == 0) {
                    const float r = a[Lay::SMALL + SM_REW];
                    const float g = io.g_reward ? __ldg(io.g_reward + sq * S + k) : 0.f;
                    a[Lay::G_SMALL + GSM_REW] = g * r * (1.f - r);
                }
                __syncwarp();
                if (lane < CL / 4)
                    a[Lay::G_SMALL + GSM_R3 + lane] = W[w.rew14_w + lane] * a[Lay::G_SMALL + GSM_REW] *
                                                      (a[Lay::SMALL + SM_R3 + lane] > 0.f ? 1.f : 0.f);
                __syncwarp();
                if (lane < CL / 2) {
                    float acc = 0.f;
                    for (int q = 0; q < CL / 4; ++q)
                        acc = fmaf(W[w.rew12_w + lane * (CL / 4 + 1) + q], a[Lay::G_SMALL + GSM_R3 + q], acc);
                    a[Lay::G_SMALL + GSM_R2 + lane] = acc * (a[Lay::SMALL + SM_R2 + lane] > 0.f ? 1.f : 0.f);
                }
                __syncwarp();
                {
                    float acc = 0.f;
                    for (int q = 0; q < CL / 2; ++q)
                        acc = fmaf(W[w.rew10_w + lane * (CL / 2 + 1) + q], a[Lay::G_SMALL + GSM_R2 + q], acc);
                    store3(a + Lay::G_RH1 + lane * ORW, acc, acc, acc);     // d rsum -> every object
                }
                __syncwarp();
                float r[3] = {0.f, 0.f, 0.f};
                obj_bwd_in(W + w.rew02_w + lane * LD, a + Lay::G_RH1, r);
                const float4 h = reinterpret_cast<const float4*>(a + Lay::RH0)[lane];
                store3(a + Lay::G_RH0P + lane * ORW, h.x > 0.f ? r[0] : 0.f, h.y > 0.f ? r[1] : 0.f,
                       h.z > 0.f ? r[2] : 0.f);
            }
            team_sync<NW>(bar);
            // ---- B2: out0 (o1 = tanh(W [f3, s] + b)): first half -> g_f3, second half -> g_s
            if (p0) {
                float r[3] = {0.f, 0.f, 0.f};
                obj_bwd_in(W + w.out0_w + lane * LD, a + Lay::G_O1P, r);
                store3(a + Lay::G_F3 + lane * ORW, r[0], r[1], r[2]);
            }
            if (p1) obj_bwd_in(W + w.out0_w + (CL + lane) * LD, a + Lay::G_O1P, gs);
            team_sync<NW>(bar);
            // ---- B3: aff2 (f3 = W f2 + b), then the tanh of aff1 (f2 = tanh(W f1 + b) + f1)
            if (p0) {
                obj_bwd_in(W + w.aff2_w + lane * LD, a + Lay::G_F3, gF2);
                const float4 f2 = reinterpret_cast<const float4*>(a + Lay::F2)[lane];
                const float4 f1 = reinterpret_cast<const float4*>(a + Lay::F1)[lane];
                const float t0 = f2.x - f1.x, t1 = f2.y - f1.y, t2 = f2.z - f1.z;
                store3(a + Lay::G_F2P + lane * ORW, gF2[0] * (1.f - t0 * t0), gF2[1] * (1.f - t1 * t1),
                       gF2[2] * (1.f - t2 * t2));
            }
            team_sync<NW>(bar);
            // ---- B4: aff1 input gradient + residual, tanh of aff0
            if (p0) {
                float r[3] = {gF2[0], gF2[1], gF2[2]};
                obj_bwd_in(W + w.aff1_w + lane * LD, a + Lay::G_F2P, r);
                const float4 f1 = reinterpret_cast<const float4*>(a + Lay::F1)[lane];
                store3(a + Lay::G_F1P + lane * ORW, r[0] * (1.f - f1.x * f1.x), r[1] * (1.f - f1.y * f1.y),
                       r[2] * (1.f - f1.z * f1.z));
            }
            team_sync<NW>(bar);
            // ---- B5: aff0 (+ rew00) -> g_d
            if (p0) {
                float r[3] = {0.f, 0.f, 0.f};
                obj_bwd_in(W + w.aff0_w + lane * LD, a + Lay::G_F1P, r);
                if (c.reward) obj_bwd_in(W + w.rew00_w + lane * LD, a + Lay::G_RH0P, r);
                store3(a + Lay::G_D + lane * ORW, r[0], r[1], r[2]);
            }
            team_sync<NW>(bar);
            // ---- B6: aggregation d_i = self_i + sum_j mask rel_ij att_ij;  self1 on part 1
            if (p0) {
                float rel[P], att[P], tot[P];
                load9(a + Lay::REL + lane * PR, rel);
                load9(a + Lay::ATT, att);
                const float4 gd4 = reinterpret_cast<const float4*>(a + Lay::G_D)[lane];
                const float gd[3] = {gd4.x, gd4.y, gd4.z};
#pragma unroll
                for (int p = 0; p < P; ++p) tot[p] = warp_sum(gd[p / O] * rel[p]);
                float grel[P];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const float mask = (p / O == p % O) ? 0.f : 1.f;
                    grel[p] = gd[p / O] * mask * att[p];
                    // att = exp(lin): d att / d lin = att
                    if (lane == p) a[Lay::G_ATT + p] = tot[p] * mask * att[p];
                }
                store9(a + Lay::G_REL + lane * PR, grel);
            }
            if (p1) {
                float r[3] = {0.f, 0.f, 0.f};
                obj_bwd_in(W + w.self1_w + lane * LD, a + Lay::G_D, r);
                const float4 gd4 = reinterpret_cast<const float4*>(a + Lay::G_D)[lane];
                const float4 h = reinterpret_cast<const float4*>(a + Lay::H)[lane];
                store3(a + Lay::G_HP + lane * ORW, (r[0] + gd4.x) * act_grad(h.x, ACT_NL, nl),
                       (r[1] + gd4.y) * act_grad(h.y, ACT_NL, nl), (r[2] + gd4.z) * act_grad(h.z, ACT_NL, nl));
            }
            team_sync<NW>(bar);
            // ---- B7: rel2 (rel = W r1 + b + r1) on part 0;  att2 (32 -> 1, exp) and self0 on part 1
            if (p0) {
                float acc[P], y[P];
                load9(a + Lay::G_REL + lane * PR, acc);                 // residual path
                pair_bwd_in(W + w.rel2_w + lane * LD, a + Lay::G_REL, acc);
                load9(a + Lay::R1 + lane * PR, y);
#pragma unroll
                for (int p = 0; p < P; ++p) acc[p] *= act_grad(y[p], ACT_NL, nl);
                store9(a + Lay::G_R1P + lane * PR, acc);
            }
            if (p1) {
                float ga[P], y[P], v[P];
                load9(a + Lay::G_ATT, ga);
                load9(a + Lay::A1 + lane * PR, y);
                const float wv = W[w.att2_w + lane];
#pragma unroll
                for (int p = 0; p < P; ++p) v[p] = wv * ga[p] * act_grad(y[p], ACT_NL, nl);
                store9(a + Lay::G_A1P + lane * PR, v);
                obj_bwd_in(W + w.self0_w + lane * LD, a + Lay::G_HP, gs);
            }
            team_sync<NW>(bar);
            // ---- B8: rel1 / att1 (64 -> 32): input rows lane and lane + 32 of each half
            if (p0) {
                float acc0[P], acc1[P], y[P];
#pragma unroll
                for (int p = 0; p < P; ++p) acc0[p] = acc1[p] = 0.f;
                pair_bwd_in2(W + w.rel1_w + lane * LD, W + w.rel1_w + (lane + 32) * LD, a + Lay::G_R1P, acc0, acc1);
                load9(a + Lay::RA0 + lane * PR, y);
#pragma unroll
                for (int p = 0; p < P; ++p) acc0[p] *= act_grad(y[p], ACT_NL, nl);
                store9(a + Lay::G_RA0P + lane * PR, acc0);
                load9(a + Lay::RA0 + (lane + 32) * PR, y);
#pragma unroll
                for (int p = 0; p < P; ++p) acc1[p] *= act_grad(y[p], ACT_NL, nl);
                store9(a + Lay::G_RA0P + (lane + 32) * PR, acc1);
            }
            if (p1) {
                float acc0[P], acc1[P], y[P];
#pragma unroll
                for (int p = 0; p < P; ++p) acc0[p] = acc1[p] = 0.f;
                pair_bwd_in2(W + w.att1_w + lane * LD, W + w.att1_w + (lane + 32) * LD, a + Lay::G_A1P, acc0, acc1);
                load9(a + Lay::RA0 + (64 + lane) * PR, y);
#pragma unroll
                for (int p = 0; p < P; ++p) acc0[p] *= act_grad(y[p], ACT_NL, nl);
                store9(a + Lay::G_RA0P + (64 + lane) * PR, acc0);
                load9(a + Lay::RA0 + (96 + lane) * PR, y);
#pragma unroll
                for (int p = 0; p < P; ++p) acc1[p] *= act_grad(y[p], ACT_NL, nl);
                store9(a + Lay::G_RA0P + (96 + lane) * PR, acc1);
            }
            team_sync<NW>(bar);
            // ---- B9a: factorised rel0|att0: GU[n][i] = sum_j g[n][ij], GV[n][j] = sum_i g[n][ij], d dist
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!((NW == 1) || part == h)) continue;
                float gdist[P];
#pragma unroll
                for (int p = 0; p < P; ++p) gdist[p] = 0.f;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int nn = 64 * h + 32 * q + lane;
                    float g[P];
                    load9(a + Lay::G_RA0P + nn * PR, g);
                    reinterpret_cast<float4*>(a + Lay::GUV + nn * 8)[0] =
                        make_float4(g[0] + g[1] + g[2], g[3] + g[4] + g[5], g[6] + g[7] + g[8], 0.f);
                    reinterpret_cast<float4*>(a + Lay::GUV + nn * 8)[1] =
                        make_float4(g[0] + g[3] + g[6], g[1] + g[4] + g[7], g[2] + g[5] + g[8], 0.f);
                    const float wd = W[w.ra0_w + 2 * CL * LD_RA0 + nn];
#pragma unroll
                    for (int p = 0; p < P; ++p) gdist[p] = fmaf(wd, g[p], gdist[p]);
                }
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const float tot = warp_sum(gdist[p]);
                    if (lane == p) a[Lay::GDIST + 16 * h + p] = tot;
                }
            }
            team_sync<NW>(bar);
            // ---- B9b: g_s[k][i] += sum_n Wa[k][n] GU[n][i] + Wb[k][n] GV[n][i]
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!((NW == 1) || part == h)) continue;
                float r[3] = {0.f, 0.f, 0.f};
                const float* wa = W + w.ra0_w + lane * LD_RA0 + 64 * h;
                const float* wb = W + w.ra0_w + (CL + lane) * LD_RA0 + 64 * h;
                const float4* guv = reinterpret_cast<const float4*>(a + Lay::GUV + 64 * h * 8);
#pragma unroll 4
                for (int q = 0; q < 64; ++q) {
                    const float va = wa[q], vb = wb[q];
                    const float4 u = guv[2 * q], v = guv[2 * q + 1];
                    r[0] = fmaf(va, u.x, r[0]); r[1] = fmaf(va, u.y, r[1]); r[2] = fmaf(va, u.z, r[2]);
                    r[0] = fmaf(vb, v.x, r[0]); r[1] = fmaf(vb, v.y, r[1]); r[2] = fmaf(vb, v.z, r[2]);
                }
                if (h == 0) store3(a + Lay::GS_A + lane * ORW, r[0], r[1], r[2]);
                else store3(a + Lay::GS_B + lane * ORW, r[0] + gs[0], r[1] + gs[1], r[2] + gs[2]);
            }
            team_sync<NW>(bar);
            // ---- B10: total d/ds, distance term, encoder (s = [s_in[:lim], enc(s_in)[lim:]])
            float gS[3] = {0.f, 0.f, 0.f};
            if (p0) {
                const float4 ga = reinterpret_cast<const float4*>(a + Lay::GS_A)[lane];
                const float4 gb = reinterpret_cast<const float4*>(a + Lay::GS_B)[lane];
                gS[0] = ga.x + gb.x; gS[1] = ga.y + gb.y; gS[2] = ga.z + gb.z;
                if (lane < 2) {
                    // dist_ij = (x_i - x_j)^2 + (y_i - y_j)^2 ; lane 0 = x, lane 1 = y
                    const float4 xv = reinterpret_cast<const float4*>(a + Lay::S)[lane];
                    const float x[3] = {xv.x, xv.y, xv.z};
                    float gd[P];
#pragma unroll
                    for (int p = 0; p < P; ++p) gd[p] = a[Lay::GDIST + p] + a[Lay::GDIST + 16 + p];
#pragma unroll
                    for (int i = 0; i < O; ++i)
#pragma unroll
                        for (int j = 0; j < O; ++j)
                            gS[i] += 2.f * (x[i] - x[j]) * (gd[i * O + j] + gd[j * O + i]);
                }
                const bool raw = lane < c.lim_enc;
                store3(a + Lay::G_ENC + lane * ORW, raw ? 0.f : gS[0], raw ? 0.f : gS[1], raw ? 0.f : gS[2]);
            }
            team_sync<NW>(bar);
            if (p0) {
                if (lane < w.in_dim) {
                    float r[3] = {0.f, 0.f, 0.f};
                    obj_bwd_in(W + w.enc_w + lane * LD, a + Lay::G_ENC, r);
                    if (lane < c.lim_enc) { r[0] += gS[0]; r[1] += gS[1]; r[2] += gS[2]; }
                    if (lane < HALF) {
#pragma unroll
                        for (int o = 0; o < O; ++o) a[Lay::GZ + o * ZD + 2 + lane] += r[o];
                    } else if (c.action_dim > 0 && lane < HALF + 4) {
#pragma unroll
                        for (int o = 0; o < O; ++o) a[Lay::G_SMALL + GSM_EMB + o * 4 + (lane - HALF)] = r[o];
                    }
                }
            }
            team_sync<NW>(bar);
            // ---- the record of this (sequence, step) for the weight-gradient kernel: the gradient half
            // (and the activation half if the forward pass did not keep it)
            {
                float4* dst = reinterpret_cast<float4*>(grec + ((sq * S + k) * (int64_t)Lay::GREC));
                const float4* src = reinterpret_cast<const float4*>(a + Lay::XREC);
                for (int i = lane + 32 * part; i < Lay::GREC / 4; i += 32 * NW) dst[i] = src[i];
                if (!SAVED) {
                    float4* xd = reinterpret_cast<float4*>(xrec + ((sq * S + k) * (int64_t)Lay::XREC));
                    const float4* xsrc = reinterpret_cast<const float4*>(a);
                    for (int i = lane + 32 * part; i < Lay::XREC / 4; i += 32 * NW) xd[i] = xsrc[i];
                }
            }
            team_sync<NW>(bar);
        }
        if (p0) {
            for (int e = lane; e < O * ZD; e += 32) io.g_z_init[sq * O * ZD + e] = a[Lay::GZ + e];
            // the initial state is [sup[:, skip-1], noise latents] (stove.py:672-676): its first six gradient
            // components belong to g_sup[:, skip-1] (zeroed by the host before the launch)
            for (int e = lane; e < O * 6; e += 32) {
                const int o = e / 6, j = e - o * 6;
                io.g_sup[((sq * T + (io.skip - 1)) * O + o) * 6 + j] = a[Lay::GZ + o * ZD + j];
            }
            // time steps before `skip` receive no other gradient through the loop (no host memset: the kernel
            // owns every element of g_sup / g_sup_std)
            for (int e = lane; e < (io.skip - 1) * O * 6; e += 32) io.g_sup[sq * T * O * 6 + e] = 0.f;
            for (int e = lane; e < io.skip * O * 6; e += 32) io.g_sup_std[sq * T * O * 6 + e] = 0.f;
        }
        team_sync<NW>(bar);
    }
}

// ---------------------------------------------------------------------------------------------
// weight gradients from the records: every thread owns a fixed set of weight-gradient entries in
// registers and streams over this CTA's records (double-buffered cp.async); one slab per CTA.
// ---------------------------------------------------------------------------------------------
// acc[q] += sum_{r<3} x[w + 16 q][r] g[lane][r]
template <int NQ>
__device__ __forceinline__ void wg_obj(const float* x, const float* g, int wp, int lane, float* acc) {
    const float4 gv = reinterpret_cast<const float4*>(g)[lane];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const float4 xv = reinterpret_cast<const float4*>(x)[wp + 16 * q];
        acc[q] = fmaf(xv.x, gv.x, acc[q]);
        acc[q] = fmaf(xv.y, gv.y, acc[q]);
        acc[q] = fmaf(xv.z, gv.z, acc[q]);
    }
}
template <int NQ>
__device__ __forceinline__ void wg_pair(const float* x, const float* g, int wp, int lane, float* acc) {
    float gv[P];
    load9(g + lane * PR, gv);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float xv[P];
        load9(x + (wp + 16 * q) * PR, xv);
#pragma unroll
        for (int p = 0; p < P; ++p) acc[q] = fmaf(xv[p], gv[p], acc[q]);
    }
}

__global__ void __launch_bounds__(512, 1) dynloop_wgrad_kernel(stove_gnn_cfg c, GnnLayout L, int64_t nrec,
                                                               const float* __restrict__ xrec,
                                                               const float* __restrict__ grec,
                                                               float* __restrict__ slabs) {
    using Lay = BwdLay;
    constexpr int REC = Lay::REC;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    float a_enc[2] = {0.f, 0.f}, a_self0[2] = {0.f, 0.f}, a_self1[2] = {0.f, 0.f}, a_aff0[2] = {0.f, 0.f},
          a_aff1[2] = {0.f, 0.f}, a_aff2[2] = {0.f, 0.f}, a_out1[2] = {0.f, 0.f}, a_rel2[2] = {0.f, 0.f},
          a_rew00[2] = {0.f, 0.f}, a_rew02[2] = {0.f, 0.f};
    float a_out0[4] = {0.f, 0.f, 0.f, 0.f}, a_rel1[4] = {0.f, 0.f, 0.f, 0.f}, a_att1[4] = {0.f, 0.f, 0.f, 0.f};
    float a_ra[4][2][2];
    float a_wd[4] = {0.f, 0.f, 0.f, 0.f};          // warp 0: w_d row of rel0|att0;  warp 1: its bias
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int q = 0; q < 2; ++q) a_ra[e][q][0] = a_ra[e][q][1] = 0.f;
    float a_att2 = 0.f, a_bias = 0.f, a_act = 0.f, a_r10 = 0.f, a_r12 = 0.f, a_r14 = 0.f, a_rb = 0.f;
    // bias rows: warp -> gradient buffer of the layer whose bias it sums
    const int bias3[10] = {Lay::G_ENC, Lay::G_HP, Lay::G_D, Lay::G_F1P, Lay::G_F2P, Lay::G_F3, Lay::G_O1P,
                           Lay::G_OUT, Lay::G_RH0P, Lay::G_RH1};
    int nmine = 0;
    for (int64_t r = blockIdx.x; r < nrec; r += gridDim.x) ++nmine;
    auto issue = [&](int i) {
        const int64_t r = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
        const float* xsrc = xrec + r * Lay::XREC;
        const float* gsrc = grec + r * Lay::GREC;
        float* dst = smem + (i & 1) * REC;
        for (int q = tid; q < Lay::XREC / 4; q += blockDim.x) cp_async16(dst + 4 * q, xsrc + 4 * q);
        for (int q = tid; q < Lay::GREC / 4; q += blockDim.x) cp_async16(dst + Lay::XREC + 4 * q, gsrc + 4 * q);
        cp_async_commit();
    };
    if (nmine > 0) issue(0);
    for (int i = 0; i < nmine; ++i) {
        if (i + 1 < nmine) {
            issue(i + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* b = smem + (i & 1) * REC;
        wg_obj<2>(b + Lay::SIN, b + Lay::G_ENC, wp, lane, a_enc);       // rows >= in_dim are zero and unused
        wg_obj<2>(b + Lay::S, b + Lay::G_HP, wp, lane, a_self0);
        wg_obj<2>(b + Lay::H, b + Lay::G_D, wp, lane, a_self1);
        wg_obj<2>(b + Lay::D, b + Lay::G_F1P, wp, lane, a_aff0);
        wg_obj<2>(b + Lay::F1, b + Lay::G_F2P, wp, lane, a_aff1);
        wg_obj<2>(b + Lay::F2, b + Lay::G_F3, wp, lane, a_aff2);
        wg_obj<4>(b + Lay::CAT, b + Lay::G_O1P, wp, lane, a_out0);
        wg_obj<2>(b + Lay::O1, b + Lay::G_OUT, wp, lane, a_out1);
        if (c.reward) {
            wg_obj<2>(b + Lay::D, b + Lay::G_RH0P, wp, lane, a_rew00);
            wg_obj<2>(b + Lay::RH0, b + Lay::G_RH1, wp, lane, a_rew02);
        }
        wg_pair<2>(b + Lay::R1, b + Lay::G_REL, wp, lane, a_rel2);
        wg_pair<4>(b + Lay::RA0, b + Lay::G_R1P, wp, lane, a_rel1);
        wg_pair<4>(b + Lay::RA0 + 2 * PB, b + Lay::G_A1P, wp, lane, a_att1);
        {   // rel0|att0, factorised
            float dist[P];
            load9(b + Lay::DIST, dist);
            const float4 x0 = reinterpret_cast<const float4*>(b + Lay::S)[wp];
            const float4 x1 = reinterpret_cast<const float4*>(b + Lay::S)[wp + 16];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float g[P];
                load9(b + Lay::G_RA0P + (lane + 32 * e) * PR, g);
                const float u0 = g[0] + g[1] + g[2], u1 = g[3] + g[4] + g[5], u2 = g[6] + g[7] + g[8];
                const float v0 = g[0] + g[3] + g[6], v1 = g[1] + g[4] + g[7], v2 = g[2] + g[5] + g[8];
                a_ra[e][0][0] += x0.x * u0 + x0.y * u1 + x0.z * u2;
                a_ra[e][0][1] += x0.x * v0 + x0.y * v1 + x0.z * v2;
                a_ra[e][1][0] += x1.x * u0 + x1.y * u1 + x1.z * u2;
                a_ra[e][1][1] += x1.x * v0 + x1.y * v1 + x1.z * v2;
                if (wp == 0) {
#pragma unroll
                    for (int p = 0; p < P; ++p) a_wd[e] = fmaf(dist[p], g[p], a_wd[e]);
                } else if (wp == 1) {
                    a_wd[e] += u0 + u1 + u2;
                }
            }
        }
        if (wp == 2) {                       // att2 weight (32 x 1)
            float y[P], ga[P];
            load9(b + Lay::A1 + lane * PR, y);
            load9(b + Lay::G_ATT, ga);
#pragma unroll
            for (int p = 0; p < P; ++p) a_att2 = fmaf(y[p], ga[p], a_att2);
        } else if (wp == 3) {                // rel2 bias; lane 0 also att2 bias
            float g[P];
            load9(b + Lay::G_REL + lane * PR, g);
#pragma unroll
            for (int p = 0; p < P; ++p) a_bias += g[p];
            if (lane == 0) {
                load9(b + Lay::G_ATT, g);
#pragma unroll
                for (int p = 0; p < P; ++p) a_att2 += g[p];
            }
        } else if (wp >= 4 && wp < 14) {
            const float4 g = reinterpret_cast<const float4*>(b + bias3[wp - 4])[lane];
            a_bias += g.x + g.y + g.z;
        } else if (wp >= 14) {               // rel1 / att1 bias
            float g[P];
            load9(b + (wp == 14 ? Lay::G_R1P : Lay::G_A1P) + lane * PR, g);
#pragma unroll
            for (int p = 0; p < P; ++p) a_bias += g[p];
        }
        if (c.action_dim > 0) {
            const int na = c.action_dim * (O * 4);
            if (tid < na) a_act = fmaf(b[Lay::ACT + tid / (O * 4)], b[Lay::G_SMALL + GSM_EMB + tid % (O * 4)], a_act);
            else if (tid >= 256 && tid < 256 + O * 4) a_act += b[Lay::G_SMALL + GSM_EMB + tid - 256];
        }
        if (c.reward) {
            a_r10 = fmaf(b[Lay::SMALL + SM_RSUM + (tid >> 4)], b[Lay::G_SMALL + GSM_R2 + (tid & 15)], a_r10);
            if (tid < 128) a_r12 = fmaf(b[Lay::SMALL + SM_R2 + (tid >> 3)], b[Lay::G_SMALL + GSM_R3 + (tid & 7)], a_r12);
            if (tid < 8) a_r14 = fmaf(b[Lay::SMALL + SM_R3 + tid], b[Lay::G_SMALL + GSM_REW], a_r14);
            // biases: threads 256.. (rew10: 16, rew12: 8, rew14: 1)
            if (tid >= 256 && tid < 272) a_rb += b[Lay::G_SMALL + GSM_R2 + tid - 256];
            else if (tid >= 288 && tid < 296) a_rb += b[Lay::G_SMALL + GSM_R3 + tid - 288];
            else if (tid == 320) a_rb += b[Lay::G_SMALL + GSM_REW];
        }
        __syncthreads();
    }
    // ---- write this CTA's slab (layout of the flat weight buffer, gnn_layout)
    float* sl = slabs + (int64_t)blockIdx.x * L.total;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int k = wp + 16 * q;
        if (k < L.in_dim) sl[L.enc_w + k * CL + lane] = a_enc[q];
        sl[L.self0_w + k * CL + lane] = a_self0[q];
        sl[L.self1_w + k * CL + lane] = a_self1[q];
        sl[L.aff0_w + k * CL + lane] = a_aff0[q];
        sl[L.aff1_w + k * CL + lane] = a_aff1[q];
        sl[L.aff2_w + k * CL + lane] = a_aff2[q];
        sl[L.out1_w + k * CL + lane] = a_out1[q];
        sl[L.rel2_w + k * CL + lane] = a_rel2[q];
        if (c.reward) {
            sl[L.rew00_w + k * CL + lane] = a_rew00[q];
            sl[L.rew02_w + k * CL + lane] = a_rew02[q];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            sl[L.ra0_w + k * (4 * CL) + lane + 32 * e] = a_ra[e][q][0];
            sl[L.ra0_w + (CL + k) * (4 * CL) + lane + 32 * e] = a_ra[e][q][1];
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = wp + 16 * q;
        sl[L.out0_w + k * CL + lane] = a_out0[q];
        sl[L.rel1_w + k * CL + lane] = a_rel1[q];
        sl[L.att1_w + k * CL + lane] = a_att1[q];
    }
    if (wp == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) sl[L.ra0_w + 2 * CL * (4 * CL) + lane + 32 * e] = a_wd[e];
    } else if (wp == 1) {
#pragma unroll
        for (int e = 0; e < 4; ++e) sl[L.ra0_b + lane + 32 * e] = a_wd[e];
    } else if (wp == 2) {
        sl[L.att2_w + lane] = a_att2;
    } else if (wp == 3) {
        sl[L.rel2_b + lane] = a_bias;
        if (lane == 0) sl[L.att2_b] = a_att2;
    } else if (wp < 14) {
        const int boff[10] = {L.enc_b, L.self0_b, L.self1_b, L.aff0_b, L.aff1_b, L.aff2_b, L.out0_b, L.out1_b,
                              L.rew00_b, L.rew02_b};
        if (boff[wp - 4] >= 0) sl[boff[wp - 4] + lane] = a_bias;
    } else {
        sl[(wp == 14 ? L.rel1_b : L.att1_b) + lane] = a_bias;
    }
    if (c.action_dim > 0) {
        const int na = c.action_dim * (O * 4);
        if (tid < na) sl[L.act_w + tid] = a_act;
        else if (tid >= 256 && tid < 256 + O * 4) sl[L.act_b + tid - 256] = a_act;
    }
    if (c.reward) {
        sl[L.rew10_w + tid] = a_r10;
        if (tid < 128) sl[L.rew12_w + tid] = a_r12;
        if (tid < 8) sl[L.rew14_w + tid] = a_r14;
        if (tid >= 256 && tid < 272) sl[L.rew10_b + tid - 256] = a_rb;
        else if (tid >= 288 && tid < 296) sl[L.rew12_b + tid - 288] = a_rb;
        else if (tid == 320) sl[L.rew14_b] = a_rb;
    }
}

// g_w[i] = sum_s slabs[s][i]
__global__ void dynloop_reduce_kernel(const float* __restrict__ slabs, int nslab, int total, float* __restrict__ g_w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float acc = 0.f;
    for (int s = 0; s < nslab; ++s) acc += slabs[(int64_t)s * total + i];
    g_w[i] = acc;
}

// ---- host -----------------------------------------------------------------------------------
static void build_tables(const stove_gnn_cfg* c, const GnnLayout& L, TW* tw, StageTable* st) {
    int at = 0, cnt = 0;
    auto seg = [&](int& wo, int& bo, int src_w, int src_b, int K, int N, int ld) {
        if (src_w < 0) { wo = bo = -1; return; }
        wo = at; st->seg[cnt++] = Seg{src_w, at, K, N, ld}; at += pad4(K * ld);
        bo = at; st->seg[cnt++] = Seg{src_b, at, 1, N, N}; at += pad4(N);
    };
    tw->in_dim = L.in_dim;
    seg(tw->act_w, tw->act_b, L.act_w, L.act_b, c->action_dim, O * 4, O * 4);
    seg(tw->enc_w, tw->enc_b, L.enc_w, L.enc_b, L.in_dim, CL, LD);
    seg(tw->self0_w, tw->self0_b, L.self0_w, L.self0_b, CL, CL, LD);
    seg(tw->self1_w, tw->self1_b, L.self1_w, L.self1_b, CL, CL, LD);
    seg(tw->ra0_w, tw->ra0_b, L.ra0_w, L.ra0_b, 2 * CL + 1, 4 * CL, LD_RA0);
    seg(tw->rel1_w, tw->rel1_b, L.rel1_w, L.rel1_b, 2 * CL, CL, LD);
    seg(tw->att1_w, tw->att1_b, L.att1_w, L.att1_b, 2 * CL, CL, LD);
    seg(tw->rel2_w, tw->rel2_b, L.rel2_w, L.rel2_b, CL, CL, LD);
    seg(tw->att2_w, tw->att2_b, L.att2_w, L.att2_b, CL, 1, 1);
    seg(tw->aff0_w, tw->aff0_b, L.aff0_w, L.aff0_b, CL, CL, LD);
    seg(tw->aff1_w, tw->aff1_b, L.aff1_w, L.aff1_b, CL, CL, LD);
    seg(tw->aff2_w, tw->aff2_b, L.aff2_w, L.aff2_b, CL, CL, LD);
    seg(tw->out0_w, tw->out0_b, L.out0_w, L.out0_b, 2 * CL, CL, LD);
    seg(tw->out1_w, tw->out1_b, L.out1_w, L.out1_b, CL, CL, LD);
    seg(tw->rew00_w, tw->rew00_b, L.rew00_w, L.rew00_b, CL, CL, LD);
    seg(tw->rew02_w, tw->rew02_b, L.rew02_w, L.rew02_b, CL, CL, LD);
    seg(tw->rew10_w, tw->rew10_b, L.rew10_w, L.rew10_b, CL, CL / 2, CL / 2 + 1);
    seg(tw->rew12_w, tw->rew12_b, L.rew12_w, L.rew12_b, CL / 2, CL / 4, CL / 4 + 1);
    seg(tw->rew14_w, tw->rew14_b, L.rew14_w, L.rew14_b, CL / 4, 1, 1);
    tw->total = at;
    st->count = cnt;
}

static bool supported(const stove_gnn_cfg* c, const GnnLayout& L) {
    return c->num_obj == O && c->cl == CL && gnn_default_state(c) && L.in_dim <= IN_MAX && c->action_dim <= A_MAX &&
           !stove_opt(OPT_DYNLOOP_GENERIC);
}

// teams per CTA: spread the sequences over the 148 SMs first, then fill each SM
static int pick_tpc(int64_t n, int max_tpc) {
    int64_t t = (n + 147) / 148;
    if (t < 1) t = 1;
    if (t > max_tpc) t = max_tpc;
    return (int)t;
}
// g_sup[:, skip-1, :, :6] = g_z_init[..., :6] for the step-by-step path (the loop kernel writes it itself)
__global__ void init_grad_to_sup_kernel(int64_t n, int T, int skip, int O, int Z, const float* __restrict__ g_z_init,
                                        float* __restrict__ g_sup) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * O * 6) return;
    const int j = (int)(i % 6);
    const int64_t bo = i / 6, b = bo / O;
    const int o = (int)(bo - b * O);
    g_sup[((b * T + (skip - 1)) * O + o) * 6 + j] = g_z_init[bo * Z + j];
}
}  // namespace tk

// generic per-step kernels (gnn.cu)
extern "C" int stove_dynstep_fwd(const stove_gnn_cfg*, const stove_fuse_cfg*, int64_t, const stove_dynstep_io*,
                                 const float*, void*);
extern "C" int stove_dynstep_bwd(const stove_gnn_cfg*, const stove_fuse_cfg*, int64_t, const stove_dynstep_io*,
                                 const float*, float*, int, int, void*, void*);

static int dynloop_check(const stove_gnn_cfg* cfg, const stove_fuse_cfg* fuse, int64_t n, const stove_dynloop_io* io,
                         const float* weights) {
    int rc = gnn_check(cfg);
    if (rc) return rc;
    STOVE_CHECK_ARG(gnn_default_state(cfg), "state_dim != cl/2 is served by stove_gnn_fwd / stove_gnn_bwd only");
    STOVE_CHECK_ARG(fuse && io && weights && n >= 0, "null pointer");
    STOVE_CHECK_ARG(io->T > io->skip && io->skip >= 1, "need T > skip >= 1");
    STOVE_CHECK_ARG(io->z_init && io->sup && io->sup_std && io->eps && io->z, "null tensor in stove_dynloop_io");
    STOVE_CHECK_ARG((cfg->action_dim > 0) == (io->actions != nullptr), "actions do not match cfg.action_dim");
    STOVE_CHECK_ARG((cfg->app_dim > 0) == (io->app != nullptr), "appearances do not match cfg.app_dim");
    STOVE_CHECK_ARG(((uintptr_t)weights & 15) == 0, "weights must be 16-byte aligned");
    return STOVE_OK;
}

static tk::LoopIO to_loop_io(const stove_dynloop_io* io) {
    tk::LoopIO l;
    l.T = io->T; l.skip = io->skip;
    l.z_init = io->z_init; l.sup = io->sup; l.sup_std = io->sup_std; l.eps = io->eps;
    l.actions = io->actions; l.app = io->app;
    l.z = io->z; l.z_dyn = io->z_dyn; l.z_dyn_std = io->z_dyn_std; l.z_std = io->z_std;
    l.logq = io->logq; l.trans = io->trans; l.reward = io->reward;
    l.g_z = io->g_z; l.g_logq = io->g_logq; l.g_trans = io->g_trans; l.g_reward = io->g_reward;
    l.g_z_init = io->g_z_init; l.g_sup = io->g_sup; l.g_sup_std = io->g_sup_std;
    l.xrec = io->xrec;
    return l;
}

// time slice k of the generic per-step interface
static stove_dynstep_io step_io(const stove_gnn_cfg* cfg, const stove_dynloop_io* io, int64_t n, int k) {
    const int O = cfg->num_obj, Z = cfg->cl / 2 + 2, T = io->T, S = io->T - io->skip, t = io->skip + k;
    stove_dynstep_io s;
    memset(&s, 0, sizeof(s));
    if (k == 0) { s.z_prev = io->z_init; s.z_prev_ss = O * Z; }
    else { s.z_prev = io->z + (int64_t)(k - 1) * O * Z; s.z_prev_ss = (int64_t)S * O * Z; }
    s.sup = io->sup + (int64_t)t * O * 6; s.sup_std = io->sup_std + (int64_t)t * O * 6; s.sup_ss = (int64_t)T * O * 6;
    s.eps = io->eps + (int64_t)k * n * O * Z; s.eps_ss = O * Z;
    if (io->actions) { s.actions = io->actions + (int64_t)(t - 1) * cfg->action_dim; s.act_ss = (int64_t)T * cfg->action_dim; }
    if (io->app) { s.app = io->app + (int64_t)(t - 1) * O * cfg->app_dim; s.app_ss = (int64_t)T * O * cfg->app_dim; }
    return s;
}

extern "C" int stove_dynloop_fwd(const stove_gnn_cfg* cfg, const stove_fuse_cfg* fuse, int64_t n,
                                 const stove_dynloop_io* io, const float* weights, void* stream) {
    int rc = dynloop_check(cfg, fuse, n, io, weights);
    if (rc) return rc;
    STOVE_CHECK_ARG(io->z_dyn && io->z_dyn_std && io->logq && io->trans, "null output in stove_dynloop_io");
    if (n == 0) return STOVE_OK;
    GnnLayout L = gnn_layout(cfg);
    cudaStream_t st = (cudaStream_t)stream;
    const int S = io->T - io->skip;
    if (!tk::supported(cfg, L)) {
        // any other shape: the CTA-wide per-step kernels, chained here
        const int O = cfg->num_obj, Z = cfg->cl / 2 + 2;
        for (int k = 0; k < S; ++k) {
            stove_dynstep_io s = step_io(cfg, io, n, k);
            s.z_out = io->z + (int64_t)k * O * Z; s.z_out_ss = (int64_t)S * O * Z;
            s.z_dyn = io->z_dyn + (int64_t)k * O * (Z - 2); s.z_dyn_std = io->z_dyn_std + (int64_t)k * O * (Z - 2);
            s.zdyn_ss = (int64_t)S * O * (Z - 2);
            if (io->z_std) { s.z_std = io->z_std + (int64_t)k * O * Z; s.z_std_ss = (int64_t)S * O * Z; }
            s.logq = io->logq + k; s.trans = io->trans + k; s.sc_ss = S;
            if (io->reward) s.reward = io->reward + k;
            rc = stove_dynstep_fwd(cfg, fuse, n, &s, weights, stream);
            if (rc) return rc;
        }
        return STOVE_OK;
    }
    tk::TW tw;
    tk::StageTable tab;
    tk::build_tables(cfg, L, &tw, &tab);
    const int NW = stove_opt(OPT_DYNLOOP_NW);
    const bool save = io->xrec != nullptr;           // training: keep the activations for the backward pass
    const int lay_total = save ? tk::BwdLay::TOTAL : tk::FwdLay::TOTAL;
    const int max_tpc = (int)((kMaxSmem / sizeof(float) - tw.total) / lay_total);
    const int tpc = tk::pick_tpc(n, max_tpc < 7 ? max_tpc : 7);
    const size_t smem = sizeof(float) * ((size_t)tw.total + (size_t)tpc * lay_total);
    const int64_t groups = (n + tpc - 1) / tpc;
    const int ctas = (int)(groups < 148 ? groups : 148);
    const FuseCfg f = make_fuse(cfg, fuse);
    const tk::LoopIO lio = to_loop_io(io);
#define DYNLOOP_FWD_LAUNCH(NW_, SAVE_)                                                                          \
    do {                                                                                                        \
        STOVE_CUDA(cudaFuncSetAttribute(tk::dynloop_fwd_kernel<NW_, SAVE_>,                                     \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
        STOVE_KERNEL(K_DYNLOOP_FWD, st, tk::dynloop_fwd_kernel<NW_, SAVE_><<<ctas, 32 * NW_ * tpc, smem, st>>>( \
            *cfg, tw, tab, f, n, lio, weights));                                                                \
    } while (0)
    if (NW == 2) { if (save) DYNLOOP_FWD_LAUNCH(2, true); else DYNLOOP_FWD_LAUNCH(2, false); }
    else { if (save) DYNLOOP_FWD_LAUNCH(1, true); else DYNLOOP_FWD_LAUNCH(1, false); }
#undef DYNLOOP_FWD_LAUNCH
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

struct DynloopBwdPlan {
    bool fast;
    int tpc, ctas, wg_ctas;
    size_t smem, xrec_bytes, grec_bytes, slab_bytes, carry_bytes, gnn_ws;
};

extern "C" size_t stove_gnn_bwd_workspace(const stove_gnn_cfg* cfg, int64_t n);

static DynloopBwdPlan dynloop_bwd_plan(const stove_gnn_cfg* cfg, const GnnLayout& L, int64_t n, int S) {
    DynloopBwdPlan p;
    memset(&p, 0, sizeof(p));
    p.fast = tk::supported(cfg, L);
    if (!p.fast) {
        p.carry_bytes = sizeof(float) * 2 * (size_t)n * cfg->num_obj * (cfg->cl / 2 + 2);
        p.carry_bytes = (p.carry_bytes + 255) / 256 * 256;
        p.gnn_ws = stove_gnn_bwd_workspace(cfg, n);
        return p;
    }
    tk::TW tw;
    tk::StageTable tab;
    tk::build_tables(cfg, L, &tw, &tab);
    const int max_tpc = (int)((kMaxSmem / sizeof(float) - tw.total) / tk::BwdLay::TOTAL);
    p.tpc = tk::pick_tpc(n, max_tpc < 3 ? max_tpc : 3);
    p.smem = sizeof(float) * ((size_t)tw.total + (size_t)p.tpc * tk::BwdLay::TOTAL);
    const int64_t groups = (n + p.tpc - 1) / p.tpc;
    p.ctas = (int)(groups < 148 ? groups : 148);
    const int64_t nrec = n * S;
    const int wg_cap = stove_opt(OPT_WGRAD_CTAS) > 0 ? stove_opt(OPT_WGRAD_CTAS) : 148;
    p.wg_ctas = (int)(nrec < wg_cap ? nrec : wg_cap);
    p.xrec_bytes = sizeof(float) * (size_t)nrec * tk::BwdLay::XREC;
    p.grec_bytes = sizeof(float) * (size_t)nrec * tk::BwdLay::GREC;
    p.slab_bytes = sizeof(float) * (size_t)p.wg_ctas * L.total;
    return p;
}

extern "C" size_t stove_dynloop_bwd_workspace(const stove_gnn_cfg* cfg, int64_t n, int T, int skip) {
    if (gnn_check(cfg) || n <= 0 || T <= skip) return 0;
    GnnLayout L = gnn_layout(cfg);
    DynloopBwdPlan p = dynloop_bwd_plan(cfg, L, n, T - skip);
    return p.fast ? p.xrec_bytes + p.grec_bytes + p.slab_bytes : p.carry_bytes + p.gnn_ws;
}

extern "C" int64_t stove_dynloop_xrec_floats(const stove_gnn_cfg* cfg, int64_t n, int T, int skip) {
    if (gnn_check(cfg) || n <= 0 || T <= skip) return 0;
    GnnLayout L = gnn_layout(cfg);
    if (!tk::supported(cfg, L) || stove_opt(OPT_DYNLOOP_RECOMPUTE)) return 0;
    return (int64_t)n * (T - skip) * tk::BwdLay::XREC;
}

extern "C" int stove_dynloop_bwd2(const stove_gnn_cfg* cfg, const stove_fuse_cfg* fuse, int64_t n,
                                  const stove_dynloop_io* io, const float* weights, float* g_weights,
                                  void* workspace, void* stream, void* wgrad_stream) {
    int rc = dynloop_check(cfg, fuse, n, io, weights);
    if (rc) return rc;
    STOVE_CHECK_ARG(io->g_z_init && io->g_sup && io->g_sup_std && g_weights && workspace, "null gradient buffer");
    GnnLayout L = gnn_layout(cfg);
    cudaStream_t st = (cudaStream_t)stream;
    const int O = cfg->num_obj, Z = cfg->cl / 2 + 2, T = io->T, S = io->T - io->skip;
    if (n == 0) {
        STOVE_CUDA(cudaMemsetAsync(g_weights, 0, sizeof(float) * L.total, st));
        return STOVE_OK;
    }
    DynloopBwdPlan p = dynloop_bwd_plan(cfg, L, n, S);
    if (!p.fast) {
        // time steps before `skip` receive no gradient through the loop (the loop kernel zeroes them itself)
        STOVE_CUDA(cudaMemsetAsync(io->g_sup, 0, sizeof(float) * (size_t)n * T * O * 6, st));
        STOVE_CUDA(cudaMemsetAsync(io->g_sup_std, 0, sizeof(float) * (size_t)n * T * O * 6, st));
        float* carry[2] = {(float*)workspace, (float*)workspace + (size_t)n * O * Z};
        void* ws = (char*)workspace + p.carry_bytes;
        for (int k = S - 1; k >= 0; --k) {
            const int t = io->skip + k;
            stove_dynstep_io s = step_io(cfg, io, n, k);
            if (io->g_z) { s.g_z_a = io->g_z + (int64_t)k * O * Z; s.g_z_a_ss = (int64_t)S * O * Z; }
            if (k < S - 1) { s.g_z_b = carry[(k + 1) % 2]; s.g_z_b_ss = O * Z; }
            s.g_sc_ss = S;
            s.g_logq = io->g_logq ? io->g_logq + k : nullptr;
            s.g_trans = io->g_trans ? io->g_trans + k : nullptr;
            s.g_reward = io->g_reward ? io->g_reward + k : nullptr;
            s.g_z_prev = (k == 0) ? io->g_z_init : carry[k % 2];
            s.g_z_prev_ss = O * Z;
            s.g_sup = io->g_sup + (int64_t)t * O * 6; s.g_sup_std = io->g_sup_std + (int64_t)t * O * 6;
            s.g_sup_ss = (int64_t)T * O * 6;
            rc = stove_dynstep_bwd(cfg, fuse, n, &s, weights, g_weights, k == S - 1, k == 0, ws, stream);
            if (rc) return rc;
        }
        const int64_t items = n * O * 6;
        STOVE_KERNEL(K_DYNSTEP_BWD, st, tk::init_grad_to_sup_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(
            n, T, io->skip, O, Z, io->g_z_init, io->g_sup));
        STOVE_LAUNCH_CHECK();
        return STOVE_OK;
    }
    tk::TW tw;
    tk::StageTable tab;
    tk::build_tables(cfg, L, &tw, &tab);
    const int NW = stove_opt(OPT_DYNLOOP_NW);
    const FuseCfg f = make_fuse(cfg, fuse);
    const tk::LoopIO lio = to_loop_io(io);
    // workspace: gradient records | slabs | activation records (only if the forward pass kept none)
    float* grec = (float*)workspace;
    float* slabs = (float*)((char*)workspace + p.grec_bytes);
    const bool saved = io->xrec != nullptr;
    float* xrec = saved ? io->xrec : (float*)((char*)workspace + p.grec_bytes + p.slab_bytes);
#define DYNLOOP_BWD_LAUNCH(NW_, SV_)                                                                              \
    do {                                                                                                          \
        STOVE_CUDA(cudaFuncSetAttribute(tk::dynloop_bwd_kernel<NW_, SV_>,                                         \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));               \
        STOVE_KERNEL(K_DYNLOOP_BWD, st, tk::dynloop_bwd_kernel<NW_, SV_><<<p.ctas, 32 * NW_ * p.tpc, p.smem, st>>>( \
            *cfg, tw, tab, f, n, lio, weights, xrec, grec));                                                      \
    } while (0)
    if (NW == 2) { if (saved) DYNLOOP_BWD_LAUNCH(2, true); else DYNLOOP_BWD_LAUNCH(2, false); }
    else { if (saved) DYNLOOP_BWD_LAUNCH(1, true); else DYNLOOP_BWD_LAUNCH(1, false); }
#undef DYNLOOP_BWD_LAUNCH
    STOVE_LAUNCH_CHECK();
    // the weight gradients are not needed by anything downstream on the chain: they may run on
    // another stream (no join here -- the caller consumes g_weights on that stream)
    cudaStream_t wst = (cudaStream_t)wgrad_stream;
    if (wst != st) {
        StoveFork* fk = stove_fork_get(2);
        if (!fk) return STOVE_ERR_CUDA;
        STOVE_CUDA(cudaEventRecord(fk->fork_ev, st));
        STOVE_CUDA(cudaStreamWaitEvent(wst, fk->fork_ev, 0));
    }
    STOVE_CUDA(cudaMemsetAsync(slabs, 0, p.slab_bytes, wst));
    const size_t wg_smem = sizeof(float) * 2 * tk::BwdLay::REC;
    STOVE_CUDA(cudaFuncSetAttribute(tk::dynloop_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg_smem));
    STOVE_KERNEL(K_DYNLOOP_WGRAD, wst, tk::dynloop_wgrad_kernel<<<p.wg_ctas, 512, wg_smem, wst>>>(*cfg, L, n * S, xrec, grec, slabs));
    STOVE_LAUNCH_CHECK();
    STOVE_KERNEL(K_GNN_REDUCE_SLABS, wst, tk::dynloop_reduce_kernel<<<(L.total + 255) / 256, 256, 0, wst>>>(slabs, p.wg_ctas, L.total, g_weights));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

// fast path of stove_gnn_rollout (gnn.cu falls back to its CTA-wide kernel when this returns 1)
int stove_team_rollout(const stove_gnn_cfg* cfg, const GnnLayout& L, int64_t n, int num, const float* z_last,
                       const float* actions, int action_len, const float* app, const float* weights,
                       const float* noise, float pos_var, float vel_std, float latent_std, float* z_out,
                       float* std_out, float* logq_out, float* rewards, cudaStream_t st) {
    if (!tk::supported(cfg, L) || stove_opt(OPT_ROLLOUT_CTA)) return 1;
    tk::TW tw;
    tk::StageTable tab;
    tk::build_tables(cfg, L, &tw, &tab);
    const int NW = stove_opt(OPT_ROLLOUT_NW);
    const int max_tpc = (int)((kMaxSmem / sizeof(float) - tw.total) / tk::FwdLay::TOTAL);
    const int tpc = tk::pick_tpc(n, max_tpc < 7 ? max_tpc : 7);
    const size_t smem = sizeof(float) * ((size_t)tw.total + (size_t)tpc * tk::FwdLay::TOTAL);
    const int64_t groups = (n + tpc - 1) / tpc;
    const int ctas = (int)(groups < 148 ? groups : 148);
    if (NW == 2) {
        STOVE_CUDA(cudaFuncSetAttribute(tk::team_rollout_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        STOVE_KERNEL(K_GNN_ROLLOUT, st, tk::team_rollout_kernel<2><<<ctas, 64 * tpc, smem, st>>>(
            *cfg, tw, tab, n, num, z_last, actions, action_len, app, weights, noise, pos_var, vel_std, latent_std,
            z_out, std_out, logq_out, rewards));
    } else {
        STOVE_CUDA(cudaFuncSetAttribute(tk::team_rollout_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        STOVE_KERNEL(K_GNN_ROLLOUT, st, tk::team_rollout_kernel<1><<<ctas, 32 * tpc, smem, st>>>(
            *cfg, tw, tab, n, num, z_last, actions, action_len, app, weights, noise, pos_var, vel_std, latent_std,
            z_out, std_out, logq_out, rewards));
    }
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}
