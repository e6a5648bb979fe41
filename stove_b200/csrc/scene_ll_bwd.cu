// Fused scene likelihood, backward chain: ONE launch takes the gradients of the per-frame / per-patch log-likelihoods
// (and of the overlap ratios) to d/dz, and leaves the node gradients the parameter-gradient kernels of
// spn_obj.cu / spn_bg.cu need in their workspaces.  Replaces spn2_bwd_nodes + spn2_bwd_input, spn1_bwd_root +
// spn1_bwd_input and scene_bwd (5 launches, 4.3 + 7.3 + 7.3 MB of gradients handed through L2 and the glimpse
// backward waiting for four SPN kernels) on the path of Supair.likelihood (supair.py:62-94).
//
// Same partition as the forward kernel (scene_ll.cu): a CTA owns a contiguous group of frames, rounds of <= MAXF.
// Phases of a round (formulas: tests/kernel_spec.py, checked against autograd of the oracle):
//   N1  (root partition r, tile): the two halves of a warp own the two mid regions of the partition; gradient of
//       their sum values -> shared memory, (c, eA, eB) -> workspace for the sum-weight gradients.
//   N2  (mid region q, tile): lane = (leaf h, patch).  Linear-domain sums from the saved values (T = exp(sum - m0 - m1));
//       each half walks half of the 100 products and the halves exchange partial leaf-vector gradients by shuffle.
//       -> gl[leaf][patch][12] in shared memory (+ workspace copies for the leaf / sum parameter gradients).
//   IN  (tile, pixel pair): lane = (pixel of the pair, patch): d/d glimpse pixel and d/d mask over the 6 leaves the
//       pixel belongs to (polynomial leaf table: three sums, x enters at the end); the (x, mask) tile becomes the
//       (g_x, g_mask) tile in place.
//   BR  warp = frame: background root -> leaf-vector gradients (8-lane groups = root partitions).
//   BI  lane = pixel (its 3 x 18 leaf parameters in registers, read ONCE from L2), loop over the frames of the round:
//       d/d background mask -> the .y half of the frame buffer.
//   SC  warp = frame: objects in reverse order: clamp backward + row / column sums of the paste gradient on the rows
//       the box touches, then the 100 sample points (glimpse and mask gradients from the tile, bilinear derivatives,
//       scatter into the background gradient).  The background state before each object is REPLAYED from the tents
//       of the earlier objects (2 FMA + clamp per object) instead of stored: no O x frame buffer.
// Sequence mode (stove_scene_seq): per-frame weights from the scalar d loss / d elbo (what elbo_bwd_kernel computes),
// gradients written straight into those of z_sup / z_s / log q / trans (what zall_bwd_kernel and an add did).
#include "common.cuh"
#include "scene_math.cuh"
#include "spn_math.cuh"
#include "scene_ll.cuh"

int spn2_param_kernels(const stove_spn2_struct* st, int64_t N, const float* x, const float* marg, const float* leaf,
                       const float* wlin, const float* rlin, void* workspace, float* g_leaf, float* g_wlog,
                       float* g_rlog, cudaStream_t s_leaf, cudaStream_t s_sum);
int spn1_param_kernels(const stove_spn1_struct* st, int64_t N, const float* x, const float* marg, const float* leaf,
                       const float* rlin, void* workspace, float* g_leaf, float* g_rlog, cudaStream_t s_leaf,
                       cudaStream_t s_root);

namespace sl {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------
// N1: root partition -> gradients of the two regions' sum values
// ------------------------------------------------------------------------------------
template <int S>
__device__ void bwd_root_task(const LLArgs& a, const SmemB& m, float* smem, int r, int tile, int npt, int64_t n0g,
                              int lane) {
    const int h = lane >> 4, pt = lane & (HT - 1);
    const int R = a.st.R, Q = 2 * R;
    const bool live = pt < npt;
    const int64_t n = n0g + pt;
    const int qo = 2 * r + h, qt = 2 * r + 1 - h;
    float own[S], oth[S], eO[S], eT[S];
#pragma unroll
    for (int i = 0; i < S; ++i) {
        own[i] = live ? a.sum_val[(int64_t)(qo * S + i) * a.npad_p + n] : 0.f;
        oth[i] = live ? a.sum_val[(int64_t)(qt * S + i) * a.npad_p + n] : 0.f;
    }
    const float mO = shift_exp<S>(own, eO), mT = shift_exp<S>(oth, eT);
    // h = 0 owns the A side (index i of w[j][i]) and needs columns: the transposed copy; h = 1 owns the B side: rows
    const float* rw = smem + (h ? m.rws : m.rwsT) + r * S * S;
    float vec[S];
    float U = 0.f;
#pragma unroll
    for (int k = 0; k < S; ++k) {
        float v = 0.f;
#pragma unroll
        for (int l = 0; l < S; ++l) v = fmaf(eT[l], lds_f1(rw + k * S + l), v);
        vec[k] = v;
        U = fmaf(eO[k], v, U);
    }
    const float U0 = __shfl_sync(0xffffffffu, U, pt);          // both halves use the value of the h = 0 lane
    const bool fast = U0 > LIN_SUM_FLOOR;
    float go = 0.f;
    if (live) {
        if (a.seq) {             // d loss / d (raw object log-likelihood) = frame weight * sx * sy (supair.py:79)
            const unsigned fu = (unsigned)n / (unsigned)a.O;
            const int64_t f = fu;
            const float4 z = load_z(a, f, (int)((unsigned)n - fu * (unsigned)a.O));
            go = frame_weight(a, f) * z.x * z.y;
        } else {
            go = a.g_obj[n];
        }
    }
    const float ov = live ? a.out_obj[n] : 0.f;
    float c = 0.f;
    if (fast) c = go * expf(mO + mT + logf(U0) - ov) / U0;
    float* gs = smem + m.t + ((size_t)tile * Q * S + qo * S) * HT + pt;
#pragma unroll
    for (int k = 0; k < S; ++k) gs[k * HT] = c * eO[k] * vec[k];
    if (live) {
        float* ar = a.aux_root + (int64_t)r * (1 + 2 * S) * a.npad_p + n;
        if (h == 0) ar[0] = c;
#pragma unroll
        for (int k = 0; k < S; ++k) ar[(int64_t)(1 + h * S + k) * a.npad_p] = fast ? eO[k] : 0.f;
    }
    __syncwarp();
    if (!fast && h == 0 && live) {
        const float* a0 = a.sum_val + (int64_t)((2 * r) * S) * a.npad_p + n;
        const float* b0 = a.sum_val + (int64_t)((2 * r + 1) * S) * a.npad_p + n;
        const float val = slow_logsumexp(a0, b0, (int)a.npad_p, S, a.rlog + r * S * S, 1);
        const float gr = go * expf(val - ov);
        if (gr != 0.f && val > -INFINITY)
            slow_sum_backward(a0, b0, (int)a.npad_p, S, a.rlog + r * S * S, 1, val, gr,
                              smem + m.t + ((size_t)tile * Q * S + (2 * r) * S) * HT + pt,
                              smem + m.t + ((size_t)tile * Q * S + (2 * r + 1) * S) * HT + pt, HT, a.g_rlog + r * S * S);
    }
}

// ------------------------------------------------------------------------------------
// N2: mid region -> gradients of its two leaf vectors
// ------------------------------------------------------------------------------------
template <int G, int S>
__device__ void bwd_region_task(const LLArgs& a, const SmemB& m, float* smem, int q, int tile, int npt, int64_t n0g,
                                int lane) {
    constexpr int SP = GP_<S>::v, SH = 5, JH = G / 2;
    static_assert(S > 8 && S <= 10 && SP == 12 && G == 10, "laid out for 10 Gaussians and 9-10 sums");
    const int h = lane >> 4, pt = lane & (HT - 1);
    const int Q = 2 * a.st.R;
    const bool live = pt < npt;
    const int64_t n = n0g + pt;
    const float* lv0 = a.leaf_val + (int64_t)(q * 2) * G * a.npad_p + n;
    float L[G], e[G], ep[G];
#pragma unroll
    for (int g = 0; g < G; ++g) L[g] = live ? lv0[(int64_t)(h * G + g) * a.npad_p] : 0.f;
    const float mo = shift_exp<G>(L, e);
    const float mp = __shfl_xor_sync(0xffffffffu, mo, 16);
#pragma unroll
    for (int g = 0; g < G; ++g) ep[g] = __shfl_xor_sync(0xffffffffu, e[g], 16);
    float e0[G], e1[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        e0[g] = h ? ep[g] : e[g];
        e1[g] = h ? e[g] : ep[g];
    }
    // linear-domain value of the sums this half owns: sum_val = m0 + m1 + log T  =>  T = exp(sum_val - m0 - m1), 5
    // exponentials instead of recomputing 100 products x 5 sums (a third of this pass); the difference of two numbers
    // of magnitude ~1e2 costs ~2e-5 of relative accuracy in T, far inside the gradient tolerance
    float T[SH];
#pragma unroll
    for (int c = 0; c < SH; ++c) {
        const int s = sum_of(h, c);
        T[c] = (live && s < S) ? expf(a.sum_val[(int64_t)(q * S + s) * a.npad_p + n] - mo - mp) : 1.f;
    }
    const float* gsrow = smem + m.t + ((size_t)tile * Q * S + q * S) * HT + pt;
    float qv[SH], qvp[SH];
    unsigned slow_mask = 0;
#pragma unroll
    for (int c = 0; c < SH; ++c) {
        const int s = sum_of(h, c);
        const float gs = (s < S) ? gsrow[s * HT] : 0.f;
        if (T[c] > LIN_SUM_FLOOR) {
            qv[c] = gs / T[c];
        } else {
            qv[c] = 0.f;
            if (gs != 0.f && live && s < S) slow_mask |= 1u << c;
        }
    }
#pragma unroll
    for (int c = 0; c < SH; ++c) qvp[c] = __shfl_xor_sync(0xffffffffu, qv[c], 16);
    float qall[12];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        qall[s] = h ? qvp[s] : qv[s];
        qall[4 + s] = h ? qv[s] : qvp[s];
    }
    qall[8] = h ? qvp[4] : qv[4];
    qall[9] = h ? qv[4] : qvp[4];
    qall[10] = 0.f;
    qall[11] = 0.f;
    // this half walks the products with j in [JH h, JH h + JH)
    float e1o[JH];
#pragma unroll
    for (int jj = 0; jj < JH; ++jj) e1o[jj] = h ? e1[JH + jj] : e1[jj];
    float acc0[G], acc1[JH];
#pragma unroll
    for (int g = 0; g < G; ++g) acc0[g] = 0.f;
#pragma unroll
    for (int jj = 0; jj < JH; ++jj) acc1[jj] = 0.f;
    {
        const float* wb = smem + m.wl + (q * G * G + JH * h * G) * SP;
#pragma unroll
        for (int jj = 0; jj < JH; ++jj) {
#pragma unroll
            for (int i = 0; i < G; ++i) {
                const float* wk = wb + (jj * G + i) * SP;
                const float4 w0 = lds_f4(wk), w1 = lds_f4(wk + 4), w2 = lds_f4(wk + 8);
                float v = qall[0] * w0.x;
                v = fmaf(qall[1], w0.y, v); v = fmaf(qall[2], w0.z, v); v = fmaf(qall[3], w0.w, v);
                v = fmaf(qall[4], w1.x, v); v = fmaf(qall[5], w1.y, v); v = fmaf(qall[6], w1.z, v);
                v = fmaf(qall[7], w1.w, v); v = fmaf(qall[8], w2.x, v); v = fmaf(qall[9], w2.y, v);
                acc0[i] = fmaf(e1o[jj], v, acc0[i]);
                acc1[jj] = fmaf(e0[i], v, acc1[jj]);
            }
        }
    }
    float gl[GLP];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const float p0 = __shfl_xor_sync(0xffffffffu, acc0[g], 16);
        const float p1 = __shfl_xor_sync(0xffffffffu, acc1[g < JH ? g : g - JH], 16);
        // h = 0 owns leaf 0: both halves' partial sums over j; h = 1 owns leaf 1: j < JH from the partner, else its own
        const float tot = h ? (g < JH ? p1 : acc1[g < JH ? 0 : g - JH]) : acc0[g] + p0;
        gl[g] = e[g] * tot;
    }
    gl[10] = 0.f;
    gl[11] = 0.f;
    float* gt = smem + m.v + (((size_t)tile * Q * 2 + q * 2 + h) * HT + pt) * GLP;
    reinterpret_cast<float4*>(gt)[0] = make_float4(gl[0], gl[1], gl[2], gl[3]);
    reinterpret_cast<float4*>(gt)[1] = make_float4(gl[4], gl[5], gl[6], gl[7]);
    reinterpret_cast<float4*>(gt)[2] = make_float4(gl[8], gl[9], 0.f, 0.f);
    if (live) {
        float* aq = a.aux_reg + (int64_t)q * (2 * G + S) * a.npad_p + n;
#pragma unroll
        for (int g = 0; g < G; ++g) aq[(int64_t)(h * G + g) * a.npad_p] = e[g];
#pragma unroll
        for (int c = 0; c < SH; ++c) {
            const int s = sum_of(h, c);
            if (s < S) aq[(int64_t)(2 * G + s) * a.npad_p] = qv[c];
        }
    }
    if (__any_sync(0xffffffffu, slow_mask != 0)) {
        // exact log-domain backward of the sums whose linear value underflowed (rare); the two halves of a patch take
        // turns because both add into the two leaf vectors of the region
        __syncwarp();
        float* g0 = smem + m.v + (((size_t)tile * Q * 2 + q * 2) * HT + pt) * GLP;
        for (int phase = 0; phase < 2; ++phase) {
            if (h == phase && slow_mask)
                for (int c = 0; c < SH; ++c)
                    if (slow_mask & (1u << c)) {
                        const int s = sum_of(h, c);
                        const float sumv = a.sum_val[(int64_t)(q * S + s) * a.npad_p + n];
                        if (sumv > -INFINITY)
                            slow_sum_backward(lv0, lv0 + (int64_t)G * a.npad_p, (int)a.npad_p, G,
                                              a.wlog + (int64_t)q * G * G * SP + s, SP, sumv, gsrow[s * HT], g0,
                                              g0 + HT * GLP, 1, a.g_wlog + (int64_t)q * G * G * SP + s);
                    }
            __syncwarp();
        }
#pragma unroll
        for (int g = 0; g < G; ++g) gl[g] = gt[g];
    }
    if (live) {
        float* gg = a.gleaf + (int64_t)((q * 2 + h) * G) * a.npad_p + n;
#pragma unroll
        for (int g = 0; g < G; ++g) gg[(int64_t)g * a.npad_p] = gl[g];
    }
}

// ------------------------------------------------------------------------------------
// IN: (x, mask) tile -> (g_x, g_mask) tile
//   L = -w (a d^2 + b):  dL/dx = -2 w a d,  dL/dmask = +(a d^2 + b) where the mask is inside [0, 1]
// ------------------------------------------------------------------------------------
template <int G>
__device__ void bwd_input_task(const LLArgs& a, const SmemB& m, float* smem, int tile, int pp, int lane) {
    constexpr int GP = GP_<G>::v;
    const int hh = lane >> 4, pt = lane & (HT - 1);
    const int D = a.st.D, R = a.st.R, Q = 2 * R;
    const int px = 2 * pp + hh;
    if (px >= D) return;
    float2* xe = reinterpret_cast<float2*>(smem + m.t) + (size_t)tile * D * HT + px * HT + ((pt + px) & (HT - 1));
    const float2 v = *xe;
    const float wv = 1.f - fminf(fmaxf(v.y, 0.f), 1.f);
    const bool inside = v.y >= 0.f && v.y <= 1.f;
    const int32_t* sl = reinterpret_cast<const int32_t*>(smem + m.slots) + px * R;
    const float* glt = smem + m.v + ((size_t)tile * Q * 2 * HT + pt) * GLP;
    // u = a (x - mu)^2 + b = c2 x^2 - c1 x + c0 and du/dx = 2 c2 x - c1 (polynomial leaf table): the sums over the
    // Gaussians and over the R leaves of the pixel are three FMAs per (leaf, Gaussian); x enters once at the end
    float A1 = 0.f, A2 = 0.f, A3 = 0.f;          // sum gl c2, sum gl c1, sum gl c0
    for (int r = 0; r < R; ++r) {
        const int pk = sl[r];
        float c1[GP], c2[GP], c0[GP];
        load_leaf_params<G, true>(smem + m.lf + (pk & 0xffff) * 3 * GP, c1, c2, c0);
        const float* gp = glt + (size_t)(pk >> 16) * HT * GLP;
        const float4 g0 = lds_f4(gp), g1 = lds_f4(gp + 4), g2 = lds_f4(gp + 8);
        const float gv[12] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x, g2.y, g2.z, g2.w};
#pragma unroll
        for (int g = 0; g < G; ++g) {
            A1 = fmaf(gv[g], c2[g], A1);
            A2 = fmaf(gv[g], c1[g], A2);
            A3 = fmaf(gv[g], c0[g], A3);
        }
    }
    // L = -w u:  dL/dx = -w (2 x A1 - A2),  dL/dmask = +u summed = x^2 A1 - x A2 + A3 where the mask is inside [0, 1]
    *xe = make_float2(-wv * fmaf(2.f * v.x, A1, -A2), inside ? fmaf(v.x, fmaf(v.x, A1, -A2), A3) : 0.f);
}

// ------------------------------------------------------------------------------------
// BR: background root, warp = frame, 8-lane group = root partition, lane j = Gaussian index
// ------------------------------------------------------------------------------------
template <int RB, int GB>
__device__ void bwd_bg_root_frame(const LLArgs& a, const SmemB& m, float* smem, int fi, int64_t f, int lane) {
    static_assert(GB <= 8 && RB <= 4, "8-lane groups");
    const int r = lane >> 3, j = lane & 7;
    const bool on = r < RB && j < GB;
    const int rr = on ? r : 0, jj = on ? j : 0;
    const float av = on ? a.bleaf_val[(int64_t)((2 * rr) * GB + jj) * a.npad_f + f] : -INFINITY;
    const float bv = on ? a.bleaf_val[(int64_t)((2 * rr + 1) * GB + jj) * a.npad_f + f] : -INFINITY;
    float mA = av, mB = bv;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, o));
        mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, o));
    }
    const float eA = on ? expf(av - mA) : 0.f, eB = on ? expf(bv - mB) : 0.f;
    float rowB = 0.f, colA = 0.f;          // rowB_j = sum_i eA_i w[j][i];  colA_i = sum_j eB_j w[j][i]  (lane index = j resp. i)
#pragma unroll
    for (int k = 0; k < GB; ++k) {
        const float ea = __shfl_sync(0xffffffffu, eA, (lane & ~7) + k);
        const float eb = __shfl_sync(0xffffffffu, eB, (lane & ~7) + k);
        if (on) {
            rowB = fmaf(ea, __ldg(a.brlin + rr * GB * GB + jj * GB + k), rowB);
            colA = fmaf(eb, __ldg(a.brlin + rr * GB * GB + k * GB + jj), colA);
        }
    }
    float U = eB * rowB;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) U += __shfl_xor_sync(0xffffffffu, U, o);
    const float go = a.seq ? frame_weight(a, f) : a.g_bg[f], ov = a.out_bg[f];
    const bool fast = U > LIN_SUM_FLOOR;
    float c = 0.f;
    if (fast) c = go * expf(mA + mB + logf(U) - ov) / U;
    const float gA = c * eA * colA, gB = c * eB * rowB;
    float* bgl = smem + m.bgl + (size_t)fi * 2 * RB * BGLP;
    if (r < RB) {
        bgl[(2 * r) * BGLP + j] = on ? gA : 0.f;
        bgl[(2 * r + 1) * BGLP + j] = on ? gB : 0.f;
    }
    if (on) {
        a.bgleaf[(int64_t)((2 * r) * GB + j) * a.npad_f + f] = gA;
        a.bgleaf[(int64_t)((2 * r + 1) * GB + j) * a.npad_f + f] = gB;
        float* ar = a.baux_root + (int64_t)r * (1 + 2 * GB) * a.npad_f + f;
        if (j == 0) ar[0] = c;
        ar[(int64_t)(1 + j) * a.npad_f] = fast ? eA : 0.f;
        ar[(int64_t)(1 + GB + j) * a.npad_f] = fast ? eB : 0.f;
    }
    __syncwarp();
    if (!fast && r < RB && j == 0) {
        // exact log-domain backward of this root partition (rare): one lane walks all pairs
        const float* a0 = a.bleaf_val + (int64_t)((2 * r) * GB) * a.npad_f + f;
        const float* b0 = a.bleaf_val + (int64_t)((2 * r + 1) * GB) * a.npad_f + f;
        const float* wl = a.brlog + r * GB * GB;
        float M = -INFINITY;
        for (int y = 0; y < GB; ++y)
            for (int x = 0; x < GB; ++x) M = fmaxf(M, a0[x * a.npad_f] + b0[y * a.npad_f] + wl[y * GB + x]);
        if (M > -INFINITY) {
            float acc = 0.f;
            for (int y = 0; y < GB; ++y)
                for (int x = 0; x < GB; ++x) acc += expf(a0[x * a.npad_f] + b0[y * a.npad_f] + wl[y * GB + x] - M);
            const float val = M + logf(acc);
            const float gr = go * expf(val - ov);
            if (gr != 0.f)
                for (int y = 0; y < GB; ++y)
                    for (int x = 0; x < GB; ++x) {
                        const float resp = gr * expf(a0[x * a.npad_f] + b0[y * a.npad_f] + wl[y * GB + x] - val);
                        bgl[(2 * r) * BGLP + x] += resp;
                        bgl[(2 * r + 1) * BGLP + y] += resp;
                        a.bgleaf[(int64_t)((2 * r) * GB + x) * a.npad_f + f] += resp;
                        a.bgleaf[(int64_t)((2 * r + 1) * GB + y) * a.npad_f + f] += resp;
                        atomicAdd(a.g_brlog + r * GB * GB + y * GB + x, resp);
                    }
        }
    }
}

// ------------------------------------------------------------------------------------
// BI: d/d background mask; lane = pixel, its leaf parameters stay in registers for all frames of the round
//   dL/dmask = + sum_{r, g} gl[r, side][g] (a d^2 + b)   (the final mask is inside [0, 1] by construction)
// ------------------------------------------------------------------------------------
template <int RB, int GB>
__device__ void bwd_bg_input_task(const LLArgs& a, const SmemB& m, float* smem, int chunk, int nfr, int lane) {
    constexpr int GPB = GP_<GB>::v;
    static_assert(GPB == BGLP, "leaf-vector rows are padded like the parameter rows");
    const int px = chunk * 32 + lane;
    const bool ok = px < a.Dbg;
    const int pxc = ok ? px : 0;
    float mu[RB][GPB], aa[RB][GPB], bb[RB][GPB];
    unsigned gaddr[RB];
    const unsigned bgl_s = smem_u32(smem + m.bgl);
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        // interleaved table, rows (r, pixel): block of 32 pixels = [6 parts][32 lanes] float4 -> coalesced 512-byte loads
        const float4* p4 = reinterpret_cast<const float4*>(a.bleaf_il) + ((int64_t)(r * a.il_stride + chunk * 32) >> 5) * (6 * 32) + lane;
        const float4 m0 = __ldg(p4), m1 = __ldg(p4 + 32), a0 = __ldg(p4 + 64), a1 = __ldg(p4 + 96), b0 = __ldg(p4 + 128),
                     b1 = __ldg(p4 + 160);
        mu[r][0] = m0.x; mu[r][1] = m0.y; mu[r][2] = m0.z; mu[r][3] = m0.w; mu[r][4] = m1.x; mu[r][5] = m1.y; mu[r][6] = m1.z; mu[r][7] = m1.w;
        aa[r][0] = a0.x; aa[r][1] = a0.y; aa[r][2] = a0.z; aa[r][3] = a0.w; aa[r][4] = a1.x; aa[r][5] = a1.y; aa[r][6] = a1.z; aa[r][7] = a1.w;
        bb[r][0] = b0.x; bb[r][1] = b0.y; bb[r][2] = b0.z; bb[r][3] = b0.w; bb[r][4] = b1.x; bb[r][5] = b1.y; bb[r][6] = b1.z; bb[r][7] = b1.w;
        gaddr[r] = bgl_s + (unsigned)((2 * r + __ldg(a.bg_side + pxc * RB + r)) * BGLP) * 4u;
    }
    unsigned addr = smem_u32(smem + m.fb) + (unsigned)pxc * 8u;
    const unsigned fstride = (unsigned)a.fs * 8u;
    // a real loop over the frames (two per trip): fully unrolled, the 13 copies of the body were 19 KB of straight-line
    // code that every warp runs twice -- 37 % of this phase's stall samples were instruction fetch
    unsigned goff = 0;
#pragma unroll 2
    for (int f = 0; f < nfr; ++f) {
        float xv;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xv) : "r"(addr));
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            float4 g0, g1;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(g0.x), "=f"(g0.y), "=f"(g0.z), "=f"(g0.w) : "r"(gaddr[r] + goff));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(g1.x), "=f"(g1.y), "=f"(g1.z), "=f"(g1.w) : "r"(gaddr[r] + goff + 16u));
            const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
            for (int g = 0; g < GB; ++g) {
                const float d = xv - mu[r][g];
                t = fmaf(gv[g], fmaf(aa[r][g] * d, d, bb[r][g]), t);
            }
        }
        if (ok) asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr + 4u), "f"(t) : "memory");
        addr += fstride;
        goff += 2 * RB * BGLP * 4;
    }
}

// ------------------------------------------------------------------------------------
// SC: glimpse / mask backward of one frame (warp = frame)
// ------------------------------------------------------------------------------------
// background mask before object o at pixel (u, v): the pastes of the earlier objects replayed from their tents
// (o is uniform over the warp: no divergence; 2 loads + FMA + clamp per earlier object)
__device__ __forceinline__ float replay_bg(const float* tt, int ts, int tXs, int o, int u, int v) {
    float b = 0.f;
    for (int k = 0; k < o; ++k) {
        const float* t = tt + k * ts;
        b = fminf(fmaxf(fmaf(t[2 * tXs + u], t[v], b), 0.f), 1.f);
    }
    return b;
}

// Branch-free corner set of a bilinear sample: clamped addresses + 0/1 validity factors.  (The first version
// predicated every corner access and replay: 45 % of this phase's instructions were branches, predicate set-up and
// reconvergence barriers, profiles/r02_ncu_scene_ll_bwd_v1.txt.)
struct Corner4 {
    int y0, y1, x0, x1;          // clamped into the frame
    float fy, fx, k00, k01, k10, k11;
};
__device__ __forceinline__ Corner4 corners4(float py, float px, int A, int B) {
    Corner4 c;
    const float fy0 = floorf(py), fx0 = floorf(px);
    c.fy = py - fy0;
    c.fx = px - fx0;
    const int y0 = (int)fminf(fmaxf(fy0, -2.f), (float)A), x0 = (int)fminf(fmaxf(fx0, -2.f), (float)B);
    const float oy0 = (y0 >= 0 && y0 < A) ? 1.f : 0.f, oy1 = (y0 + 1 >= 0 && y0 + 1 < A) ? 1.f : 0.f;
    const float ox0 = (x0 >= 0 && x0 < B) ? 1.f : 0.f, ox1 = (x0 + 1 >= 0 && x0 + 1 < B) ? 1.f : 0.f;
    c.k00 = oy0 * ox0; c.k01 = oy0 * ox1; c.k10 = oy1 * ox0; c.k11 = oy1 * ox1;
    c.y0 = min(max(y0, 0), A - 1); c.y1 = min(max(y0 + 1, 0), A - 1);
    c.x0 = min(max(x0, 0), B - 1); c.x1 = min(max(x0 + 1, 0), B - 1);
    return c;
}

__device__ void scene_frame_bwd(const LLArgs& a, const SmemB& m, float* smem, int fi, int64_t f, int lane) {
    const int PP = a.pa * a.pb, D = a.st.D;
    const int tXs = up4(a.B), tYs = up4(a.A), ts = m.tent_stride;
    float2* fb = reinterpret_cast<float2*>(smem + m.fb) + (size_t)fi * a.fs;       // (x, gradient w.r.t. the background mask)
    float* tt = smem + m.v + (size_t)fi * a.O * ts;
    int32_t* rng = reinterpret_cast<int32_t*>(smem + m.rng) + fi * a.O * 2;
    for (int o = 0; o < a.O; ++o) {
        const float4 zz = load_z(a, f, o);
        float* t = tt + o * ts;
        int ulo, uhi;
        warp_tents(a, zz.x, zz.y, zz.z, zz.w, t, t + 2 * tXs, t + tXs, t + 2 * tXs + tYs, lane, ulo, uhi);
        if (lane == 0) { rng[2 * o] = ulo; rng[2 * o + 1] = uhi; }
    }
    const float rpb = recip_n(a.pb, a.align), rpa = recip_n(a.pa, a.align);
    const float kB = unnorm_slope(a.B, a.align), kA = unnorm_slope(a.A, a.align);
    const float oB = unnorm_offset(a.B, a.align), oA = unnorm_offset(a.A, a.align);
    const float rB = recip_n(a.B, a.align), rA = recip_n(a.A, a.align);
    const float rPP = 1.f / (float)PP;
    const int nc = (a.B + 31) >> 5;                       // columns per lane (B <= 32 SCENE_MAXC)
    const float2* xw = reinterpret_cast<const float2*>(smem + m.t);
    __syncwarp();
    for (int o = a.O - 1; o >= 0; --o) {
        float zq = 0.f;                                           // sy / sx (sequence mode)
        const float4 zz = load_z(a, f, o, &zq);
        const float sx = zz.x, sy = zz.y, tx = zz.z, ty = zz.w;
        const float isx = 1.f / sx, isy = 1.f / sy;
        const float* t = tt + o * ts;
        const float *tX = t, *dX = t + tXs, *tY = t + 2 * tXs, *dY = t + 2 * tXs + tYs;
        const int ulo = rng[2 * o], uhi = rng[2 * o + 1];
        float gsx = 0.f, gsy = 0.f, gtx = 0.f, gty = 0.f;        // per-lane partial sums, reduced once per object
        // (1) clamp backward (gradient passes where 0 <= bg + paste <= 1) and, with paste = tY[u] tX[v], the gradient of
        //     the two tents.  Rows outside [ulo, uhi] get no paste: bg + 0 is inside [0, 1], the tent derivative is 0.
        for (int c = 0; c < nc; ++c) {
            const int v = lane + 32 * c;
            const bool vok = v < a.B;
            const int vc = vok ? v : 0;
            const float txv = vok ? tX[vc] : 0.f;
            float colacc = 0.f;
            for (int u = ulo; u <= uhi; ++u) {
                const float tyu = tY[u];
                const float pre = fmaf(tyu, txv, replay_bg(tt, ts, tXs, o, u, vc));
                float g = vok ? fb[u * a.B + vc].y : 0.f;     // (idle lanes must not read what lane 0 is writing: racecheck)
                if (!(pre >= 0.f && pre <= 1.f)) g = 0.f;
                if (vok) fb[u * a.B + vc].y = g;
                // d paste / d tY[u] = tX[v]; tY[u] = tent(p_u), p_u affine in 1 / sy and ty / sy
                const float gy = g * txv * dY[u] * kA;
                gsy = fmaf(gy, -(base_coord_r(u, rA, a.align) - ty) * isy * isy, gsy);
                gty = fmaf(gy, -isy, gty);
                colacc = fmaf(g, tyu, colacc);
            }
            const float gx = colacc * (vok ? dX[vc] : 0.f) * kB;
            gsx = fmaf(gx, -(base_coord_r(vc, rB, a.align) - tx) * isx * isx, gsx);
            gtx = fmaf(gx, -isx, gtx);
        }
        __syncwarp();
        // (2) the sample points of the glimpse and of the mask
        const float fw = a.seq ? frame_weight(a, f) : 0.f;
        const float gov = a.seq ? -a.sq.beta * fw * rPP : (a.g_overlap ? __ldg(a.g_overlap + f * a.O + o) * rPP : 0.f);
        const int pl = fi * a.O + o, tile = pl / HT, pt = pl - tile * HT;
        const float2* xt = xw + (size_t)tile * D * HT;
        const float mx = sx * kB, ox = fmaf(tx, kB, oB), my = sy * kA, oy = fmaf(ty, kA, oA);
        // (a real loop: the unrolled copies of this body made the phase 53 KB of straight-line code, 30 % of its stall
        //  samples were instruction fetch; the glimpse coordinates are recomputed per pixel instead of kept in registers)
#pragma unroll 1
        for (int idx = lane; idx < PP; idx += 32) {
            {
                const int gi = idx / a.pb, gj = idx - gi * a.pb;
                const float xbi = base_coord_r(gj, rpb, a.align), ybi = base_coord_r(gi, rpa, a.align);
                const Corner4 c = corners4(fmaf(ybi, my, oy), fmaf(xbi, mx, ox), a.A, a.B);
                const float2 gt = xt[idx * HT + ((pt + idx) & (HT - 1))];       // (g_x, g_mask) of this glimpse pixel
                const float gm = gov + gt.y, gp = gt.x;
                const int a00 = c.y0 * a.B + c.x0, a01 = c.y0 * a.B + c.x1, a10 = c.y1 * a.B + c.x0, a11 = c.y1 * a.B + c.x1;
                // corner values of (1 - background before o) and of the frame, zero padded
                float b00 = 0.f, b01 = 0.f, b10 = 0.f, b11 = 0.f;
                for (int k = 0; k < o; ++k) {
                    const float* tk = tt + k * ts;
                    const float y0v = tk[2 * tXs + c.y0], y1v = tk[2 * tXs + c.y1], x0v = tk[c.x0], x1v = tk[c.x1];
                    b00 = fminf(fmaxf(fmaf(y0v, x0v, b00), 0.f), 1.f);
                    b01 = fminf(fmaxf(fmaf(y0v, x1v, b01), 0.f), 1.f);
                    b10 = fminf(fmaxf(fmaf(y1v, x0v, b10), 0.f), 1.f);
                    b11 = fminf(fmaxf(fmaf(y1v, x1v, b11), 0.f), 1.f);
                }
                const float m00 = c.k00 * (1.f - b00), m01 = c.k01 * (1.f - b01), m10 = c.k10 * (1.f - b10), m11 = c.k11 * (1.f - b11);
                const float q00 = c.k00 * fb[a00].x, q01 = c.k01 * fb[a01].x, q10 = c.k10 * fb[a10].x, q11 = c.k11 * fb[a11].x;
                const float mdy = (m10 + c.fx * (m11 - m10)) - (m00 + c.fx * (m01 - m00));
                const float mdx = (1.f - c.fy) * (m01 - m00) + c.fy * (m11 - m10);
                const float qdy = (q10 + c.fx * (q11 - q10)) - (q00 + c.fx * (q01 - q00));
                const float qdx = (1.f - c.fy) * (q01 - q00) + c.fy * (q11 - q10);
                // marg = 1 - sample(1 - bg): d marg / d p = -d sample
                const float dpx = fmaf(gp, qdx, -gm * mdx), dpy = fmaf(gp, qdy, -gm * mdy);
                gsx = fmaf(dpx * kB, xbi, gsx);
                gtx = fmaf(dpx, kB, gtx);
                gsy = fmaf(dpy * kA, ybi, gsy);
                gty = fmaf(dpy, kA, gty);
                // d marg / d bg_o = + bilinear weights (zero weight outside the frame)
                if (gm != 0.f) {
                    const float wy0 = 1.f - c.fy, wy1 = c.fy, wx0 = 1.f - c.fx, wx1 = c.fx;
                    if (c.k00 != 0.f) atomicAdd(&fb[a00].y, gm * wy0 * wx0);
                    if (c.k01 != 0.f) atomicAdd(&fb[a01].y, gm * wy0 * wx1);
                    if (c.k10 != 0.f) atomicAdd(&fb[a10].y, gm * wy1 * wx0);
                    if (c.k11 != 0.f) atomicAdd(&fb[a11].y, gm * wy1 * wx1);
                }
            }
        }
        gsx = warp_sum(gsx); gsy = warp_sum(gsy); gtx = warp_sum(gtx); gty = warp_sum(gty);
        if (!a.seq) {
            if (lane == 0) reinterpret_cast<float4*>(a.g_z)[f * a.O + o] = make_float4(gsx, gsy, gtx, gty);
        } else {
            // + the sx * sy weighting of the object term, then (sx, sy) -> (sx, q = sy / sx) and straight into the
            // gradients of the sequence tensors (what stove_elbo_bwd + stove_zall_bwd + one add used to do)
            const float lo = a.out_obj[f * a.O + o];
            gsx = fmaf(fw * lo, sy, gsx);
            gsy = fmaf(fw * lo, sx, gsy);
            const float g0 = fmaf(gsy, zq, gsx), g1 = gsy * sx;
            const int T = a.sq.T, S_ = T - a.sq.skip;
            const unsigned bu = (unsigned)f / (unsigned)(T - 1);
            const int64_t b = bu;
            const int t = (int)((unsigned)f - bu * (unsigned)(T - 1)) + 1;
            float4* gsup = reinterpret_cast<float4*>(a.sq.g_z_sup);
            if (lane == 0) {
                gsup[(b * T + t) * a.O + o] = t < a.sq.skip ? make_float4(g0, g1, gtx, gty) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (t == 1) gsup[(b * T) * a.O + o] = make_float4(0.f, 0.f, 0.f, 0.f);       // frame 0 is not scored
            }
            if (t >= a.sq.skip) {
                float* dst = a.sq.g_z_s + ((b * S_ + (t - a.sq.skip)) * a.O + o) * a.sq.Z;
                for (int k = lane; k < a.sq.Z; k += 32) dst[k] = k == 0 ? g0 : k == 1 ? g1 : k == 2 ? gtx : k == 3 ? gty : 0.f;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------
template <int G, int S, int RB, int GB>
__global__ void __launch_bounds__(MAXNW * 32, 1) scene_ll_bwd_kernel(const __grid_constant__ LLArgs a) {
    constexpr int GP = GP_<G>::v, SP = GP_<S>::v;
    extern __shared__ __align__(16) float smem[];
    const SmemB m = smem_layout_bwd(a, G, S, GB, RB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int R = a.st.R, Q = 2 * R, D = a.st.D, AB = a.A * a.B;
    const int64_t per = a.F / gridDim.x, rem = a.F % gridDim.x;
    const int64_t f0 = blockIdx.x * per + min((int64_t)blockIdx.x, rem);
    const int cnt = (int)(per + (blockIdx.x < rem ? 1 : 0));

    if (a.seq) {
        // d elbo / d log q = -1 / (n S), d elbo / d trans = +1 / (n S)   (stove.py:747)
        const int64_t nS = a.sq.n * (a.sq.T - a.sq.skip);
        const float we = __ldg(a.sq.g_elbo) * a.w_elbo;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + tid; i < nS; i += (int64_t)gridDim.x * blockDim.x) {
            a.sq.g_logq[i] = -we;
            a.sq.g_trans[i] = we;
        }
    }
    if (blockIdx.x == gridDim.x - 1) {
        // columns [N, npad) of the workspaces belong to nobody; the parameter-gradient kernels copy whole 32-column
        // tiles (and ignore those columns): give them defined values (compute-sanitizer initcheck)
        for (int64_t n = a.Np + tid; n < a.npad_p; n += blockDim.x) {
            for (int row = 0; row < Q * 2 * G; ++row) a.gleaf[(int64_t)row * a.npad_p + n] = 0.f;
            for (int row = 0; row < Q * (2 * G + S); ++row) a.aux_reg[(int64_t)row * a.npad_p + n] = 0.f;
            for (int row = 0; row < R * (1 + 2 * S); ++row) a.aux_root[(int64_t)row * a.npad_p + n] = 0.f;
        }
        for (int64_t n = a.F + tid; n < a.npad_f; n += blockDim.x) {
            for (int row = 0; row < RB * 2 * GB; ++row) a.bgleaf[(int64_t)row * a.npad_f + n] = 0.f;
            for (int row = 0; row < RB * (1 + 2 * GB); ++row) a.baux_root[(int64_t)row * a.npad_f + n] = 0.f;
        }
    }
    for (int fr0 = 0; fr0 < cnt; fr0 += a.rf) {
        const int nfr = min(a.rf, cnt - fr0);
        const int64_t fbase = f0 + fr0;
        const int npatch = nfr * a.O, ntile = (npatch + HT - 1) / HT;
        const int64_t nbase = fbase * a.O;
        // ---- sum / root weights -> U
        {
            const float4* src = reinterpret_cast<const float4*>(a.wlin);
            float4* dst = reinterpret_cast<float4*>(smem + m.wl);
            for (int i = tid; i < Q * G * G * SP / 4; i += blockDim.x) cp_async16(dst + i, src + i);
            for (int i = tid; i < R * S * S; i += blockDim.x) cp_async4(smem + m.rws + i, a.rlin + i);
            cp_async_commit();
            for (int i = tid; i < R * S * S; i += blockDim.x) {
                const int r = i / (S * S), k = i - r * S * S, x = k / S, y = k - x * S;     // rwsT[r][x][y] = w[r][y][x]
                smem[m.rwsT + i] = __ldg(a.rlin + r * S * S + y * S + x);
            }
            // background root while the weights are in flight: global memory -> bgl + workspace, independent of the rest
            for (int fi = warp; fi < nfr; fi += nw) bwd_bg_root_frame<RB, GB>(a, m, smem, fi, fbase + fi, lane);
            cp_async_wait<0>();
        }
        __syncthreads();
        // ---- N1, N2
        for (int t = warp; t < R * ntile; t += nw) {
            const int tile = t / R, r = t - tile * R;
            bwd_root_task<S>(a, m, smem, r, tile, min(HT, npatch - tile * HT), nbase + tile * HT, lane);
        }
        __syncthreads();
        for (int t = warp; t < Q * ntile; t += nw) {
            const int tile = t / Q, q = t - tile * Q;
            bwd_region_task<G, S>(a, m, smem, q, tile, min(HT, npatch - tile * HT), nbase + tile * HT, lane);
        }
        __syncthreads();
        // ---- leaf table + slot table -> U, (x, mask) tile -> T
        {
            stage_leaf_poly(smem + m.lf, a.leaf, Q * a.st.pmax, GP, tid, blockDim.x);      // (c1, c2, c0) rows, see scene_ll.cuh
            int32_t* sl = reinterpret_cast<int32_t*>(smem + m.slots);
            for (int i = tid; i < D * R; i += blockDim.x) {
                const int slot = __ldg(a.st.slot + i);
                const int q = slot / a.st.pmax, p = slot - q * a.st.pmax;
                sl[i] = slot | ((q * 2 + (p >= __ldg(a.st.n0 + q) ? 1 : 0)) << 16);
            }
            float2* xw = reinterpret_cast<float2*>(smem + m.t);
            const float* xs = a.patches + nbase * D;
            const float* ms = a.marg_patch + nbase * D;
            if ((D & 3) == 0) {
                // warp = patch, lane = 4 pixels: two 16-byte loads, four tile entries
                for (int p = warp; p < npatch; p += nw) {
                    const int tile = p / HT, pt = p - tile * HT;
                    float2* xt = xw + (size_t)tile * D * HT;
                    for (int l = lane; l < D / 4; l += 32) {
                        const float4 x4 = __ldg(reinterpret_cast<const float4*>(xs + (size_t)p * D) + l);
                        const float4 m4 = __ldg(reinterpret_cast<const float4*>(ms + (size_t)p * D) + l);
                        const int px = 4 * l;
                        xt[px * HT + ((pt + px) & (HT - 1))] = make_float2(x4.x, m4.x);
                        xt[(px + 1) * HT + ((pt + px + 1) & (HT - 1))] = make_float2(x4.y, m4.y);
                        xt[(px + 2) * HT + ((pt + px + 2) & (HT - 1))] = make_float2(x4.z, m4.z);
                        xt[(px + 3) * HT + ((pt + px + 3) & (HT - 1))] = make_float2(x4.w, m4.w);
                    }
                }
            } else {
                for (int i = tid; i < npatch * D; i += blockDim.x) {
                    const int p = i / D, px = i - p * D, tile = p / HT, pt = p - tile * HT;
                    xw[(size_t)tile * D * HT + px * HT + ((pt + px) & (HT - 1))] = make_float2(__ldg(xs + i), __ldg(ms + i));
                }
            }
        }
        __syncthreads();
        // ---- IN
        {
            const int npp = (D + 1) / 2;
            for (int t = warp; t < ntile * npp; t += nw) {
                const int tile = t / npp, pp = t - tile * npp;
                bwd_input_task<G>(a, m, smem, tile, pp, lane);
            }
        }
        __syncthreads();
        // ---- frames -> U as (x, 0): the frames of a round are contiguous in global memory; 8 loads in flight per thread
        if ((AB & 3) == 0) {
            const float4* src = reinterpret_cast<const float4*>(a.img + fbase * AB);
            const int q4 = AB / 4, tot = nfr * q4;
            for (int i0 = 0; i0 < tot; i0 += 8 * (int)blockDim.x) {
                float4 t[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = i0 + k * (int)blockDim.x + tid;
                    t[k] = i < tot ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = i0 + k * (int)blockDim.x + tid;
                    if (i < tot) {
                        const int fi = i / q4, j = i - fi * q4;
                        float4* d4 = reinterpret_cast<float4*>(reinterpret_cast<float2*>(smem + m.fb) + (size_t)fi * a.fs + 4 * j);
                        d4[0] = make_float4(t[k].x, 0.f, t[k].y, 0.f);
                        d4[1] = make_float4(t[k].z, 0.f, t[k].w, 0.f);
                    }
                }
            }
        } else {
            for (int fi = 0; fi < nfr; ++fi) {
                float2* dst = reinterpret_cast<float2*>(smem + m.fb) + (size_t)fi * a.fs;
                const float* src = a.img + (fbase + fi) * AB;
                for (int i = tid; i < AB; i += blockDim.x) dst[i] = make_float2(__ldg(src + i), 0.f);
            }
        }
        __syncthreads();
        // ---- BI
        {
            // every CTA walks the whole leaf table: start at a CTA-dependent chunk so that the 148 SMs do not ask the
            // same L2 slices for the same sectors at the same moment
            const int nch = (a.Dbg + 31) / 32, rot = (int)((blockIdx.x * 5u) % (unsigned)nch);
            for (int t = warp; t < nch; t += nw) {
                const int ch = t + rot;
                bwd_bg_input_task<RB, GB>(a, m, smem, ch < nch ? ch : ch - nch, nfr, lane);
            }
        }
        __syncthreads();
        // ---- SC
        for (int fi = warp; fi < nfr; fi += nw) scene_frame_bwd(a, m, smem, fi, fbase + fi, lane);
        __syncthreads();
    }
}

}  // namespace sl

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
extern "C" int stove_scene_ll_bwd(int64_t F, int O, int A, int B, int pa, int pb, int align_corners, const float* img,
                                  const float* z, const stove_spn2_struct* obj, const float* leaf, const float* wlin,
                                  const float* wlog, const float* rlin, const float* rlog,
                                  const stove_spn1_struct* bg, const int32_t* bg_scope, const int32_t* bg_cnt,
                                  const float* bleaf, const float* brlin, const float* brlog, const float* bleaf_il,
                                  int il_stride, const float* patches,
                                  const float* marg_patch, const float* marg_bg, const float* leaf_val,
                                  const float* sum_val, const float* out_obj, const float* bleaf_val,
                                  const float* out_bg, const float* g_obj, const float* g_bg, const float* g_overlap,
                                  float* g_z, float* g_leaf, float* g_wlog, float* g_rlog, float* g_bleaf,
                                  float* g_brlog, void* ws_obj, void* ws_bg, const stove_scene_seq* seq, void* stream,
                                  void* join_obj, void* join_bg) {
    sl::LLArgs a{};
    STOVE_CHECK_ARG(obj && bg, "null structure");
    if (seq) {
        STOVE_CHECK_ARG(!z && !g_obj && !g_bg && !g_overlap && !g_z, "sequence mode: z and the per-term gradients must be NULL");
        STOVE_CHECK_ARG(seq->n > 0 && seq->T > seq->skip && seq->skip >= 1 && seq->Z >= 4 && seq->z_sup && seq->z_s && seq->g_elbo &&
                            seq->g_z_sup && seq->g_z_s && seq->g_logq && seq->g_trans && F == seq->n * (seq->T - 1),
                        "bad sequence block");
        STOVE_CHECK_ARG(((uintptr_t)seq->g_z_sup & 15) == 0, "g_z_sup must be 16-byte aligned");
        STOVE_CHECK_ARG(F * O < (1ll << 31), "too many patches for the sequence mode");
        a.seq = 1;
        a.sq = *seq;
        a.w_elbo = (float)(1.0 / ((double)seq->n * (seq->T - seq->skip)));
        a.w_sup = seq->skip > 1 ? (float)(1.0 / ((double)seq->n * (seq->skip - 1))) : 0.f;
    } else {
        STOVE_CHECK_ARG(z && g_obj && g_bg && g_z, "null pointer");
        STOVE_CHECK_ARG(((uintptr_t)z & 15) == 0 && ((uintptr_t)g_z & 15) == 0, "z / g_z must be 16-byte aligned");
    }
    STOVE_CHECK_ARG(F >= 0 && O > 0 && A > 0 && B > 0 && pa > 0 && pb > 0 && img && bg_scope && bg_cnt, "bad argument");
    STOVE_CHECK_ARG(leaf && wlin && wlog && rlin && rlog && bleaf && brlin && brlog && patches && marg_patch && marg_bg &&
                        leaf_val && sum_val && out_obj && bleaf_val && out_bg && g_leaf &&
                        g_wlog && g_rlog && g_bleaf && g_brlog && ws_obj && ws_bg, "null pointer");
    STOVE_CHECK_ARG(bleaf_il && il_stride > 0 && (il_stride & 31) == 0 && ((uintptr_t)bleaf_il & 15) == 0, "bad interleaved table");
    if (F == 0) return STOVE_OK;
    a.bleaf_il = bleaf_il; a.il_stride = il_stride;
    a.O = O; a.A = A; a.B = B; a.pa = pa; a.pb = pb; a.align = align_corners; a.F = F;
    a.img = img; a.z = z;
    a.st.D = obj->D; a.st.R = obj->R; a.st.pmax = obj->pmax;
    a.st.scope = obj->region_scope; a.st.n0 = obj->region_n0; a.st.nt = obj->region_n; a.st.slot = obj->pix_slot;
    a.Np = F * O;
    a.npad_p = round_up64(a.Np, 32);
    a.npad_f = round_up64(F, 32);
    a.Dbg = bg->D; a.bg_side = bg->side; a.bg_scope = bg_scope; a.bg_cnt = bg_cnt;
    a.leaf = leaf; a.wlin = wlin; a.wlog = wlog; a.rlin = rlin; a.rlog = rlog;
    a.bleaf = bleaf; a.brlin = brlin; a.brlog = brlog;
    a.patches = const_cast<float*>(patches); a.marg_patch = const_cast<float*>(marg_patch);
    a.marg_bg = const_cast<float*>(marg_bg);
    a.leaf_val = const_cast<float*>(leaf_val); a.sum_val = const_cast<float*>(sum_val);
    a.out_obj = const_cast<float*>(out_obj); a.bleaf_val = const_cast<float*>(bleaf_val);
    a.out_bg = const_cast<float*>(out_bg);
    a.g_obj = g_obj; a.g_bg = g_bg; a.g_overlap = g_overlap; a.g_z = g_z;
    a.g_wlog = g_wlog; a.g_rlog = g_rlog; a.g_brlog = g_brlog;
    {
        // workspace layouts of spn_obj.cu (spn2_ws_layout) and spn_bg.cu (spn1_ws_layout)
        const int Q = 2 * obj->R, G = obj->G, S = obj->S;
        float* p = (float*)ws_obj;
        a.gleaf = p; p += (int64_t)Q * 2 * G * a.npad_p;
        a.aux_reg = p; p += (int64_t)Q * (2 * G + S) * a.npad_p;
        a.aux_root = p;
        float* b = (float*)ws_bg;
        a.bgleaf = b; b += (int64_t)bg->R * 2 * bg->G * a.npad_f;
        a.baux_root = b;
    }
    int grid;
    size_t smem;
    if (bg->side == nullptr || !sl_plan(a, obj->G, obj->S, bg->R, bg->G, &grid, &smem, true)) {
        stove_set_error("stove_scene_ll_bwd: configuration not supported by the fused kernel (see stove_scene_ll_supported)");
        return STOVE_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    auto kernel = sl::scene_ll_bwd_kernel<10, 10, 3, 6>;
    STOVE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    STOVE_KERNEL(K_SCENE_LL_BWD, s, kernel<<<grid, a.nw * 32, smem, s>>>(a));
    STOVE_LAUNCH_CHECK();
    // parameter gradients: off the chain, on library side streams, joined where the packing's backward runs
    int rc;
    const int64_t Np = F * O;
    StoveFork* fk0 = !stove_opt(OPT_FORK) ? nullptr : stove_fork_get(0);
    if (fk0 && (rc = stove_fork(fk0, s, 2))) return rc;
    if ((rc = spn2_param_kernels(obj, Np, patches, marg_patch, leaf, wlin, rlin, ws_obj, g_leaf, g_wlog, g_rlog,
                                 fk0 ? fk0->side[0] : s, fk0 ? fk0->side[1] : s)))
        return rc;
    if (fk0 && (rc = stove_join(fk0, join_obj ? (cudaStream_t)join_obj : s, 2))) return rc;
    StoveFork* fk1 = !stove_opt(OPT_FORK) ? nullptr : stove_fork_get(1);
    if (fk1 && (rc = stove_fork(fk1, s, 2))) return rc;
    if ((rc = spn1_param_kernels(bg, F, img, marg_bg, bleaf, brlin, ws_bg, g_bleaf, g_brlog, fk1 ? fk1->side[0] : s,
                                 fk1 ? fk1->side[1] : s)))
        return rc;
    if (fk1 && (rc = stove_join(fk1, join_bg ? (cudaStream_t)join_bg : s, 2))) return rc;
    return STOVE_OK;
}
