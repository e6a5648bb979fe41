// Fused scene likelihood: glimpses + sequential compositing masks + object RAT-SPN + background RAT-SPN in ONE
// forward launch (and one backward chain launch, scene_ll_bwd.cu), fp32, sm_100a.
//
// Replaces the op sequence of Supair.likelihood (model/video_prediction/supair.py:62-76):
//   masks_from_z (:278-356) -> bg_spn.forward (rat_torch.py:333-357, structure probabilistic_models.py:25-39)
//   patches_from_z (:241-276) -> obj_spn.forward (structure probabilistic_models.py:8-22)
// The unfused kernels (scene.cu -> spn_obj.cu / spn_bg.cu) hand 2 x (F*O, 100) glimpse / mask rows and the
// (F, W*H) background mask from one launch to the next through L2.  Here a CTA owns a contiguous group of
// frames (F / #SM of them, 12-13 at batch 256); their glimpses and masks are produced straight into the
// object SPN's shared-memory tile layout and consumed by the leaf pass of the same CTA, the final background
// mask of every frame stays in shared memory for the background SPN's leaf pass.
//
// Phases of a CTA (a "round" = up to MAXF frames; one round at the benchmark sizes):
//   S    warp = frame: frame + running background in shared memory, objects in order: bilinear glimpse and
//        marginalisation mask -> tile xw[pixel][patch] = (x, 1 - mask); paste of the box = tent(y) * tent(x);
//        clamp.  (scene.cu runs a 256-thread CTA per frame with 3 CTA-wide barriers per object and was issue bound:
//        11 k warp instructions per frame, profiles/r01_ncu_full_spn_v3.txt.)
//   BG   warp = (background leaf, slice of its scope), lane = pixel of the scope: the leaf parameters of a pixel
//        are read ONCE from L2 and applied to all frames of the round (accumulators [frame][gaussian] in registers),
//        then a transposed butterfly reduces them over the lanes; slices are combined in a fixed order.
//   --   background root per frame while the object SPN's tables are staged (cp.async) over the frame buffers.
//   OBJ  warp = (mid region q, half-warp tile of 16 patches), lane = (leaf h of the region, patch): each lane
//        accumulates ONE leaf vector (25 pixels x 10 Gaussians), the two halves exchange exp-shifted vectors by
//        shuffle and each computes half of the region's sums (max-shifted linear domain, exact log-domain slow path).
//        Half-warp tiles keep 81-94 % of the lanes busy for the 36-39 patches a CTA owns; the 32-patch tiles of
//        spn_obj.cu would leave 40 % idle and cost the same issue slots.
//        The leaf table is staged in polynomial form (scene_ll.cuh: stage_leaf_poly): three FMAs per (pixel, Gaussian).
//   root partitions per (r, tile), final logsumexp per patch.
// Sequence mode (stove_scene_seq): the states come straight from the sequence tensors z_sup / z_s of
// Stove.stove_forward (stove.py:731-736) and the object terms also leave weighted by sx * sy (supair.py:79).
// Everything the backward pass needs is written once (leaf / sum values, root values) in the layouts of
// spn_obj.cu / spn_bg.cu, so the unfused kernels remain usable on the fused forward's outputs (tests do that).
#include "common.cuh"
#include "scene_math.cuh"
#include "spn_math.cuh"
#include "scene_ll.cuh"

namespace sl {

// ------------------------------------------------------------------------------------
// phase S: one warp composites one frame
// ------------------------------------------------------------------------------------
// The frame buffer holds (x, running background mask) pairs: one 8-byte load per bilinear corner serves the glimpse
// and the mask, and the background leaf pass reads both with one load as well.
__device__ __forceinline__ void bilinear_pair(const float2* fb, int B, const Corner& c, float& xval, float& mval) {
    float2 v00 = make_float2(0.f, 0.f), v01 = v00, v10 = v00, v11 = v00;       // (x, 1 - background), zero padded
    if (c.oky0 && c.okx0) { const float2 t = fb[c.y0 * B + c.x0]; v00 = make_float2(t.x, 1.f - t.y); }
    if (c.oky0 && c.okx1) { const float2 t = fb[c.y0 * B + c.x0 + 1]; v01 = make_float2(t.x, 1.f - t.y); }
    if (c.oky1 && c.okx0) { const float2 t = fb[(c.y0 + 1) * B + c.x0]; v10 = make_float2(t.x, 1.f - t.y); }
    if (c.oky1 && c.okx1) { const float2 t = fb[(c.y0 + 1) * B + c.x0 + 1]; v11 = make_float2(t.x, 1.f - t.y); }
    const float tx_ = v00.x + c.fx * (v01.x - v00.x), bx_ = v10.x + c.fx * (v11.x - v10.x);
    xval = tx_ + c.fy * (bx_ - tx_);
    const float tm_ = v00.y + c.fx * (v01.y - v00.y), bm_ = v10.y + c.fx * (v11.y - v10.y);
    mval = tm_ + c.fy * (bm_ - tm_);
}

__device__ void scene_frame_fwd(const LLArgs& a, int64_t f, int pl0, float2* fb, float* tX, float* tY, float2* xw,
                                int lane) {
    const int AB = a.A * a.B, PP = a.pa * a.pb, D = a.st.D;
    const float* src = a.img + f * AB;
    if ((AB & 3) == 0) {
        // all loads of the frame in flight before the first store (the scalar loop paid one L2 round trip per 4 pixels)
        const float4* s4 = reinterpret_cast<const float4*>(src);
        for (int i0 = 0; i0 < AB / 4; i0 += 32 * 8) {
            float4 t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = i0 + 32 * k + lane;
                t[k] = (i < AB / 4) ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = i0 + 32 * k + lane;
                if (i < AB / 4) {
                    float4* d4 = reinterpret_cast<float4*>(fb + 4 * i);
                    d4[0] = make_float4(t[k].x, 0.f, t[k].y, 0.f);
                    d4[1] = make_float4(t[k].z, 0.f, t[k].w, 0.f);
                }
            }
        }
    } else {
        for (int i = lane; i < AB; i += 32) fb[i] = make_float2(__ldg(src + i), 0.f);
    }
    if (lane == 0) fb[AB] = make_float2(0.f, 1.f);        // null pixel: weight 1 - mask = 0 (idle lanes of the background pass)
    // normalised glimpse coordinates of this lane's pixels (the same for every object)
    float xb[MAXIT], yb[MAXIT];
    {
        const float rb = recip_n(a.pb, a.align), ra = recip_n(a.pa, a.align);
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            const int idx = lane + 32 * it, i = idx / a.pb, j = idx - i * a.pb;
            xb[it] = base_coord_r(j, rb, a.align);
            yb[it] = base_coord_r(i, ra, a.align);
        }
    }
    const float kB = unnorm_slope(a.B, a.align), kA = unnorm_slope(a.A, a.align);
    const float oB = unnorm_offset(a.B, a.align), oA = unnorm_offset(a.A, a.align);
    const float rPP = 1.f / (float)PP;
    const int du = 32 / a.B, dv = 32 - du * a.B;
    __syncwarp();
    float4 znext = load_z(a, f, 0);
    for (int o = 0; o < a.O; ++o) {
        const float4 zz = znext;
        if (o + 1 < a.O) znext = load_z(a, f, o + 1);        // the next object's state travels while this one is composited
        const float sx = zz.x, sy = zz.y, tx = zz.z, ty = zz.w;
        const float mx = sx * kB, ox = fmaf(tx, kB, oB), my = sy * kA, oy = fmaf(ty, kA, oA);
        const int pl = pl0 + o, tile = pl / HT, pt = pl - tile * HT;
        float2* xt = xw + (size_t)tile * D * HT;
        const int64_t nb = (f * a.O + o) * PP;
        float msum = 0.f;
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            const int idx = lane + 32 * it;
            if (idx < PP) {
                const Corner c = corners(fmaf(yb[it], my, oy), fmaf(xb[it], mx, ox), a.A, a.B);
                float val, inv;
                bilinear_pair(fb, a.B, c, val, inv);
                const float mg = 1.f - inv;
                msum += mg;
                a.patches[nb + idx] = val;
                a.marg_patch[nb + idx] = mg;
                // (x, weight = 1 - clamp(mask)); the column is rotated by the pixel index so that the 32 pixel rows a
                // warp writes here fall into 16 different bank pairs (the readers rotate back)
                xt[idx * HT + ((pt + idx) & (HT - 1))] = make_float2(val, 1.f - fminf(fmaxf(mg, 0.f), 1.f));
            }
        }
        int ulo, uhi;
        warp_tents(a, sx, sy, tx, ty, tX, tY, nullptr, nullptr, lane, ulo, uhi);
        msum = warp_sum(msum);
        if (lane == 0) a.overlap[f * a.O + o] = msum * rPP;
        __syncwarp();
        if (uhi >= ulo) {
            int u = ulo + lane / a.B, v = lane % a.B;
            for (int i = ulo * a.B + lane; i < (uhi + 1) * a.B; i += 32) {
                fb[i].y = fminf(fmaxf(fb[i].y + tY[u] * tX[v], 0.f), 1.f);
                u += du;
                v += dv;
                if (v >= a.B) { v -= a.B; ++u; }
            }
        }
        __syncwarp();
    }
    float* dst = a.marg_bg + f * AB;
    for (int i = lane; i < AB; i += 32) dst[i] = fb[i].y;
}

// ------------------------------------------------------------------------------------
// phase BG: background leaf pass over the frames of a round
// ------------------------------------------------------------------------------------
// v[k], k < N / 32, becomes the warp-wide sum of the original v[(N / 32) * lane + k]: a butterfly in which every
// stage halves the number of values a lane carries (2 N instructions in total instead of 10 N for N warp_sums)
template <int N, int OFF>
struct ReduceScatter {
    static __device__ __forceinline__ void run(float* v, int lane) {
        const bool up = (lane & OFF) != 0;
#pragma unroll
        for (int k = 0; k < N / 2; ++k) {
            const float lo = v[k], hi = v[k + N / 2];
            const float send = up ? lo : hi, keep = up ? hi : lo;
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        if constexpr (OFF > 1) ReduceScatter<N / 2, OFF / 2>::run(v, lane);
    }
};

// G0 .. G0 + NG - 1 = the Gaussians of this pass (NG <= 4, G0 a multiple of 4: one float4 per parameter array)
template <int RB, int GB, int G0, int NG>
__device__ __forceinline__ void bg_leaf_pass(const LLArgs& a, int task, int l, int r, int i0, int i1, int nfr,
                                             const float2* fb, float* bgpart, int lane) {
    constexpr int GPB = GP_<GB>::v;
    constexpr int NV = (MAXF * NG + 31) / 32 * 32;
    const int32_t* sc = a.bg_scope + (size_t)l * a.Dbg;
    float acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.f;
    const unsigned fb_s = (unsigned)__cvta_generic_to_shared(fb), fstride = (unsigned)a.fs * 8u;
    // (the iterations of a slice start at a CTA-dependent position and wrap around: every CTA walks the whole leaf
    // table, and 148 SMs asking the same L2 slices for the same sectors at the same moment serialise there)
    const int nit = (i1 - i0 + 31) >> 5;
    int itr = nit > 0 ? (int)(blockIdx.x % (unsigned)nit) : 0;
    for (int k = 0; k < nit; ++k) {
        const int ib = i0 + 32 * itr;
        itr = (itr + 1 == nit) ? 0 : itr + 1;
        const int i = ib + lane;
        const bool ok = i < i1;
        // lanes past the end of the slice read the frame's null pixel (x = 0, mask = 1: weight 0)
        const int px = ok ? __ldg(sc + i) : a.Dbg;
        // interleaved table: block of 32 rows = [6 parts][32 lanes] float4, rows of leaf l start at l * il_stride;
        // rows past the end of the scope are zero (slices start at multiples of 32)
        const float4* p4 = reinterpret_cast<const float4*>(a.bleaf_il) + ((int64_t)(l * a.il_stride + ib) >> 5) * (6 * 32) + lane;
        const float4 m4 = __ldg(p4 + (G0 / 4) * 32), a4 = __ldg(p4 + (GPB / 4 + G0 / 4) * 32),
                     b4 = __ldg(p4 + (2 * (GPB / 4) + G0 / 4) * 32);
        const float mu[4] = {m4.x, m4.y, m4.z, m4.w}, aa[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
        unsigned addr = fb_s + (unsigned)px * 8u;
        // one address increment per frame instead of index arithmetic on the runtime frame stride (the first version
        // spent a third of this phase's instructions on it and on selecting the weight of idle lanes)
#pragma unroll
        for (int f = 0; f < MAXF; ++f) {
            if (f < nfr) {
                float2 v;                                                 // (x, final background mask in [0, 1])
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
                addr += fstride;
                const float wv = 1.f - v.y;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const float d = v.x - mu[g];
                    acc[f * NG + g] = fmaf(-wv, fmaf(d * d, aa[g], bb[g]), acc[f * NG + g]);
                }
            }
        }
    }
    ReduceScatter<NV, 16>::run(acc, lane);
    constexpr int K = NV / 32;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int idx = K * lane + k;                                    // = f * NG + g
        if (idx < MAXF * NG) bgpart[task * (MAXF * GB) + (idx / NG) * GB + G0 + idx % NG] = acc[k];
    }
}

template <int RB, int GB>
__device__ void bg_leaf_task(const LLArgs& a, int task, int nfr, const float2* fb, float* bgpart, int lane) {
    static_assert(GB > 4 && GB <= 8 && GP_<GB>::v == 8, "two passes: Gaussians 0-3, then 4 .. GB-1; 8-float parameter rows");
    const int l = task / a.ns, s = task - l * a.ns, r = l >> 1;
    const int cnt = __ldg(a.bg_cnt + l);
    const int per = ((cnt + a.ns - 1) / a.ns + 31) & ~31;            // slices start at multiples of 32 (interleaved table)
    const int i0 = min(cnt, s * per), i1 = min(cnt, i0 + per);
    // 13 frames x 6 Gaussians of accumulators + 18 parameters do not fit the 96 registers a 576-thread CTA gets
    // (the first version spilled inside the loop: profiles/r02_ncu_scene_ll_fwd_v1.txt)
    bg_leaf_pass<RB, GB, 0, 4>(a, task, l, r, i0, i1, nfr, fb, bgpart, lane);
    bg_leaf_pass<RB, GB, 4, GB - 4>(a, task, l, r, i0, i1, nfr, fb, bgpart, lane);
}

// products + root sum of the background SPN for one frame (rat_torch.py:147-163, 202-222): warp = frame, lane =
// (root partition r, index j within an 8-lane group).  One thread per frame took 5.6 k instructions in a row while
// 17 warps waited at the barrier behind it (20 % of the first version's time).
template <int RB, int GB>
__device__ void bg_root_frame(const LLArgs& a, int fi, int64_t f, const float* bgpart, int lane) {
    static_assert(GB <= 8 && RB <= 4, "8-lane groups");
    const int r = lane >> 3, j = lane & 7;
    const bool on = r < RB && j < GB;
    float av = -INFINITY, bv = -INFINITY;
    if (on) {
        av = 0.f; bv = 0.f;
        for (int s = 0; s < a.ns; ++s) {                                  // fixed order: deterministic
            av += bgpart[((2 * r) * a.ns + s) * (MAXF * GB) + fi * GB + j];
            bv += bgpart[((2 * r + 1) * a.ns + s) * (MAXF * GB) + fi * GB + j];
        }
        a.bleaf_val[(int64_t)((2 * r) * GB + j) * a.npad_f + f] = av;
        a.bleaf_val[(int64_t)((2 * r + 1) * GB + j) * a.npad_f + f] = bv;
    }
    float mA = av, mB = bv;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, o));
        mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, o));
    }
    const float eA = on ? expf(av - mA) : 0.f, eB = on ? expf(bv - mB) : 0.f;
    float inner = 0.f;                                                    // sum_i eA[i] w[j][i]
#pragma unroll
    for (int i = 0; i < GB; ++i) {
        const float ei = __shfl_sync(0xffffffffu, eA, (lane & ~7) + i);
        if (on) inner = fmaf(ei, __ldg(a.brlin + r * GB * GB + j * GB + i), inner);
    }
    float U = eB * inner;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) U += __shfl_xor_sync(0xffffffffu, U, o);
    float val = mA + mB + logf(U);
    if (!(U > LIN_SUM_FLOOR) && r < RB) {
        // exact log-domain value; every lane of the group walks all pairs (rare path)
        const float* pa_ = bgpart + fi * GB;
        float M = -INFINITY;
        for (int jj = 0; jj < GB; ++jj)
            for (int i = 0; i < GB; ++i) {
                float A_ = 0.f, B_ = 0.f;
                for (int s = 0; s < a.ns; ++s) {
                    A_ += pa_[((2 * r) * a.ns + s) * (MAXF * GB) + i];
                    B_ += pa_[((2 * r + 1) * a.ns + s) * (MAXF * GB) + jj];
                }
                M = fmaxf(M, A_ + B_ + __ldg(a.brlog + r * GB * GB + jj * GB + i));
            }
        float acc = 0.f;
        if (M > -INFINITY)
            for (int jj = 0; jj < GB; ++jj)
                for (int i = 0; i < GB; ++i) {
                    float A_ = 0.f, B_ = 0.f;
                    for (int s = 0; s < a.ns; ++s) {
                        A_ += pa_[((2 * r) * a.ns + s) * (MAXF * GB) + i];
                        B_ += pa_[((2 * r + 1) * a.ns + s) * (MAXF * GB) + jj];
                    }
                    acc += expf(A_ + B_ + __ldg(a.brlog + r * GB * GB + jj * GB + i) - M);
                }
        val = (M > -INFINITY) ? M + logf(acc) : M;
    }
    // logsumexp over the root partitions (lanes 0, 8, 16, ...)
    float M = -INFINITY;
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) M = fmaxf(M, __shfl_sync(0xffffffffu, val, 8 * rr));
    float acc = 0.f;
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) acc += expf(__shfl_sync(0xffffffffu, val, 8 * rr) - M);
    if (lane == 0) a.out_bg[f] = (M > -INFINITY) ? M + logf(acc) : M;
}

// ------------------------------------------------------------------------------------
// phase OBJ
// ------------------------------------------------------------------------------------
template <int G, int S>
__device__ void obj_region_task(const LLArgs& a, const Smem& m, float* smem, int q, int tile, int npt, int64_t n0g,
                                int lane) {
    constexpr int GP = GP_<G>::v, SP = GP_<S>::v, SH = 5;
    static_assert(S > 8 && S <= 10 && SP == 12, "half-warp split of the sums is laid out for 9-10 sums");
    const int h = lane >> 4, pt = lane & (HT - 1);
    const int D = a.st.D, Q = 2 * a.st.R;
    const int nq0 = __ldg(a.st.n0 + q), nq = __ldg(a.st.nt + q);
    const int pbeg = h ? nq0 : 0, pcnt = h ? nq - nq0 : nq0;
    const int plen = max(nq0, nq - nq0);
    const float2* xt = reinterpret_cast<const float2*>(smem + m.xw) + (size_t)tile * D * HT;
    const float* leaf_q = smem + m.lf + q * a.st.pmax * 3 * GP;
    const int32_t* sc = reinterpret_cast<const int32_t*>(smem + m.scs) + q * a.st.pmax;
    float L[G];
#pragma unroll
    for (int g = 0; g < G; ++g) L[g] = 0.f;
#pragma unroll 2
    for (int p = 0; p < plen; ++p) {
        const bool ok = p < pcnt;
        const int pp = pbeg + (ok ? p : 0);
        const int px = sc[pp];
        const float2 v = xt[px * HT + ((pt + px) & (HT - 1))];
        const float wv = ok ? v.y : 0.f;
        const float wx = wv * v.x, wx2 = wx * v.x;
        float c1[GP], c2[GP], c0[GP];
        load_leaf_params<G, true>(leaf_q + pp * 3 * GP, c1, c2, c0);
#pragma unroll
        for (int g = 0; g < G; ++g) {            // L -= w (c2 x^2 - c1 x + c0) = w (a (x - mu)^2 + b)
            L[g] = fmaf(-wx2, c2[g], L[g]);
            L[g] = fmaf(wx, c1[g], L[g]);
            L[g] = fmaf(-wv, c0[g], L[g]);
        }
    }
    const bool live = pt < npt;
    const int64_t n = n0g + pt;
    float* lv = a.leaf_val + (int64_t)(q * 2) * G * a.npad_p + n;       // leaf 0 of the region; leaf 1 is G rows further
    if (live) {
#pragma unroll
        for (int g = 0; g < G; ++g) lv[(int64_t)(h * G + g) * a.npad_p] = L[g];
    }
    float e[G], ep[G];
    const float mo = shift_exp<G>(L, e);
    const float mp = __shfl_xor_sync(0xffffffffu, mo, 16);
#pragma unroll
    for (int g = 0; g < G; ++g) ep[g] = __shfl_xor_sync(0xffffffffu, e[g], 16);
    __syncwarp();                       // the slow path below reads both leaf vectors back from global memory
    float e0[G], e1[G];                 // exp-shifted vectors of leaf 0 / leaf 1 of the region
#pragma unroll
    for (int g = 0; g < G; ++g) {
        e0[g] = h ? ep[g] : e[g];
        e1[g] = h ? e[g] : ep[g];
    }
    float T[SH];
#pragma unroll
    for (int c = 0; c < SH; ++c) T[c] = 0.f;
    const float* wq = smem + m.wl + q * G * G * SP + 4 * h;
    // j as a real loop (e1[j] picked by a select chain): fully unrolled, the 100 products were 900 straight-line
    // instructions that a warp runs twice -- a quarter of this phase's stall samples were instruction fetch
#pragma unroll 1
    for (int j = 0; j < G; ++j) {
        float e1j = e1[0];
#pragma unroll
        for (int t = 1; t < G; ++t) e1j = (j == t) ? e1[t] : e1j;
#pragma unroll
        for (int i = 0; i < G; ++i) {
            const float pk = e0[i] * e1j;
            const float* wk = wq + (j * G + i) * SP;
            const float4 w = lds_f4(wk);
            const float w4 = lds_f1(wk + 8 - 3 * h);                // element 8 + h of the row
            T[0] = fmaf(pk, w.x, T[0]);
            T[1] = fmaf(pk, w.y, T[1]);
            T[2] = fmaf(pk, w.z, T[2]);
            T[3] = fmaf(pk, w.w, T[3]);
            T[4] = fmaf(pk, w4, T[4]);
        }
    }
    float* ss = smem + m.ss + (size_t)tile * Q * S * HT;
#pragma unroll
    for (int c = 0; c < SH; ++c) {
        const int s = sum_of(h, c);
        if (s < S) {
            float val = mo + mp + logf(T[c]);
            if (!(T[c] > LIN_SUM_FLOOR))
                val = live ? slow_logsumexp(lv, lv + (int64_t)G * a.npad_p, (int)a.npad_p, G,
                                            a.wlog + (int64_t)q * G * G * SP + s, SP)
                           : 0.f;
            ss[(q * S + s) * HT + pt] = val;
            if (live) a.sum_val[(int64_t)(q * S + s) * a.npad_p + n] = val;
        }
    }
}

template <int S>
__device__ void obj_root_task(const LLArgs& a, const Smem& m, float* smem, int r, int tile, int npt, int lane) {
    const int h = lane >> 4, pt = lane & (HT - 1);
    const int R = a.st.R, Q = 2 * R;
    const float* ss = smem + m.ss + (size_t)tile * Q * S * HT;
    float A_[S], B_[S], eA[S], eB[S];
#pragma unroll
    for (int i = 0; i < S; ++i) {
        A_[i] = ss[((2 * r) * S + i) * HT + pt];
        B_[i] = ss[((2 * r + 1) * S + i) * HT + pt];
    }
    const float mA = shift_exp<S>(A_, eA), mB = shift_exp<S>(B_, eB);
    const float* rw = smem + m.rws + r * S * S;
    constexpr int JH = (S + 1) / 2;
    float U = 0.f;
#pragma unroll
    for (int jj = 0; jj < JH; ++jj) {
        const int j = h * JH + jj;
        if (j < S) {
            float inner = 0.f;
#pragma unroll
            for (int i = 0; i < S; ++i) inner = fmaf(eA[i], lds_f1(rw + j * S + i), inner);
            U = fmaf(h ? eB[(JH + jj < S) ? JH + jj : S - 1] : eB[jj], inner, U);
        }
    }
    U += __shfl_xor_sync(0xffffffffu, U, 16);
    float val = mA + mB + logf(U);
    if (!(U > LIN_SUM_FLOOR))
        val = (pt < npt) ? slow_logsumexp(ss + (2 * r) * S * HT + pt, ss + (2 * r + 1) * S * HT + pt, HT, S,
                                          a.rlog + r * S * S, 1)
                         : 0.f;
    if (h == 0) smem[m.vr + (tile * R + r) * HT + pt] = val;
}

// ------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------
template <int G, int S, int RB, int GB>
__global__ void __launch_bounds__(MAXNW * 32, 1) scene_ll_fwd_kernel(const __grid_constant__ LLArgs a) {
    constexpr int GP = GP_<G>::v, SP = GP_<S>::v;
    extern __shared__ __align__(16) float smem[];
    const Smem m = smem_layout(a, G, S, GB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int R = a.st.R, Q = 2 * R, D = a.st.D;
    // frames of this CTA
    const int64_t per = a.F / gridDim.x, rem = a.F % gridDim.x;
    const int64_t f0 = blockIdx.x * per + min((int64_t)blockIdx.x, rem);
    const int cnt = (int)(per + (blockIdx.x < rem ? 1 : 0));
    const int tXs = up4(a.B), tYs = up4(a.A);

    // unused tile columns are read by idle lanes: keep them finite
    for (int i = tid; i < a.ntile * D * HT * 2; i += blockDim.x) smem[m.xw + i] = 0.f;
    if (blockIdx.x == gridDim.x - 1) {
        // columns [N, npad) of the saved activations belong to nobody; the unfused backward kernels load whole tiles
        for (int64_t n = a.Np + tid; n < a.npad_p; n += blockDim.x) {
            for (int row = 0; row < Q * 2 * G; ++row) a.leaf_val[(int64_t)row * a.npad_p + n] = 0.f;
            for (int row = 0; row < Q * S; ++row) a.sum_val[(int64_t)row * a.npad_p + n] = 0.f;
        }
        for (int64_t n = a.F + tid; n < a.npad_f; n += blockDim.x)
            for (int row = 0; row < RB * 2 * GB; ++row) a.bleaf_val[(int64_t)row * a.npad_f + n] = 0.f;
    }
    __syncthreads();

    for (int fr0 = 0; fr0 < cnt; fr0 += a.rf) {
        const int nfr = min(a.rf, cnt - fr0);
        const int64_t fbase = f0 + fr0;
        // ---- S
        float2* fbuf = reinterpret_cast<float2*>(smem + m.frames_img);
        for (int fi = warp; fi < nfr; fi += nw)
            scene_frame_fwd(a, fbase + fi, fi * a.O, fbuf + (size_t)fi * a.fs, smem + m.frames_tx + fi * tXs,
                            smem + m.frames_ty + fi * tYs, reinterpret_cast<float2*>(smem + m.xw), lane);
        __syncthreads();
        // ---- BG leaf pass
        float* bgpart = smem + m.ss;
        for (int t = warp; t < 2 * RB * a.ns; t += nw)
            bg_leaf_task<RB, GB>(a, t, nfr, fbuf, bgpart, lane);
        __syncthreads();
        // ---- stage the object SPN's tables over the frame buffers; background root meanwhile
        {
            const float4* src = reinterpret_cast<const float4*>(a.wlin);
            float4* dst = reinterpret_cast<float4*>(smem + m.wl);
            for (int i = tid; i < Q * G * G * SP / 4; i += blockDim.x) cp_async16(dst + i, src + i);
            for (int i = tid; i < R * S * S; i += blockDim.x) cp_async4(smem + m.rws + i, a.rlin + i);
            cp_async_commit();
            stage_leaf_poly(smem + m.lf, a.leaf, Q * a.st.pmax, GP, tid, blockDim.x);      // (c1, c2, c0) rows, see scene_ll.cuh
            int32_t* scs = reinterpret_cast<int32_t*>(smem + m.scs);
            for (int i = tid; i < Q * a.st.pmax; i += blockDim.x) scs[i] = max(__ldg(a.st.scope + i), 0);
        }
        for (int fi = warp; fi < nfr; fi += nw) bg_root_frame<RB, GB>(a, fi, fbase + fi, bgpart, lane);
        cp_async_wait<0>();
        __syncthreads();
        // ---- OBJ
        const int npatch = nfr * a.O, ntile = (npatch + HT - 1) / HT;
        const int64_t nbase = fbase * a.O;
        for (int t = warp; t < Q * ntile; t += nw) {
            const int tile = t / Q, q = t - tile * Q;
            obj_region_task<G, S>(a, m, smem, q, tile, min(HT, npatch - tile * HT), nbase + tile * HT, lane);
        }
        __syncthreads();
        for (int t = warp; t < R * ntile; t += nw) {
            const int tile = t / R, r = t - tile * R;
            obj_root_task<S>(a, m, smem, r, tile, min(HT, npatch - tile * HT), lane);
        }
        __syncthreads();
        for (int i = tid; i < npatch; i += blockDim.x) {
            const int tile = i / HT, pt = i - tile * HT;
            const float* vr = smem + m.vr + tile * R * HT + pt;
            float M = vr[0];
            for (int r = 1; r < R; ++r) M = fmaxf(M, vr[r * HT]);
            float acc = 0.f;
            for (int r = 0; r < R; ++r) acc += expf(vr[r * HT] - M);
            const float val = (M > -INFINITY) ? M + logf(acc) : M;
            a.out_obj[nbase + i] = val;
            if (a.seq) {             // the term of the ELBO: weighted by sx * sy (supair.py:79)
                const int64_t f = (nbase + i) / a.O;
                const float4 z = load_z(a, f, (int)(nbase + i - f * a.O));
                a.sq.patch_w[nbase + i] = val * z.x * z.y;
            }
        }
        __syncthreads();               // the next round overwrites the tables with frames
    }
}

}  // namespace sl

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
static int sl_sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// partition + shared-memory plan; returns 0 if the configuration does not fit the fused kernels
int sl_plan(sl::LLArgs& a, int G, int S, int RB, int GB, int* grid, size_t* smem_bytes, bool backward) {
    using namespace sl;
    if (!(G == 10 && S == 10 && RB == 3 && GB == 6)) return 0;
    if (a.F <= 0 || a.O <= 0 || a.O > 16 || a.st.D != a.pa * a.pb || a.Dbg != a.A * a.B || a.pa * a.pb > 32 * MAXIT) return 0;
    const int nsm = sl_sm_count();
    *grid = (int)((a.F < nsm) ? a.F : nsm);
    const int cnt_max = (int)((a.F + *grid - 1) / *grid);
    a.fs = up4(a.A * a.B + 1);          // + the null pixel
    const int Q = 2 * a.st.R;
    for (int rf = cnt_max < MAXF ? cnt_max : MAXF; rf >= 1; --rf) {
        a.rf = rf;
        a.ntile = (rf * a.O + HT - 1) / HT;
        const int tasks = Q * a.ntile;
        const int rounds = (tasks + MAXNW - 1) / MAXNW;
        int nw = (tasks + rounds - 1) / rounds;
        if (nw < rf) nw = rf < MAXNW ? rf : MAXNW;
        if (nw < 4) nw = 4;
        a.nw = nw;
        a.ns = nw / (2 * RB) > 0 ? nw / (2 * RB) : 1;
        const int total = backward ? smem_layout_bwd(a, G, S, GB, RB).total : smem_layout(a, G, S, GB).total;
        if ((size_t)total * sizeof(float) <= 227 * 1024) {
            *smem_bytes = (size_t)total * sizeof(float);
            return 1;
        }
    }
    return 0;
}

static int sl_fill(sl::LLArgs& a, int64_t F, int O, int A, int B, int pa, int pb, int align_corners, const float* img,
                   const float* z, const stove_spn2_struct* obj, const stove_spn1_struct* bg, const int32_t* bg_scope,
                   const int32_t* bg_cnt) {
    STOVE_CHECK_ARG(obj && bg && obj->region_scope && obj->region_n0 && obj->region_n && obj->pix_slot && bg->side,
                    "null structure");
    STOVE_CHECK_ARG(F >= 0 && O > 0 && A > 0 && B > 0 && pa > 0 && pb > 0 && img && z && bg_scope && bg_cnt, "bad argument");
    STOVE_CHECK_ARG(((uintptr_t)z & 15) == 0, "z must be 16-byte aligned");
    a.O = O; a.A = A; a.B = B; a.pa = pa; a.pb = pb; a.align = align_corners; a.F = F;
    a.img = img; a.z = z;
    a.st.D = obj->D; a.st.R = obj->R; a.st.pmax = obj->pmax;
    a.st.scope = obj->region_scope; a.st.n0 = obj->region_n0; a.st.nt = obj->region_n; a.st.slot = obj->pix_slot;
    a.Np = F * O;
    a.npad_p = round_up64(a.Np > 0 ? a.Np : 1, 32);
    a.npad_f = round_up64(F > 0 ? F : 1, 32);
    a.Dbg = bg->D; a.bg_side = bg->side; a.bg_scope = bg_scope; a.bg_cnt = bg_cnt;
    return STOVE_OK;
}

extern "C" int stove_scene_ll_supported(int64_t F, int O, int C, int A, int B, int pa, int pb,
                                        const stove_spn2_struct* obj, const stove_spn1_struct* bg) {
    if (!obj || !bg || C != 1 || F <= 0) return 0;
    sl::LLArgs a{};
    a.O = O; a.A = A; a.B = B; a.pa = pa; a.pb = pb; a.F = F;
    a.st.D = obj->D; a.st.R = obj->R; a.st.pmax = obj->pmax; a.Dbg = bg->D;
    if (obj->R > 8 || B > 32 * SCENE_MAXC) return 0;
    int grid;
    size_t smem;
    sl::LLArgs b = a;
    return sl_plan(a, obj->G, obj->S, bg->R, bg->G, &grid, &smem, false) &&
           sl_plan(b, obj->G, obj->S, bg->R, bg->G, &grid, &smem, true);
}

extern "C" int stove_scene_ll_fwd(int64_t F, int O, int A, int B, int pa, int pb, int align_corners, const float* img,
                                  const float* z, const stove_spn2_struct* obj, const float* leaf, const float* wlin,
                                  const float* wlog, const float* rlin, const float* rlog,
                                  const stove_spn1_struct* bg, const int32_t* bg_scope, const int32_t* bg_cnt,
                                  const float* bleaf, const float* brlin, const float* brlog, const float* bleaf_il,
                                  int il_stride, float* patches,
                                  float* marg_patch, float* marg_bg, float* overlap, float* leaf_val, float* sum_val,
                                  float* out_obj, float* bleaf_val, float* out_bg, const stove_scene_seq* seq,
                                  void* stream) {
    sl::LLArgs a{};
    if (seq) {
        STOVE_CHECK_ARG(!z && seq->n > 0 && seq->T > seq->skip && seq->skip >= 1 && seq->Z >= 4 && seq->z_sup && seq->z_s &&
                            seq->patch_w, "bad sequence block");
        STOVE_CHECK_ARG(F == seq->n * (seq->T - 1) && F * O < (1ll << 31), "F must be n * (T - 1) in sequence mode");
        a.seq = 1;
        a.sq = *seq;
        z = seq->z_sup;              // (only checked for null / alignment below)
    }
    int rc = sl_fill(a, F, O, A, B, pa, pb, align_corners, img, z, obj, bg, bg_scope, bg_cnt);
    if (rc) return rc;
    STOVE_CHECK_ARG(leaf && wlin && wlog && rlin && rlog && bleaf && brlin && brlog && patches && marg_patch && marg_bg &&
                        overlap && leaf_val && sum_val && out_obj && bleaf_val && out_bg && bleaf_il, "null pointer");
    STOVE_CHECK_ARG(il_stride > 0 && (il_stride & 31) == 0 && ((uintptr_t)bleaf_il & 15) == 0, "bad interleaved table");
    if (F == 0) return STOVE_OK;
    a.bleaf_il = bleaf_il; a.il_stride = il_stride;
    a.leaf = leaf; a.wlin = wlin; a.wlog = wlog; a.rlin = rlin; a.rlog = rlog;
    a.bleaf = bleaf; a.brlin = brlin; a.brlog = brlog;
    a.patches = patches; a.marg_patch = marg_patch; a.marg_bg = marg_bg; a.overlap = overlap;
    a.leaf_val = leaf_val; a.sum_val = sum_val; a.out_obj = out_obj; a.bleaf_val = bleaf_val; a.out_bg = out_bg;
    int grid;
    size_t smem;
    if (!sl_plan(a, obj->G, obj->S, bg->R, bg->G, &grid, &smem, false)) {
        stove_set_error("stove_scene_ll_fwd: configuration not supported by the fused kernel (see stove_scene_ll_supported)");
        return STOVE_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    auto kernel = sl::scene_ll_fwd_kernel<10, 10, 3, 6>;
    STOVE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    STOVE_KERNEL(K_SCENE_LL_FWD, s, kernel<<<grid, a.nw * 32, smem, s>>>(a));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}
