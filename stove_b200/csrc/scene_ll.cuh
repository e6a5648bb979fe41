// Argument block and shared-memory plan of the fused scene-likelihood kernels (scene_ll.cu, scene_ll_bwd.cu).
#pragma once
#include "common.cuh"
#include "spn_math.cuh"

namespace sl {

constexpr int HT = 16;       // patches per half-warp tile of the object SPN
constexpr int MAXF = 13;     // frames per round: the background leaf pass keeps MAXF x 6 accumulators in registers
constexpr int MAXIT = 4;     // glimpse pixels per lane: pa * pb <= 128
constexpr int MAXNW = 18;    // warps per CTA (576 threads -> 112 registers per thread)

struct LLArgs {
    // scene
    int O, A, B, pa, pb, align;
    int64_t F;
    const float* img;            // (F, 1, A, B)
    const float* z;              // (F, O, 4) = (sx, sy, x, y)
    float* patches;              // (F*O, pa*pb)   written for the parameter-gradient kernels / prop_dict
    float* marg_patch;           // (F*O, pa*pb)
    float* marg_bg;              // (F, A*B)
    float* overlap;              // (F, O)
    // plan (sl_plan)
    int rf, ntile, fs, nw, ns;
    // object SPN (D2), packed parameters of spn_obj.cu
    Spn2Dev st;
    const float *leaf, *wlin, *wlog, *rlin, *rlog;
    float *leaf_val, *sum_val, *out_obj;
    int64_t Np, npad_p;
    // background SPN (D1), packed parameters of spn_bg.cu
    int Dbg;
    const int32_t* bg_side;      // [D][R]
    const int32_t* bg_scope;     // [2R][D]: pixels of leaf l = 2 r + side, ascending
    const int32_t* bg_cnt;       // [2R]
    const float *bleaf, *brlin, *brlog;
    float *bleaf_val, *out_bg;
    int64_t npad_f;
    // backward only
    const float *g_obj, *g_bg, *g_overlap;
    float* g_z;
    float *gleaf, *aux_reg, *aux_root;       // object SPN workspace (layout of spn_obj.cu)
    float *bgleaf, *baux_root;               // background SPN workspace (layout of spn_bg.cu)
    float *g_wlog, *g_rlog, *g_brlog;        // slow-path contributions (atomics)
};

struct Smem {                  // offsets in floats
    int frames_img, frames_bg, frames_tx, frames_ty;     // union U, phases S / BG
    int lf, wl, rws, scs;                                 // union U, phase OBJ
    int xw, ss, vr, total;
};

__host__ __device__ inline int imax_(int a, int b) { return a > b ? a : b; }

__host__ __device__ inline Smem smem_layout(const LLArgs& a, int G, int S, int GB) {
    Smem m;
    const int GP = up4(G), SP = up4(S), Q = 2 * a.st.R;
    const int tXs = up4(a.B), tYs = up4(a.A);
    m.frames_img = 0;
    m.frames_bg = a.rf * a.fs;
    m.frames_tx = 2 * a.rf * a.fs;
    m.frames_ty = m.frames_tx + a.rf * tXs;
    const int frames_sz = m.frames_ty + a.rf * tYs;
    m.lf = 0;
    m.wl = m.lf + Q * a.st.pmax * 3 * GP;
    m.rws = m.wl + Q * G * G * SP;
    m.scs = m.rws + up4(a.st.R * S * S);
    const int tables_sz = m.scs + up4(Q * a.st.pmax);
    m.xw = up4(imax_(frames_sz, tables_sz));
    m.ss = m.xw + a.ntile * a.st.D * HT * 2;
    m.vr = m.ss + up4(imax_(a.ntile * Q * S * HT, 6 * a.ns * MAXF * GB));
    m.total = m.vr + a.ntile * a.st.R * HT;
    return m;
}

__host__ __device__ inline Smem smem_layout_bwd(const LLArgs& a, int G, int S, int GB) {
    return smem_layout(a, G, S, GB);       // placeholder until scene_ll_bwd.cu lands
}

}  // namespace sl

int sl_plan(sl::LLArgs& a, int G, int S, int RB, int GB, int* grid, size_t* smem_bytes, bool backward);
