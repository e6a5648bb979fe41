// Shared by gnn.cu (CTA-wide generic kernels) and dynloop.cu (warp-team kernels for O = 3, cl = 32):
// flat weight layout of the C ABI, activations, host-side argument checks.
#pragma once
#include <stdlib.h>
#include "common.cuh"

enum { ACT_NONE = 0, ACT_NL = 1, ACT_TANH = 2, ACT_RELU = 3, ACT_SIGMOID = 4, ACT_EXP = 5 };

struct GnnLayout {
    int in_dim;
    int act_w, act_b, enc_w, enc_b, self0_w, self0_b, self1_w, self1_b, ra0_w, ra0_b, rel1_w, rel1_b,
        att1_w, att1_b, rel2_w, rel2_b, att2_w, att2_b, aff0_w, aff0_b, aff1_w, aff1_b, aff2_w, aff2_b,
        out0_w, out0_b, out1_w, out1_b, rew00_w, rew00_b, rew02_w, rew02_b, rew10_w, rew10_b, rew12_w,
        rew12_b, rew14_w, rew14_b;
    int total;
};

static inline int pad4(int v) { return (v + 3) / 4 * 4; }
// width of the state rows: cl/2 for the STOVE model, enc_input_size for the supervised ablation
__host__ __device__ static inline int gnn_sdim(const stove_gnn_cfg& c) { return c.state_dim > 0 ? c.state_dim : c.cl / 2; }

static inline GnnLayout gnn_layout(const stove_gnn_cfg* c) {
    GnnLayout L;
    const int cl = c->cl, O = c->num_obj;
    L.in_dim = gnn_sdim(*c) + (c->action_dim > 0 ? 4 : 0) + c->app_dim;
    int at = 0;
    auto seg = [&](int& w, int& b, int K, int N) {
        w = at; at += pad4(K * N);
        b = at; at += pad4(N);
    };
    L.act_w = L.act_b = -1;
    if (c->action_dim > 0) seg(L.act_w, L.act_b, c->action_dim, O * 4);
    seg(L.enc_w, L.enc_b, L.in_dim, cl);
    seg(L.self0_w, L.self0_b, cl, cl);
    seg(L.self1_w, L.self1_b, cl, cl);
    seg(L.ra0_w, L.ra0_b, 2 * cl + 1, 4 * cl);
    seg(L.rel1_w, L.rel1_b, 2 * cl, cl);
    seg(L.att1_w, L.att1_b, 2 * cl, cl);
    seg(L.rel2_w, L.rel2_b, cl, cl);
    seg(L.att2_w, L.att2_b, cl, 1);
    seg(L.aff0_w, L.aff0_b, cl, cl);
    seg(L.aff1_w, L.aff1_b, cl, cl);
    seg(L.aff2_w, L.aff2_b, cl, cl);
    seg(L.out0_w, L.out0_b, 2 * cl, cl);
    seg(L.out1_w, L.out1_b, cl, cl);
    L.rew00_w = L.rew00_b = L.rew02_w = L.rew02_b = L.rew10_w = L.rew10_b = L.rew12_w = L.rew12_b =
        L.rew14_w = L.rew14_b = -1;
    if (c->reward) {
        seg(L.rew00_w, L.rew00_b, cl, cl);
        seg(L.rew02_w, L.rew02_b, cl, cl);
        seg(L.rew10_w, L.rew10_b, cl, cl / 2);
        seg(L.rew12_w, L.rew12_b, cl / 2, cl / 4);
        seg(L.rew14_w, L.rew14_b, cl / 4, 1);
    }
    L.total = at;
    return L;
}

__device__ __forceinline__ float apply_act(float v, int act, int nonlin) {
    switch (act) {
        case ACT_NL: return nonlin ? (v > 0.f ? v : expm1f(v)) : (v >= 0.f ? v : 0.01f * v);
        case ACT_TANH: return tanhf(v);
        case ACT_RELU: return fmaxf(v, 0.f);
        case ACT_SIGMOID: return sigmoidf_(v);
        case ACT_EXP: return expf(v);
        default: return v;
    }
}
// derivative of the activation expressed through its OUTPUT y
__device__ __forceinline__ float act_grad(float y, int act, int nonlin) {
    switch (act) {
        case ACT_NL: return nonlin ? (y > 0.f ? 1.f : y + 1.f) : (y > 0.f ? 1.f : 0.01f);
        case ACT_TANH: return 1.f - y * y;
        case ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case ACT_SIGMOID: return y * (1.f - y);
        case ACT_EXP: return y;
        default: return 1.f;
    }
}

struct FuseCfg {
    float scale[3];          // pos_var, 0.04, debug_latent_q_std  (constrain_z_dyn)
    float trans_std[32];     // transition_lik_std per state feature (cl/2 used)
};

static inline FuseCfg make_fuse(const stove_gnn_cfg* cfg, const stove_fuse_cfg* fc) {
    FuseCfg f;
    f.scale[0] = fc->pos_var; f.scale[1] = fc->vel_std; f.scale[2] = fc->latent_std;
    for (int i = 0; i < 32; ++i) f.trans_std[i] = (i < cfg->cl / 2) ? fc->trans_std[i] : 1.f;
    return f;
}

// the loop / rollout kernels integrate the state themselves: they need the STOVE state layout
static inline bool gnn_default_state(const stove_gnn_cfg* c) { return c->state_dim == 0 || c->state_dim == c->cl / 2; }

static inline int gnn_check(const stove_gnn_cfg* c) {
    STOVE_CHECK_ARG(c, "null cfg");
    STOVE_CHECK_ARG(c->num_obj > 0 && c->num_obj <= 16, "num_obj out of range");
    STOVE_CHECK_ARG(c->cl >= 8 && c->cl % 8 == 0 && c->cl <= 64, "cl must be a multiple of 8 in [8, 64]");
    STOVE_CHECK_ARG(c->state_dim >= 0 && c->state_dim <= c->cl, "state_dim must be in [0, cl]");
    STOVE_CHECK_ARG(c->action_dim >= 0 && c->app_dim >= 0 && c->lim_enc >= 0 && c->lim_enc <= gnn_sdim(*c), "bad cfg");
    return STOVE_OK;
}

static const size_t kMaxSmem = 227 * 1024;

