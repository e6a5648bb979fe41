/*
 * stove_b200 -- C ABI of the B200-native (sm_100a) STOVE hot path.
 *
 * The reference (jlko/STOVE) has no native code and no plugin ABI: its hot path sits
 * behind Python nn.Module methods.  This header is therefore the
 * boundary a maintainer would bind from Python (ctypes); every entry
 * point names the reference function it replaces (file:line under /root/reference).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says host.
 *   - the caller owns every buffer (outputs, workspaces); nothing is allocated inside.
 *   - all arithmetic is fp32; tensors are dense, row-major, innermost dimension last.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous, no host sync.
 *   - return 0 on success, negative on error; `stove_last_error()` gives the message.
 *   - buffers documented "accumulated" must be zeroed by the caller before the first call.
 */
#ifndef STOVE_B200_H
#define STOVE_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STOVE_OK 0
#define STOVE_ERR_ARG (-1)
#define STOVE_ERR_CUDA (-2)
#define STOVE_ERR_UNSUPPORTED (-3)

const char* stove_last_error(void);
int stove_abi_version(void);

/* Launch accounting (measurement support, used by bench.py):
 *   stove_launch_count   kernels launched by this library since the last reset
 *   stove_profile_enable bracket every launch with CUDA events on its stream (not while capturing)
 *   stove_profile_read   HOST arrays ids/ms (max_n entries): waits for the events, returns how many
 *                        (kernel id, device milliseconds) records were written, clears the log
 *   stove_kernel_name    name of a kernel id in [0, stove_kernel_count()) */
int64_t stove_launch_count(int reset);
int stove_profile_enable(int on);
int stove_profile_read(int32_t* ids, float* ms, int max_n);
const char* stove_kernel_name(int id);
int stove_kernel_count(void);

/* Library options: alternative code paths kept for the parity tests and for per-kernel timing passes.
 * Nothing in the library reads the environment.  Names: "fork" (1: independent kernels of one call run on
 * library side streams), "spn2_nodes_stage", "dynloop_generic", "dynloop_nw", "dynloop_recompute",
 * "rollout_cta", "rollout_nw", "gnn_seq_fwd", "gnn_seq_bwd", "head_par_ctas", "wgrad_ctas".  set returns the previous value (or a
 * negative error code for an unknown name). */
int stove_set_option(const char* name, int value);
int stove_get_option(const char* name);

/* ------------------------------------------------------------------------------------
 * bw_transform: sum colour channels, clamp to [0,1]   (model/utils/utils.py:10-15)
 *   x [n][C][hw] -> y [n][hw]
 * ---------------------------------------------------------------------------------- */
int stove_bw_transform(const float* x, float* y, int64_t n, int channels, int64_t hw, void* stream);
/* The same with uint8 frames accepted (x_is_u8: values are scaled by 1/255 on the fly -- host frames then
 * cross PCIe at one byte per value) and, optionally, the (hi, lo) TF32 operand planes y_planes [2][n][hw] of
 * y written in the same pass (the left operand of the recognition LSTM's input GEMM, see below).
 * x_is_cell: x is not the frames but a DEVICE cell holding their (16-byte aligned) address, read when the
 * kernel runs -- a captured CUDA graph is re-pointed at the next batch by writing 8 bytes instead of copying
 * the batch into a static buffer. */
int stove_bw_transform_ex(const void* x, int x_is_u8, int x_is_cell, float* y, float* y_planes, int64_t n,
                          int channels, int64_t hw, void* stream);

/* ------------------------------------------------------------------------------------
 * RAT-SPN parameter packing (model/spn/rat_torch.py:85-99 leaf variance,
 * :209-210 log_softmax of sum weights).
 *
 * leaf rows:  means/sigma_params [rows][G]  ->  packed [prow][3][GP] = (mu, 1/(2 var),
 *             0.5 log var + 0.5 log 2pi), prow = dst_row[row], GP = G rounded up to 4.
 * sum blocks: raw [nb][K][S] -> wlog/wlin [nb][K][SP] (column-wise log_softmax over K and
 *             its exp), SP = S rounded up to 4 (or 1 when S == 1).
 * The *_bwd calls map gradients w.r.t. the packed tensors back to the raw parameters.
 * ---------------------------------------------------------------------------------- */
int stove_spn_pack_leaf_fwd(const float* means, const float* sigma_params, const int32_t* dst_row,
                            int rows, int G, int GP, float min_var, float max_var,
                            float* packed, void* stream);
int stove_spn_pack_leaf_bwd(const float* sigma_params, const int32_t* dst_row, int rows, int G, int GP,
                            float min_var, float max_var, const float* g_packed,
                            float* g_means, float* g_sigma_params, void* stream);
int stove_spn_pack_sum_fwd(const float* raw, int nb, int K, int S, int SP,
                           float* wlog, float* wlin, void* stream);
int stove_spn_pack_sum_bwd(const float* wlog, int nb, int K, int S, int SP, const float* g_wlog,
                           float* g_raw, void* stream);

/* ------------------------------------------------------------------------------------
 * Object SPN ("D2" structure: R root partitions, each = product of two mid regions,
 * each mid region = one sum vector over the product of two Gauss leaves), i.e. what
 * probabilistic_models.py:8-22 builds.  Replaces RatSpn.forward for that structure
 * (rat_torch.py:333-357; leaves :83-109, products :147-163, sums :202-222).
 *
 * Structure tables (int32, device):
 *   region_scope [2R][pmax]  input indices of region q: leaf-0 pixels first, then leaf-1
 *   region_n0    [2R]        number of leaf-0 pixels
 *   region_n     [2R]        number of pixels of the region
 *   pix_slot     [D][R]      for input p and repetition r: q*pmax + position in region_scope
 * Region 2r is the first input of root product r, region 2r+1 the second.
 * Packed parameters: leaf [2R*pmax][3][GP]; wlin/wlog [2R][G*G][SP]; rlin/rlog [R][S*S].
 * Saved activations: leaf_val [2R*2*G][npad], sum_val [2R*S][npad], npad = N rounded up to 32.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int32_t D, R, G, S, pmax;
    const int32_t* region_scope;
    const int32_t* region_n0;
    const int32_t* region_n;
    const int32_t* pix_slot;
} stove_spn2_struct;

int stove_spn2_fwd(const stove_spn2_struct* st, int64_t N, const float* x, const float* marg,
                   const float* leaf, const float* wlin, const float* wlog,
                   const float* rlin, const float* rlog,
                   float* leaf_val, float* sum_val, float* out, void* stream);
size_t stove_spn2_bwd_workspace(const stove_spn2_struct* st, int64_t N);
/* g_x / g_marg may be NULL.  g_leaf, g_wlog, g_rlog are accumulated.  The parameter-gradient kernels run on
 * library-owned side streams; they are joined into `join_stream` (the stream that consumes the parameter
 * gradients, e.g. where the packing's backward runs) or, if that is NULL, into `stream`.  With a separate
 * join_stream the caller's stream only carries the node pass and the input gradients. */
int stove_spn2_bwd(const stove_spn2_struct* st, int64_t N, const float* x, const float* marg,
                   const float* leaf, const float* wlin, const float* wlog,
                   const float* rlin, const float* rlog,
                   const float* leaf_val, const float* sum_val, const float* out, const float* g_out,
                   float* g_x, float* g_marg, float* g_leaf, float* g_wlog, float* g_rlog,
                   void* workspace, void* stream, void* join_stream);

/* ------------------------------------------------------------------------------------
 * Background SPN ("D1" structure: R root partitions, each the product of two Gauss
 * leaves that split all D inputs), probabilistic_models.py:25-39.
 *   side [D][R]   0/1: which leaf of repetition r owns input p (0 = first product input)
 * Packed: leaf [D*R][3][GP]; rlin/rlog [R][G*G].  Saved: leaf_val [R*2*G][npad].
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int32_t D, R, G;
    const int32_t* side;
} stove_spn1_struct;

size_t stove_spn1_fwd_workspace(const stove_spn1_struct* st, int64_t N);
int stove_spn1_fwd(const stove_spn1_struct* st, int64_t N, const float* x, const float* marg,
                   const float* leaf, const float* rlin, const float* rlog,
                   float* leaf_val, float* out, void* workspace, void* stream);
size_t stove_spn1_bwd_workspace(const stove_spn1_struct* st, int64_t N);
int stove_spn1_bwd(const stove_spn1_struct* st, int64_t N, const float* x, const float* marg,
                   const float* leaf, const float* rlin, const float* rlog,
                   const float* leaf_val, const float* out, const float* g_out,
                   float* g_x, float* g_marg, float* g_leaf, float* g_rlog,
                   void* workspace, void* stream, void* join_stream);

/* ------------------------------------------------------------------------------------
 * Glimpse + marginalisation masks: Supair.patches_from_z (supair.py:241-276) and
 * Supair.masks_from_z (supair.py:278-356), one launch for all objects of a frame.
 *   img [F][C][A][B], z [F][O][4] = (sx, sy, x, y)
 *   patches, marg_patch [F*O][C][pa][pb];  marg_bg [F][C][A][B];  overlap [F][O]
 * x walks the last image axis, y the second-to-last (supair.py:244-247).
 * The backward returns d/dz only (frames are data).
 * ---------------------------------------------------------------------------------- */
int stove_scene_fwd(int64_t F, int O, int C, int A, int B, int pa, int pb, int align_corners,
                    const float* img, const float* z,
                    float* patches, float* marg_patch, float* marg_bg, float* overlap, void* stream);
int stove_scene_bwd(int64_t F, int O, int C, int A, int B, int pa, int pb, int align_corners,
                    const float* img, const float* z,
                    const float* g_patches, const float* g_marg_patch, const float* g_marg_bg,
                    const float* g_overlap, float* g_z, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused scene likelihood (csrc/scene_ll.cu): stove_scene_fwd + stove_spn2_fwd + stove_spn1_fwd in ONE launch for
 * single-channel frames -- the op sequence of Supair.likelihood, model/video_prediction/supair.py:62-76 (masks ->
 * background SPN, glimpses -> object SPN).  A CTA owns a group of frames; glimpses and masks are produced into
 * the object SPN's shared-memory tiles and consumed there, the background mask stays on chip for the background
 * SPN's leaf pass.  Outputs are those of the three unfused calls, in their layouts:
 *   patches / marg_patch (F*O, pa*pb), marg_bg (F, A*B), overlap (F, O)         [stove_scene_fwd]
 *   leaf_val [2R*2*G][npad], sum_val [2R*S][npad], out_obj (F*O)  npad = roundup(F*O, 32)   [stove_spn2_fwd]
 *   bleaf_val [R*2*G][npad_f], out_bg (F)                          npad_f = roundup(F, 32)   [stove_spn1_fwd]
 * bg_scope [2R][D] / bg_cnt [2R]: pixels of background leaf l = 2 r + side, ascending (int32, device).
 * stove_scene_ll_supported returns 1 when the configuration fits the fused kernels (C = 1, the D2 / D1
 * structures with 10 Gaussians / 10 sums and 3 x 6, shared memory); otherwise use the unfused calls.
 * ------------------------------------------------------------------------------------ */
/* Sequence mode of the fused scene likelihood (optional, NULL = plain mode): the kernels read the states of the
 * scored frames straight from the sequence tensors (Stove.stove_forward, model/video_prediction/stove.py:731-736:
 * z_sup for 1 <= t < skip, the sampled z for t >= skip, both as [sx, sy/sx, x, y]; frame f = b (T-1) + (t-1)) and
 * weight the object terms by sx * sy themselves -- no stove_zall_* launches, no (F, O, 4) state tensor, and
 * stove_elbo_fwd (z_all = NULL) only sums per-frame terms.  In this mode `z`, and in the backward pass g_obj / g_bg /
 * g_overlap / g_z, are NULL: the backward kernel derives the per-frame weights from the scalar g_elbo (what
 * stove_elbo_bwd does) and writes the gradients of z_sup (all rows), z_s (all Z components), logq and trans (what
 * stove_zall_bwd and one gradient add did). */
typedef struct {
    int64_t n;                 /* sequences */
    int32_t T, skip, Z;        /* frames per sequence, first dynamics frame, width of a z_s row (>= 4) */
    float beta;                /* overlap prior: log Exponential(beta) */
    const float* z_sup;        /* (n, T, O, 4) */
    const float* z_s;          /* (n, T - skip, O, Z) */
    float* patch_w;            /* forward out (F*O): object log-likelihoods weighted by sx * sy (supair.py:79), the
                                  `patch` input of stove_elbo_fwd with z_all = NULL */
    const float* g_elbo;       /* backward: d loss / d elbo (device scalar) */
    float* g_z_sup;            /* backward out (n, T, O, 4) */
    float* g_z_s;              /* backward out (n, T - skip, O, Z) */
    float* g_logq;             /* backward out (n, T - skip) */
    float* g_trans;            /* backward out (n, T - skip) */
} stove_scene_seq;

/* bleaf_il: lane-interleaved copy of the background leaf table (stove_spn_interleave_leaf) -- blocks of 32 rows laid
 * out [6 float4 parts][32 rows] so that a warp whose lanes own consecutive rows loads it fully coalesced.  Row order
 * of the forward pass: (leaf l, position in bg_scope[l]), leaves il_stride rows apart; of the backward pass:
 * (repetition r, pixel), repetitions il_stride rows apart.  il_stride is a multiple of 32; unused rows are zero. */
int stove_spn_interleave_leaf(const float* leaf, const int32_t* row_map, int64_t rows, float* out, void* stream);
int stove_scene_ll_supported(int64_t F, int O, int C, int A, int B, int pa, int pb,
                             const stove_spn2_struct* obj, const stove_spn1_struct* bg);
int stove_scene_ll_fwd(int64_t F, int O, int A, int B, int pa, int pb, int align_corners,
                       const float* img, const float* z,
                       const stove_spn2_struct* obj, const float* leaf, const float* wlin, const float* wlog,
                       const float* rlin, const float* rlog,
                       const stove_spn1_struct* bg, const int32_t* bg_scope, const int32_t* bg_cnt,
                       const float* bleaf, const float* brlin, const float* brlog,
                       const float* bleaf_il, int il_stride,
                       float* patches, float* marg_patch, float* marg_bg, float* overlap,
                       float* leaf_val, float* sum_val, float* out_obj, float* bleaf_val, float* out_bg,
                       const stove_scene_seq* seq, void* stream);

/* Backward of stove_scene_ll_fwd (csrc/scene_ll_bwd.cu): one chain launch replaces spn2_bwd (node + input pass),
 * spn1_bwd (root + input pass) and stove_scene_bwd -- gradients of the glimpses and masks stay in shared memory --
 * followed by the four parameter-gradient kernels of stove_spn2_bwd / stove_spn1_bwd on library side streams
 * (joined into join_obj / join_bg, or into `stream` when those are NULL).  g_obj (F*O), g_bg (F), g_overlap (F, O) or
 * NULL -> g_z (F, O, 4); g_leaf / g_wlog / g_rlog / g_bleaf / g_brlog are accumulated (zero them first).
 * ws_obj / ws_bg: stove_spn2_bwd_workspace(obj, F*O) / stove_spn1_bwd_workspace(bg, F) bytes. */
int stove_scene_ll_bwd(int64_t F, int O, int A, int B, int pa, int pb, int align_corners,
                       const float* img, const float* z,
                       const stove_spn2_struct* obj, const float* leaf, const float* wlin, const float* wlog,
                       const float* rlin, const float* rlog,
                       const stove_spn1_struct* bg, const int32_t* bg_scope, const int32_t* bg_cnt,
                       const float* bleaf, const float* brlin, const float* brlog,
                       const float* bleaf_il, int il_stride,
                       const float* patches, const float* marg_patch, const float* marg_bg,
                       const float* leaf_val, const float* sum_val, const float* out_obj,
                       const float* bleaf_val, const float* out_bg,
                       const float* g_obj, const float* g_bg, const float* g_overlap,
                       float* g_z, float* g_leaf, float* g_wlog, float* g_rlog, float* g_bleaf, float* g_brlog,
                       void* ws_obj, void* ws_bg, const stove_scene_seq* seq, void* stream, void* join_obj,
                       void* join_bg);

/* ------------------------------------------------------------------------------------
 * Sequence glue before the dynamics loop, one launch: Supair.constrain_zp (supair.py:112-149),
 * Stove.match_objects (stove.py:200-329 / 331-430 / 432-514), Stove.fix_supair
 * (stove.py:516-571), Stove.v_from_state / v_std_from_pos (stove.py:54-101).
 *   zp [n][T][O][8] raw encoder output, app [n][T][O][3] or NULL
 *   -> z_sup [n][T][O][4] (matched + smoothed means), z_full / std_full [n][T][O][6]
 *      (means / stds with finite-difference velocities, zeros at t = 0), app_out [n][T][O][3]
 *      (matched appearances, may be NULL), idx / flag [n][T][O] int32 (saved for the backward)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int32_t T, num_obj;
    int32_t match_kind;          /* 0 = '3_only', 1 = 'greedy', 2 = 'volatile' */
    int32_t app_dim;             /* 0 or 3 */
    int32_t match_appearance;    /* debug_match_appearance */
    int32_t fix_supair;          /* debug_fix_supair */
    float min_obj_scale, max_obj_scale, min_y_scale, max_y_scale, obj_pos_bound, scale_var, pos_var;
    float fix_eps;               /* 0.095 in the reference */
} stove_sup_cfg;

int stove_sup_prepare_fwd(const stove_sup_cfg* cfg, int64_t n, const float* zp, const float* app,
                          float* z_sup, float* z_full, float* std_full, float* app_out,
                          int32_t* idx, int32_t* flag, void* stream);
int stove_sup_prepare_bwd(const stove_sup_cfg* cfg, int64_t n, const float* zp, const int32_t* idx,
                          const int32_t* flag, const float* std_full, const float* g_z_sup,
                          const float* g_z_full, const float* g_std_full, float* g_zp, void* stream);

/* ------------------------------------------------------------------------------------
 * GNN dynamics: Dynamics.forward + core (dynamics.py:181-265) for core_idx 0, and the
 * rollout loop Stove.rollout (stove.py:777-861).
 *
 * Weights: one flat buffer, every matrix stored [in][out]; order
 *   [act_emb W (A x O*4), b]            if action_dim > 0
 *   enc W (in_dim x cl), b
 *   self0, self1 (cl x cl)
 *   rel0|att0 (2cl+1 x 4cl: rel columns first), b (4cl)
 *   rel1 (2cl x cl), att1 (2cl x cl), rel2 (cl x cl), att2 (cl x 1)
 *   aff0, aff1, aff2 (cl x cl), out0 (2cl x cl), out1 (cl x cl)
 *   [rew00, rew02 (cl x cl), rew10 (cl x cl/2), rew12 (cl/2 x cl/4), rew14 (cl/4 x 1)] if reward
 * each followed by its bias.  `stove_gnn_weight_count` returns the float count.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int32_t num_obj;      /* O */
    int32_t cl;           /* 32 */
    int32_t action_dim;   /* A, 0 = not action conditioned */
    int32_t app_dim;      /* 3 if appearances are concatenated, else 0 */
    int32_t reward;       /* 1 = reward head present */
    int32_t lim_enc;      /* raw pass-through dims (2) */
    int32_t nonlin;       /* 0 = leaky_relu(0.01) (reference default), 1 = elu */
    int32_t state_dim;    /* width of the state rows fed to the encoder; 0 = cl/2 (the STOVE model).  Other
                           * values (supairvised/dynamics.py:24-25: enc_input_size 16 with lim_enc 4) are
                           * served by the single-step kernels stove_gnn_fwd / stove_gnn_bwd only */
} stove_gnn_cfg;

int64_t stove_gnn_weight_count(const stove_gnn_cfg* cfg);
/* float offsets of every weight/bias segment (each padded to a multiple of 4 floats), in the
 * order act, enc, self0, self1, rel0|att0, rel1, att1, rel2, att2, aff0, aff1, aff2, out0,
 * out1, rew00, rew02, rew10, rew12, rew14 (w then b each; -1 if absent), then the total.
 * Returns the number of entries written (39). */
int stove_gnn_weight_offsets(const stove_gnn_cfg* cfg, int32_t* out, int max_out);
size_t stove_gnn_bwd_workspace(const stove_gnn_cfg* cfg, int64_t n);
/* s [n][O][state_dim or cl/2], actions [n][A] or NULL, app [n][O][app_dim] or NULL
 * -> out [n][O][cl], reward [n] (sigmoid applied; NULL if no reward head) */
int stove_gnn_fwd(const stove_gnn_cfg* cfg, int64_t n, const float* s, const float* actions,
                  const float* app, const float* weights, float* out, float* reward, void* stream);
/* g_weights is overwritten (not accumulated). g_reward may be NULL. */
int stove_gnn_bwd(const stove_gnn_cfg* cfg, int64_t n, const float* s, const float* actions,
                  const float* app, const float* weights, const float* g_out, const float* g_reward,
                  float* g_s, float* g_weights, void* workspace, void* stream);
/* z_last [n][O][cl/2+2]; actions [n][L][A] (wrapped modulo L) or NULL; app [n][O][app_dim]
 * or NULL; noise [n][num][O][cl/2] or NULL (=> mean prediction).
 * z_out [n][num][O][cl/2+2]; std_out / logq_out [n][num][O][cl/2] or NULL;
 * rewards [n][num] or NULL. */
int stove_gnn_rollout(const stove_gnn_cfg* cfg, int64_t n, int num, const float* z_last,
                      const float* actions, int action_len, const float* app, const float* weights,
                      const float* noise, float pos_var, float vel_std, float latent_std,
                      float* z_out, float* std_out, float* logq_out, float* rewards, void* stream);

/* ------------------------------------------------------------------------------------
 * One step of the dynamics loop of Stove.stove_forward (stove.py:696-713), fused: GNN core
 * (dynamics.py:181-265) + constrain_z_dyn (dynamics.py:147-179) + position integration +
 * full_state (Gaussian fusion with the SuPAIR state, reparametrised sample, log q;
 * stove.py:103-170) + transition_lik of the sample (stove.py:172-198).
 * Every tensor is addressed as base + sequence * stride (strides in floats) so that time
 * slices of (n, T, ...) tensors can be passed without copies.  Per sequence:
 *   z_prev [O][cl/2+2], sup / sup_std [O][6] (sx, sy/sx, x, y, vx, vy), eps [O][cl/2+2],
 *   actions [A], app [O][app_dim]  ->  z_out [O][cl/2+2], z_dyn / z_dyn_std [O][cl/2],
 *   z_std [O][cl/2+2] (optional), logq / trans / reward: one float each.
 * Backward: g_z = g_z_a + g_z_b (either may be NULL), g_logq / g_trans / g_reward one float per
 * sequence -> g_z_prev [O][cl/2+2], g_sup / g_sup_std [O][6] (overwritten), g_weights.
 * Workspace: stove_gnn_bwd_workspace bytes.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    float pos_var, vel_std, latent_std;   /* constrain_z_dyn scales */
    float trans_std[32];                  /* transition_lik_std, cl/2 entries used */
} stove_fuse_cfg;

typedef struct {
    const float* z_prev; int64_t z_prev_ss;
    const float* sup; const float* sup_std; int64_t sup_ss;
    const float* eps; int64_t eps_ss;
    const float* actions; int64_t act_ss;
    const float* app; int64_t app_ss;
    float* z_out; int64_t z_out_ss;
    float* z_dyn; float* z_dyn_std; int64_t zdyn_ss;
    float* z_std; int64_t z_std_ss;
    float* logq; float* trans; float* reward; int64_t sc_ss;
    const float* g_z_a; int64_t g_z_a_ss;
    const float* g_z_b; int64_t g_z_b_ss;
    const float* g_logq; const float* g_trans; const float* g_reward; int64_t g_sc_ss;
    float* g_z_prev; int64_t g_z_prev_ss;
    float* g_sup; float* g_sup_std; int64_t g_sup_ss;
} stove_dynstep_io;

int stove_dynstep_fwd(const stove_gnn_cfg* cfg, const stove_fuse_cfg* fuse, int64_t n,
                      const stove_dynstep_io* io, const float* weights, void* stream);
/* `first` / `last` mark the first and last call of one backward pass over the time steps: the
 * per-CTA weight-gradient slabs in `workspace` are cleared on the first call, accumulate over
 * the calls, and are reduced into g_weights (overwritten) on the last. */
int stove_dynstep_bwd(const stove_gnn_cfg* cfg, const stove_fuse_cfg* fuse, int64_t n,
                      const stove_dynstep_io* io, const float* weights, float* g_weights,
                      int first, int last, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------
 * The WHOLE dynamics loop of Stove.stove_forward (stove.py:696-713) in one call: for every
 * t in [skip, T) the fused step above, chained through z_t on the device.  For O = 3, cl = 32
 * (every BASELINE config but multiball) this is one persistent warp-team kernel forward and
 * three kernels backward (chain, weight gradients, slab reduction; csrc/dynloop.cu); any other
 * shape runs the per-step kernels above, chained inside the library.
 *   z_init [n][O][Z] (Z = cl/2 + 2); sup / sup_std [n][T][O][6]; eps [T-skip][n][O][Z];
 *   actions [n][T][A] or NULL and app [n][T][O][app_dim] or NULL (step t reads index t-1)
 *   -> z [n][S][O][Z], z_dyn / z_dyn_std [n][S][O][Z-2], z_std [n][S][O][Z] (optional),
 *      logq / trans / reward [n][S]  (S = T - skip; reward may be NULL)
 * Backward: g_z [n][S][O][Z], g_logq / g_trans / g_reward [n][S] (each may be NULL)
 *   -> g_z_init [n][O][Z], g_sup / g_sup_std [n][T][O][6] (fully overwritten), g_weights
 *      (overwritten).  `z` must hold the forward result.  Workspace: stove_dynloop_bwd_workspace.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int32_t T, skip;
    const float* z_init; const float* sup; const float* sup_std; const float* eps;
    const float* actions; const float* app;
    float* z; float* z_dyn; float* z_dyn_std; float* z_std;
    float* logq; float* trans; float* reward;
    const float* g_z; const float* g_logq; const float* g_trans; const float* g_reward;
    float* g_z_init; float* g_sup; float* g_sup_std;
    float* xrec;   /* optional [stove_dynloop_xrec_floats]: forward keeps its activations here and the
                      backward reloads them instead of recomputing each step (NULL: recompute) */
} stove_dynloop_io;

int stove_dynloop_fwd(const stove_gnn_cfg* cfg, const stove_fuse_cfg* fuse, int64_t n,
                      const stove_dynloop_io* io, const float* weights, void* stream);
size_t stove_dynloop_bwd_workspace(const stove_gnn_cfg* cfg, int64_t n, int T, int skip);
/* Same, with the weight-gradient kernels (which nothing on the sequential chain waits for) launched on
 * `wgrad_stream` after the chain kernel on `stream`; there is NO join: g_weights and the workspace are
 * valid on `wgrad_stream`.  Shapes without the fast path run everything on `stream`. */
int stove_dynloop_bwd2(const stove_gnn_cfg* cfg, const stove_fuse_cfg* fuse, int64_t n,
                       const stove_dynloop_io* io, const float* weights, float* g_weights,
                       void* workspace, void* stream, void* wgrad_stream);
/* floats of the optional activation buffer `xrec` (0: this shape has no such path) */
int64_t stove_dynloop_xrec_floats(const stove_gnn_cfg* cfg, int64_t n, int T, int skip);
/* ------------------------------------------------------------------------------------
 * z of every scored frame and the ELBO assembly of Stove.stove_forward (stove.py:731-748,
 * supair.py:84-110, Supair.sy_from_quotient supair.py:151-158).
 *   zall: z_sup [n][T][O][4], z_s [n][T-skip][O][Z] ([sx, sy/sx, x, y | ...]) -> z_all [n][T-1][O][4]
 *         ([sx, sy, x, y] of frames 1 .. T-1: SuPAIR state for t < skip, sampled state after);
 *         backward overwrites g_z_sup [n][T][O][4] and g_z_s [n][T-skip][O][Z] completely.
 *   elbo: bg [F], patch [F][O] (raw object-SPN log-likelihoods), z_all [F][O][4], overlap [F][O]
 *         (F = n (T-1)), logq / trans [n][T-skip] -> stats[8] = {average ELBO, mean bg, mean patch,
 *         mean overlap prior (frames t >= skip), mean log q, mean transition lik, mean SuPAIR-frame
 *         likelihood, 0}, elbo_out[1] = the average ELBO again (its own buffer);  backward: g = d loss / d ELBO (one float on the device).
 * ---------------------------------------------------------------------------------- */
/* stove_elbo_fwd with z_all = NULL: `patch` already carries the sx * sy weight (sequence mode of stove_scene_ll_fwd). */
int stove_zall_fwd(int64_t n, int T, int skip, int O, int Z, const float* z_sup, const float* z_s,
                   float* z_all, void* stream);
int stove_zall_bwd(int64_t n, int T, int skip, int O, int Z, const float* z_sup, const float* z_s,
                   const float* g_z_all, float* g_z_sup, float* g_z_s, void* stream);
int stove_elbo_fwd(int64_t n, int T, int skip, int O, float beta, const float* bg, const float* patch,
                   const float* z_all, const float* overlap, const float* logq, const float* trans,
                   float* stats, float* elbo_out, void* stream);
int stove_elbo_bwd(int64_t n, int T, int skip, int O, float beta, const float* g, const float* patch,
                   const float* z_all, float* g_bg, float* g_patch, float* g_z_all, float* g_overlap,
                   float* g_logq, float* g_trans, void* stream);

/* ------------------------------------------------------------------------------------
 * Recognition LSTM (encoder.py:28-51: nn.LSTM fed the same frame num_obj times) on the tensor cores,
 * forward and backward (csrc/lstm_tc.cu).  Every contraction is a 3xTF32 product on tcgen05 (kind::tf32,
 * accumulator in TMEM, operands by TMA): an fp32 operand X [rows][K] is stored once as two PLANES
 * [2][rows][K] -- hi = X with the low 13 mantissa bits cleared (TF32-exact), lo = X - hi -- and
 * X Y^T ~= hi hi^T + hi lo^T + lo hi^T accumulates in one fp32 accumulator (fp32-level accuracy).
 * All pointers 16-byte aligned; leading dimensions / plane strides multiples of 4 floats.
 * ---------------------------------------------------------------------------------- */

/* x [rows][cols] (row stride ldx) -> pl [2][rows][cols] and / or the transposed planes plT [2][cols][ldT]
 * (ldT >= rows; columns rows .. ldT-1 are written as zero).  Either output may be NULL. */
int stove_split_planes(int64_t rows, int cols, const float* x, int64_t ldx, float* pl, float* plT,
                       int64_t ldT, void* stream);

/* One LSTM step: gates = A B^T + addend (A_pl [2][n][K] = planes of the frame or of h_{t-1}; B_pl [2][4H][K] =
 * planes of W_ih or W_hh, nn.LSTM gate order i, f, g, o along the rows), then the LSTM cell in the GEMM's
 * epilogue.  addend = bias [4H] (addend_is_bias) or the input-GEMM gates [n][4H]; c_prev [n][H] or NULL
 * (zero state); gx_out [n][4H] (optional) receives GEMM + addend; h_out has row stride h_ld (a slice of the
 * stacked (n, steps, H) output); act [n][4H] keeps the activated gates for the backward pass; h_pl [2][n][H]
 * (optional) = planes of h for the next step; hT_pl (optional, needs h_pl) = TRANSPOSED planes of h with
 * row stride ldT and plane stride hT_plane: rows = hidden units, columns 0 .. n-1 = frames, columns
 * n .. spanT-1 zero (a step's block of a buffer stacked along the columns for the W_hh gradient).
 * H % 32 == 0, K % 4 == 0. */
int stove_lstm_gemm_cell_fwd(int64_t n, int H, int64_t K, const float* A_pl, const float* B_pl, const float* addend,
                             int addend_is_bias, const float* c_prev, float* gx_out, float* h_out, int64_t h_ld,
                             float* c_out, float* act, float* h_pl, float* hT_pl, int64_t ldT, int64_t spanT,
                             int64_t hT_plane, void* stream);

/* D = A B^T over K columns: A_pl [2][M][lda], B_pl [2][N][ldb] (plane strides a_plane / b_plane floats),
 * written as `parts` split-K partial products D[p][M][ldd] (part stride part_stride floats; the caller sums
 * them, stove_sum_parts).  The library may lower `parts` (never below 1): stove_tc3_gemm_parts returns the
 * count it will use for (M, N, K, want).  Used for the hidden-state gradient (gate gradient x W_hh) and for
 * both weight gradients of the LSTM (contractions over the frames: transposed planes as operands).
 * bn: tile width (0 = automatic: 128, or 64 for N <= 64; 256 = fewest operand bytes, half the CTAs). */
int stove_tc3_gemm(int64_t M, int64_t N, int64_t K, const float* A_pl, int64_t lda, int64_t a_plane,
                   const float* B_pl, int64_t ldb, int64_t b_plane, float* D, int64_t ldd, int parts,
                   int64_t part_stride, int bn, void* stream);
int stove_tc3_gemm_parts(int64_t M, int64_t N, int64_t K, int want);

/* Gate/state backward of one LSTM step.  g_h = g_h_a (row stride g_h_a_ld) + the g_h_b_parts split-K parts
 * g_h_b [parts][n][H] (NULL: none); g_c, c_prev, g_c_prev may be NULL.  Outputs: g_pl [2][n][4H] = planes of
 * this step's gate gradient (optional); gT_pl = transposed planes (rows = 4H gate units, row stride ldT,
 * plane stride gT_plane, columns n .. spanT-1 zero) of this step's gate gradient or, with emit_acc, of the
 * running sum over the steps (optional); g_acc [n][4H] = running sum (acc_mode 0: start, 1: add; not written
 * when emit_acc; may be NULL when acc_mode == 0 and emit_acc); bias_part [ceil(max(n, spanT) / 32)][4H] =
 * column sums of the emitted gradient over blocks of 32 frames (optional; sum them with stove_sum_parts). */
int stove_lstm_cell_bwd_t(int64_t n, int H, const float* act, const float* c_prev, const float* c_out,
                          const float* g_h_a, int64_t g_h_a_ld, const float* g_h_b, int g_h_b_parts,
                          const float* g_c, float* g_pl, float* gT_pl, int64_t ldT, int64_t spanT,
                          int64_t gT_plane, float* g_acc, int acc_mode, int emit_acc, float* bias_part,
                          float* g_c_prev, void* stream);

/* out[i] = scale * sum over p < parts of in[p * stride + i], fixed order; numel % 4 == 0, stride % 4 == 0
 * (scale = 1 / world size lets the last reduction of a weight gradient write straight into the data-parallel
 * gradient bucket). */
int stove_sum_parts(int64_t numel, int parts, int64_t stride, const float* in, float* out, float scale, void* stream);

/* Output head of the recognition network (encoder.py:53-56): out = fc2(sigmoid(fc1(x))), x [R][K],
 * w1 [J][K], b1 [J], w2 [P][J], b2 [P] (nn.Linear layouts); hidden [R][J] = sigmoid(fc1(x)) is kept for the
 * backward pass.  K <= 256 and K % 32 == 0, J <= 64, P <= 16, else STOVE_ERR_UNSUPPORTED.  The backward
 * pass is two calls sharing ws (stove_enc_head_bwd_workspace bytes): _data writes g_x [R][K] -- what the rest
 * of the backward chain waits for -- and leaves the pre-activation gradients in ws; _params (after _data, on
 * any stream ordered behind it) turns them into the parameter gradients (overwritten, fixed summation order). */
int stove_enc_head_fwd(int64_t R, int K, int J, int P, const float* x, const float* w1, const float* b1,
                       const float* w2, const float* b2, float* hidden, float* out, void* stream);
size_t stove_enc_head_bwd_workspace(int64_t R, int K, int J, int P);
int stove_enc_head_bwd_data(int64_t R, int K, int J, int P, const float* w1, const float* w2,
                            const float* hidden, const float* g_out, float* g_x, float* ws, void* stream);
int stove_enc_head_bwd_params(int64_t R, int K, int J, int P, const float* x, const float* hidden,
                              const float* g_out, float* g_w1, float* g_b1, float* g_w2, float* g_b2,
                              float* ws, void* stream);

/* Gather `count` device tensors into one flat fp32 buffer (the data-parallel gradient bucket):
 * dst[offsets[i] .. offsets[i] + numels[i]) = srcs[i][0 .. numels[i]).  srcs / offsets / numels are HOST
 * arrays (device addresses travel as kernel parameters: capturable, no table upload).  Every value is
 * multiplied by `scale` on the way (1 / world size folds the averaging of the all-reduce into the gather). */
int stove_gather_flat(const void* const* srcs, const int64_t* offsets, const int64_t* numels, int count,
                      float* dst, float scale, void* stream);

/* Optimizer step of the reference trainer on the flat gradient bucket (train.py:46-49 Adam with amsgrad,
 * :471-473 clip_grad_norm_(parameters, max_norm) then step).  params / offsets / numels are HOST arrays
 * (count entries): parameter i occupies flat_grad[offsets[i] .. + numels[i]) and is updated in place at
 * params[i]; exp_avg / exp_avg_sq / max_exp_avg_sq (NULL: no amsgrad) are flat like the bucket; partial
 * holds stove_adam_workspace_floats() floats; lr and step are DEVICE scalars (step is incremented by the
 * call, so a captured graph keeps counting); max_norm <= 0 disables clipping.  The clipped gradient is
 * written back to flat_grad, as clip_grad_norm_ does. */
int stove_adam_workspace_floats(void);
int stove_adam_step(const void* const* params, const int64_t* offsets, const int64_t* numels, int count,
                    int64_t total, float* flat_grad, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq,
                    float* partial, const float* lr, float* step, float beta1, float beta2, float eps,
                    float max_norm, void* stream);

/* Renderer: frames from states (Supair.reconstruct_from_z, supair.py:425-501):
 *   out [F][C][A][B] = clamp(bg + sum_o paste(patches[f, o], z[f, o]), 0, 1),
 * paste = F.grid_sample(patch, F.affine_grid(inverse of [[sx,0,x],[0,sy,y]], (A, B))), bilinear, zero padding.
 * bg [C][A][B] (or [F][C][A][B] if bg_per_frame); patches [O][C][pa][pb] (or [F][O][C][pa][pb] if
 * patches_per_frame); z [F][O][4] = (sx, sy, x, y).  No gradient. */
int stove_render(int64_t F, int O, int C, int A, int B, int pa, int pb, int align_corners, const float* bg,
                 int bg_per_frame, const float* patches, int patches_per_frame, const float* z, float* out,
                 void* stream);

/* Measurement infrastructure (csrc/microbench.cu; not on the product path): the roofline denominators that
 * MEASURED_PEAKS.json does not carry.  FP32 FMA throughput of ctas x 1024 threads x iters x 16 independent FMAs,
 * and the issue peak of tcgen05.mma kind::tf32 128 x n_cols x 8 (n_cols 128 or 256) with operands resident in
 * shared memory, both in TFLOP/s; scripts/microbench.py writes them to profiles/r02_microbench.json. */
int stove_microbench_ffma(int ctas, int iters, float* scratch, double* tflops, void* stream);
int stove_microbench_tf32(int ctas, int iters, int n_cols, double* tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif
