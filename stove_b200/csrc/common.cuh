// Shared helpers for the stove_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/stove_b200.h"

#define HALF_LOG_2PI 0.91893853320467274178f
// below this a linear-domain mixture sum is recomputed exactly in the log domain
#define LIN_SUM_FLOOR 1e-30f

void stove_set_error(const char* fmt, ...);

// every kernel launch goes through STOVE_KERNEL: it counts launches (stove_launch_count) and,
// when profiling is enabled (stove_profile_enable), brackets the launch with CUDA events on the
// launching stream so bench.py can report per-kernel device time from the timed region.
#include "kernel_ids.inc"
int stove_prof_begin(int id, cudaStream_t s);
void stove_prof_end(int slot, cudaStream_t s);
#define STOVE_KERNEL(id, stream, ...)                 \
    do {                                              \
        const int ps_ = stove_prof_begin(id, stream); \
        __VA_ARGS__;                                  \
        stove_prof_end(ps_, stream);                  \
    } while (0)

#define STOVE_CHECK_ARG(cond, msg)                                  \
    do {                                                            \
        if (!(cond)) {                                              \
            stove_set_error("%s: %s", __func__, msg);               \
            return STOVE_ERR_ARG;                                   \
        }                                                           \
    } while (0)

#define STOVE_CUDA(call)                                                              \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            stove_set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e_)); \
            return STOVE_ERR_CUDA;                                                    \
        }                                                                             \
    } while (0)

#define STOVE_LAUNCH_CHECK() STOVE_CUDA(cudaGetLastError())

// Fork/join of independent kernels inside one library call: `fork` makes up to two library-owned side
// streams wait for everything issued so far on the caller's stream, `join` makes the caller's stream
// wait for them.  Works eagerly and under stream capture (the branches become parallel graph nodes).
// `family` selects a private set of streams/events (0 = object SPN, 1 = background SPN, ...), so two
// callers on different streams do not serialise each other.
#define STOVE_FORK_FAMILIES 4
struct StoveFork {
    cudaStream_t side[2];
    cudaEvent_t fork_ev, join_ev[2];
};
StoveFork* stove_fork_get(int family);          // nullptr on failure (error string set)
int stove_fork(StoveFork* f, cudaStream_t s, int nside);
int stove_join(StoveFork* f, cudaStream_t s, int nside);

// Library options (stove_set_option): alternative code paths kept for the parity tests and for per-kernel
// timing passes.  They are set through the API only -- nothing in the library reads the environment.
enum StoveOption {
    OPT_FORK,                 // 1: independent kernels of one call run on library side streams (default); 0: serial
    OPT_SPN2_NODES_STAGE,     // 1: spn2_bwd_nodes stages the sum weights in shared memory (default)
    OPT_DYNLOOP_GENERIC,      // 1: dynamics loop through the generic CTA-wide kernels
    OPT_DYNLOOP_NW,           // warps per sequence of the warp-team dynamics loop (1 or 2; default 2)
    OPT_DYNLOOP_RECOMPUTE,    // 1: the loop backward recomputes each step instead of reloading activations
    OPT_ROLLOUT_CTA,          // 1: rollouts through the generic CTA-wide kernel
    OPT_ROLLOUT_NW,           // warps per sequence of the rollout kernel (1 or 2; default 2)
    OPT_GNN_SEQ_FWD,          // sequences per CTA of the generic kernels (0 = automatic)
    OPT_GNN_SEQ_BWD,
    OPT_HEAD_PAR_CTAS,
    OPT_WGRAD_CTAS,           // CTAs of dynloop_wgrad (0 = one per SM; default 74): each takes a whole SM's registers for
                              // its lifetime, so half the SMs stay free for the chain kernels that become ready while it
                              // runs (sweep in profiles/r02_tune_wgrad_ctas.txt: 0.697 ms/step at 148, 0.680 at 74)
        // CTAs of the head's parameter-gradient kernel (0 = one per SM); it runs beside the LSTM backward
    OPT_GNN_THREADS,          // threads per CTA of the generic dynamics kernels (gnn_fwd/bwd, dynstep_fwd/bwd): 256, 384, 512; 0 = by object count
    OPT_COUNT
};
int stove_opt(int id);

static inline int64_t round_up64(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }
