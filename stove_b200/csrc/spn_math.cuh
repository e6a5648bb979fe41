// Helpers shared by the object-SPN kernels (spn_obj.cu) and the fused scene-likelihood kernels (scene_ll.cu):
// asynchronous staging copies, ordered shared-memory loads, max-shifted exponentials and the exact log-domain
// slow path of a sum node (model/spn/rat_torch.py:202-222).
#pragma once
#include "common.cuh"

struct Spn2Dev {
    int D, R, pmax;
    const int32_t* scope;
    const int32_t* n0;
    const int32_t* nt;
    const int32_t* slot;
};

template <int G>
struct GP_ {
    static constexpr int v = (G + 3) / 4 * 4;
};

// ------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------
// 16-byte asynchronous global -> shared copies.  ncu on the first version of these kernels: 7-13
// stall cycles per issued instruction on the load scoreboard -- every (pixel, region) step waited
// for its own leaf-parameter round trip to L2.  The kernels now stage the parameter blocks they
// will walk through (leaf table, sum weights, scope indices) with cp.async while the patch tile
// is loaded, and the inner loops read shared memory.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }
__device__ __forceinline__ float lds_f1(const float* p) {
    float v;
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
// volatile on purpose: with plain shared-memory loads ptxas hoists all 600 LDS.128 of the two unrolled product
// loops to the top and spills 7.6 KB per thread; pinned in program order the staged kernel keeps 100 registers
__device__ __forceinline__ float4 lds_f4(const float* p) {
    float4 v;
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

__host__ __device__ static inline int up4(int v) { return (v + 3) & ~3; }

template <int G>
__device__ __forceinline__ float shift_exp(const float (&L)[G], float (&e)[G]) {
    float m = L[0];
#pragma unroll
    for (int g = 1; g < G; ++g) m = fmaxf(m, L[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) e[g] = expf(L[g] - m);
    return m;
}

// exact log-domain value of one sum node: logsumexp_k(in0[i] + in1[j] + wlog[k*ldw])
// in0/in1 are read with a stride (global or shared memory)
static __device__ __noinline__ float slow_logsumexp(const float* in0, const float* in1, int stride, int G,
                                             const float* wlog, int ldw) {
    float M = -INFINITY;
    for (int j = 0; j < G; ++j)
        for (int i = 0; i < G; ++i)
            M = fmaxf(M, in0[i * stride] + in1[j * stride] + wlog[(j * G + i) * ldw]);
    if (!(M > -INFINITY)) return M;
    float acc = 0.f;
    for (int j = 0; j < G; ++j)
        for (int i = 0; i < G; ++i)
            acc += expf(in0[i * stride] + in1[j * stride] + wlog[(j * G + i) * ldw] - M);
    return M + logf(acc);
}


// SMEM = the parameter block has been staged in shared memory (plain loads), else read-only global loads
template <int G, bool SMEM = false>
__device__ __forceinline__ void load_leaf_params(const float* __restrict__ lp, float (&mu)[GP_<G>::v],
                                                 float (&a)[GP_<G>::v], float (&b)[GP_<G>::v]) {
    constexpr int GP = GP_<G>::v;
    const float4* p4 = reinterpret_cast<const float4*>(lp);
#pragma unroll
    for (int v = 0; v < GP / 4; ++v) {
        float4 t = SMEM ? p4[v] : __ldg(p4 + v);
        mu[4 * v] = t.x; mu[4 * v + 1] = t.y; mu[4 * v + 2] = t.z; mu[4 * v + 3] = t.w;
        t = SMEM ? p4[GP / 4 + v] : __ldg(p4 + GP / 4 + v);
        a[4 * v] = t.x; a[4 * v + 1] = t.y; a[4 * v + 2] = t.z; a[4 * v + 3] = t.w;
        t = SMEM ? p4[2 * (GP / 4) + v] : __ldg(p4 + 2 * (GP / 4) + v);
        b[4 * v] = t.x; b[4 * v + 1] = t.y; b[4 * v + 2] = t.z; b[4 * v + 3] = t.w;
    }
}


// exact log-domain backward of one sum node (slow path of the node passes): adds the responsibilities to the two
// input vectors (stride gstride) and, atomically, to the gradient of the log weights
static __device__ __noinline__ void slow_sum_backward(const float* in0, const float* in1, int stride, int G,
                                               const float* wlog, int ldw, float sumv, float gs,
                                               float* g0, float* g1, int gstride, float* g_wlog) {
    for (int j = 0; j < G; ++j)
        for (int i = 0; i < G; ++i) {
            const int k = j * G + i;
            const float resp = gs * expf(in0[i * stride] + in1[j * stride] + wlog[k * ldw] - sumv);
            g0[i * gstride] += resp;
            g1[j * gstride] += resp;
            atomicAdd(g_wlog + k * ldw, resp);
        }
}

