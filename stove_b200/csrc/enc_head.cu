// Output head of the recognition network (model/video_prediction/encoder.py:53-56):
//     zp = fc2(sigmoid(fc1(h)))          h [R][K] (K = 256 LSTM features), fc1 K -> J (50), fc2 J -> P (8)
// One kernel forward, two + a reduction backward, instead of ~14 library launches (two SIMT-fp32 GEMMs
// with split-K reductions, bias/sigmoid element-wise kernels and their backward GEMMs / column sums):
// the layer is 79 M MAC per step -- a few microseconds of FP32 work -- and sits on the critical chain
// of both passes.
//   head_fwd      CTA = 32 rows; W1 transposed in shared memory (row stride 65: conflict-free for the
//                 transposing store and for lane = hidden unit reads), warp = 4 rows, lane = hidden j, j + 32
//   head_bwd_data g_pre = (g_out W2) h (1 - h);  g_x = g_pre W1;  lane = input feature k + 32 i
//   head_bwd_par  weight / bias gradients: thread = k keeps the J partial sums of g_W1[:, k] in registers
//                 over the CTA's rows; one slab per CTA, summed in a fixed order by head_bwd_reduce.
#include "common.cuh"

namespace eh {
constexpr int KMAX = 256, JP = 64, PMAX = 16, ROWS = 32, THREADS = 256, LDW = JP + 1, PAR_ROWS = 48;
constexpr int FROWS = 48, FRW = FROWS / 8;        // forward: rows per CTA (one CTA per SM for 6144 rows) / per warp

__device__ __forceinline__ float sigm(float v) { return 1.0f / (1.0f + expf(-v)); }
// staging without a register round trip: every copy of the CTA is in flight at once
__device__ __forceinline__ void cp4(float* smem_dst, const float* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(THREADS)
head_fwd_kernel(int64_t R, int K, int J, int P, const float* __restrict__ x, const float* __restrict__ w1,
                const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                float* __restrict__ hidden, float* __restrict__ out) {
    extern __shared__ float sm[];
    float* W1T = sm;                       // [K][LDW]
    float* XS = W1T + K * LDW;             // [ROWS][K]
    float* W2S = XS + FROWS * K;            // [P][JP]
    float* B1S = W2S + PMAX * JP;          // [JP]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t r0 = (int64_t)blockIdx.x * FROWS;
    // W1 [J][K] -> W1T [k][LDW] (4-byte async copies, coalesced along k); rows j >= J stay zero
    for (int i = tid; i < K * (JP - J); i += THREADS) W1T[(i / (JP - J)) * LDW + J + i % (JP - J)] = 0.f;
    {
        const int sh = 31 - __clz(K);                  // shift / mask indexing when K is a power of two
        if ((K & (K - 1)) == 0) {
            for (int i = tid; i < J * K; i += THREADS) cp4(W1T + (i & (K - 1)) * LDW + (i >> sh), w1 + i);
        } else {
            for (int i = tid; i < J * K; i += THREADS) cp4(W1T + (i % K) * LDW + i / K, w1 + i);
        }
    }
    for (int i = tid; i < PMAX * JP; i += THREADS) {
        const int o = i / JP, j = i - o * JP;
        W2S[i] = (o < P && j < J) ? __ldg(w2 + o * J + j) : 0.f;
    }
    if (tid < JP) B1S[tid] = tid < J ? __ldg(b1 + tid) : 0.f;
    const int K4 = K >> 2;
    for (int i = tid; i < FROWS * K4; i += THREADS) {
        const int r = i / K4, c = i - r * K4;
        if (r0 + r < R) cp16(XS + r * K + 4 * c, x + (r0 + r) * K + 4 * c);
        else reinterpret_cast<float4*>(XS + r * K)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_wait_all();
    __syncthreads();
    float acc[FRW][2];
#pragma unroll
    for (int r = 0; r < FRW; ++r) acc[r][0] = acc[r][1] = 0.f;
    const float* xr = XS + warp * FRW * K;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 xv[FRW];
#pragma unroll
        for (int r = 0; r < FRW; ++r) xv[r] = *reinterpret_cast<const float4*>(xr + r * K + k);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float wa = W1T[(k + kk) * LDW + lane], wb = W1T[(k + kk) * LDW + lane + 32];
#pragma unroll
            for (int r = 0; r < FRW; ++r) {
                const float xs = kk == 0 ? xv[r].x : kk == 1 ? xv[r].y : kk == 2 ? xv[r].z : xv[r].w;
                acc[r][0] = fmaf(xs, wa, acc[r][0]);
                acc[r][1] = fmaf(xs, wb, acc[r][1]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < FRW; ++r) {
        const int64_t row = r0 + warp * FRW + r;
        const float ha = lane < J ? sigm(acc[r][0] + B1S[lane]) : 0.f;
        const float hb = lane + 32 < J ? sigm(acc[r][1] + B1S[lane + 32]) : 0.f;
        if (row < R) {
            if (lane < J) hidden[row * J + lane] = ha;
            if (lane + 32 < J) hidden[row * J + lane + 32] = hb;
        }
        float mine = 0.f;
        for (int o = 0; o < P; ++o) {
            const float s = warp_sum(ha * W2S[o * JP + lane] + hb * W2S[o * JP + lane + 32]);
            if (lane == o) mine = s + __ldg(b2 + o);
        }
        if (row < R && lane < P) out[row * P + lane] = mine;
    }
}

// g_pre [R][JP] (padded with zeros), g_x [R][K]
__global__ void __launch_bounds__(THREADS)
head_bwd_data_kernel(int64_t R, int K, int J, int P, const float* __restrict__ w1, const float* __restrict__ w2,
                     const float* __restrict__ hidden, const float* __restrict__ g_out,
                     float* __restrict__ g_pre, float* __restrict__ g_x) {
    extern __shared__ float sm[];
    float* W1S = sm;                       // [J][K] as in global memory
    float* W2S = W1S + JP * KMAX;          // [P][JP]
    float* GP = W2S + PMAX * JP;           // [JP][ROWS]: g_pre of the CTA's rows, j-major
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t r0 = (int64_t)blockIdx.x * ROWS;
    for (int i = tid; i < J * (K >> 2); i += THREADS) cp16(W1S + 4 * i, w1 + 4 * i);
    for (int i = tid; i < PMAX * JP; i += THREADS) {
        const int o = i / JP, j = i - o * JP;
        W2S[i] = (o < P && j < J) ? __ldg(w2 + o * J + j) : 0.f;
    }
    cp_wait_all();
    __syncthreads();
    // g_pre of the warp's four rows for hidden units lane and lane + 32, kept in registers and written to the
    // j-major tile as ONE float4 per unit: with scalar stores all 32 lanes hit one bank (row stride 32: 30 % of the
    // kernel's shared-memory wavefronts were conflicts); the 4-row group of unit j sits at group position
    // warp ^ (j & 7), so the eight lanes of a store phase touch eight different bank groups
    float ga4[4], gb4[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int rl = warp * 4 + r;
        const int64_t row = r0 + rl;
        float ga = 0.f, gb = 0.f;
        if (row < R) {
            float sa = 0.f, sb = 0.f;
            for (int o = 0; o < P; ++o) {
                const float g = __ldg(g_out + row * P + o);
                sa = fmaf(g, W2S[o * JP + lane], sa);
                sb = fmaf(g, W2S[o * JP + lane + 32], sb);
            }
            if (lane < J) {
                const float ha = hidden[row * J + lane];
                ga = sa * ha * (1.f - ha);
            }
            if (lane + 32 < J) {
                const float hb = hidden[row * J + lane + 32];
                gb = sb * hb * (1.f - hb);
            }
            g_pre[row * JP + lane] = ga;
            g_pre[row * JP + lane + 32] = gb;
        }
        ga4[r] = ga;
        gb4[r] = gb;
    }
    *reinterpret_cast<float4*>(GP + lane * ROWS + ((warp ^ (lane & 7)) << 2)) = make_float4(ga4[0], ga4[1], ga4[2], ga4[3]);
    *reinterpret_cast<float4*>(GP + (lane + 32) * ROWS + ((warp ^ (lane & 7)) << 2)) = make_float4(gb4[0], gb4[1], gb4[2], gb4[3]);
    __syncwarp();                                           // a warp only reads back its own four rows
    const int NI = K >> 5;                                  // input features per lane (k = lane + 32 i)
    float acc[4][KMAX / 32];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < KMAX / 32; ++i) acc[r][i] = 0.f;
    for (int j = 0; j < J; ++j) {
        const float4 g = *reinterpret_cast<const float4*>(GP + j * ROWS + ((warp ^ (j & 7)) << 2));
#pragma unroll
        for (int i = 0; i < KMAX / 32; ++i) {
            if (i < NI) {
                const float w = W1S[j * K + lane + 32 * i];
                acc[0][i] = fmaf(g.x, w, acc[0][i]);
                acc[1][i] = fmaf(g.y, w, acc[1][i]);
                acc[2][i] = fmaf(g.z, w, acc[2][i]);
                acc[3][i] = fmaf(g.w, w, acc[3][i]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t row = r0 + warp * 4 + r;
        if (row < R) {
#pragma unroll
            for (int i = 0; i < KMAX / 32; ++i)
                if (i < NI) g_x[row * K + lane + 32 * i] = acc[r][i];
        }
    }
}

// slab per CTA: [J][K] g_w1 | [JP] g_b1 | [P][JP] g_w2 | [PMAX] g_b2
__host__ __device__ inline int slab_floats(int K, int J) { return J * K + JP + PMAX * JP + PMAX; }

__global__ void __launch_bounds__(THREADS)
head_bwd_par_kernel(int64_t R, int K, int J, int P, int64_t rows_per_cta, const float* __restrict__ x,
                    const float* __restrict__ hidden, const float* __restrict__ g_out,
                    const float* __restrict__ g_pre, float* __restrict__ slabs) {
    extern __shared__ __align__(16) float sm[];
    float (*XS)[KMAX] = reinterpret_cast<float (*)[KMAX]>(sm);
    float (*GP)[JP] = reinterpret_cast<float (*)[JP]>(sm + PAR_ROWS * KMAX);
    float (*HS)[JP] = reinterpret_cast<float (*)[JP]>(sm + PAR_ROWS * (KMAX + JP));
    float (*GO)[PMAX] = reinterpret_cast<float (*)[PMAX]>(sm + PAR_ROWS * (KMAX + 2 * JP));
    const int tid = threadIdx.x;
    const int64_t lo = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t hi = lo + rows_per_cta < R ? lo + rows_per_cta : R;
    float acc[JP];                       // g_w1[j][k = tid]
#pragma unroll
    for (int j = 0; j < JP; ++j) acc[j] = 0.f;
    float accb1 = 0.f, accw2[PMAX];      // thread j < JP: g_b1[j], g_w2[:, j]
#pragma unroll
    for (int o = 0; o < PMAX; ++o) accw2[o] = 0.f;
    float accb2 = 0.f;                   // thread o < P: g_b2[o]
    const int K4 = K >> 2;
    for (int64_t base = lo; base < hi; base += PAR_ROWS) {
        const int nr = (int)(hi - base < PAR_ROWS ? hi - base : PAR_ROWS);
        __syncthreads();
        for (int i = tid; i < PAR_ROWS * K4; i += THREADS) {
            const int r = i / K4, c = i - r * K4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nr) v = __ldg(reinterpret_cast<const float4*>(x + (base + r) * K) + c);
            reinterpret_cast<float4*>(&XS[r][0])[c] = v;
        }
        for (int i = tid; i < PAR_ROWS * JP; i += THREADS) {
            const int r = i / JP, j = i - r * JP;
            GP[r][j] = r < nr ? g_pre[(base + r) * JP + j] : 0.f;
            HS[r][j] = (r < nr && j < J) ? hidden[(base + r) * J + j] : 0.f;
        }
        for (int i = tid; i < PAR_ROWS * PMAX; i += THREADS) {
            const int r = i / PMAX, o = i - r * PMAX;
            GO[r][o] = (r < nr && o < P) ? __ldg(g_out + (base + r) * P + o) : 0.f;
        }
        __syncthreads();
        if (tid < K) {
#pragma unroll 2
            for (int r = 0; r < PAR_ROWS; ++r) {
                const float xv = XS[r][tid];
#pragma unroll
                for (int j4 = 0; j4 < JP / 4; ++j4) {
                    const float4 g = *reinterpret_cast<const float4*>(&GP[r][j4 * 4]);
                    acc[j4 * 4 + 0] = fmaf(g.x, xv, acc[j4 * 4 + 0]);
                    acc[j4 * 4 + 1] = fmaf(g.y, xv, acc[j4 * 4 + 1]);
                    acc[j4 * 4 + 2] = fmaf(g.z, xv, acc[j4 * 4 + 2]);
                    acc[j4 * 4 + 3] = fmaf(g.w, xv, acc[j4 * 4 + 3]);
                }
            }
        }
        if (tid < JP) {
            for (int r = 0; r < PAR_ROWS; ++r) {
                accb1 += GP[r][tid];
                const float h = HS[r][tid];
#pragma unroll
                for (int o = 0; o < PMAX; ++o) accw2[o] = fmaf(GO[r][o], h, accw2[o]);
            }
        }
        if (tid >= JP && tid < JP + PMAX) {
            for (int r = 0; r < PAR_ROWS; ++r) accb2 += GO[r][tid - JP];
        }
    }
    float* slab = slabs + (int64_t)blockIdx.x * slab_floats(K, J);
    if (tid < K) {
#pragma unroll
        for (int j = 0; j < JP; ++j)
            if (j < J) slab[j * K + tid] = acc[j];
    }
    if (tid < JP) {
        slab[J * K + tid] = accb1;
#pragma unroll
        for (int o = 0; o < PMAX; ++o) slab[J * K + JP + o * JP + tid] = accw2[o];
    }
    if (tid >= JP && tid < JP + PMAX) slab[J * K + JP + PMAX * JP + (tid - JP)] = accb2;
}

// eight lanes per element: lane q sums slabs q, q + 8, ... (independent loads in flight), fixed-order butterfly
__global__ void head_bwd_reduce_kernel(int K, int J, int P, int nslab, const float* __restrict__ slabs,
                                       float* __restrict__ g_w1, float* __restrict__ g_b1,
                                       float* __restrict__ g_w2, float* __restrict__ g_b2) {
    const int sf = slab_floats(K, J);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 3, q = t & 7;
    float s = 0.f;
    if (i < sf) {
#pragma unroll 4
        for (int c = q; c < nslab; c += 8) s += slabs[(int64_t)c * sf + i];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (i >= sf || q != 0) return;
    if (i < J * K) {
        g_w1[i] = s;
    } else if (i < J * K + JP) {
        const int j = i - J * K;
        if (j < J) g_b1[j] = s;
    } else if (i < J * K + JP + PMAX * JP) {
        const int o = (i - J * K - JP) / JP, j = (i - J * K - JP) % JP;
        if (o < P && j < J) g_w2[o * J + j] = s;
    } else {
        const int o = i - J * K - JP - PMAX * JP;
        if (o < P) g_b2[o] = s;
    }
}

static int par_ctas(int64_t R) {
    int64_t c = (R + PAR_ROWS - 1) / PAR_ROWS;
    const int cap = stove_opt(OPT_HEAD_PAR_CTAS) > 0 ? stove_opt(OPT_HEAD_PAR_CTAS) : 148;
    return (int)(c < cap ? (c < 1 ? 1 : c) : cap);
}
static bool dims_ok(int K, int J, int P) { return K > 0 && K <= KMAX && K % 32 == 0 && J > 0 && J <= JP && P > 0 && P <= PMAX; }
}  // namespace eh

extern "C" int stove_enc_head_fwd(int64_t R, int K, int J, int P, const float* x, const float* w1, const float* b1,
                                  const float* w2, const float* b2, float* hidden, float* out, void* stream) {
    using namespace eh;
    STOVE_CHECK_ARG(R >= 0 && x && w1 && b1 && w2 && b2 && hidden && out, "bad argument");
    if (!dims_ok(K, J, P)) {
        stove_set_error("stove_enc_head_fwd: unsupported head %d -> %d -> %d (need K <= 256, K %% 32 == 0, J <= 64, P <= 16)", K, J, P);
        return STOVE_ERR_UNSUPPORTED;
    }
    STOVE_CHECK_ARG((((uintptr_t)x) & 15) == 0, "x must be 16-byte aligned");
    if (R == 0) return STOVE_OK;
    const size_t smem = (size_t)(K * LDW + FROWS * K + PMAX * JP + JP) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        STOVE_CUDA(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)((KMAX * LDW + FROWS * KMAX + PMAX * JP + JP) * sizeof(float))));
        attr = true;
    }
    cudaStream_t s = (cudaStream_t)stream;
    STOVE_KERNEL(K_ENC_HEAD_FWD, s, head_fwd_kernel<<<(unsigned)((R + FROWS - 1) / FROWS), THREADS, smem, s>>>(
        R, K, J, P, x, w1, b1, w2, b2, hidden, out));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" size_t stove_enc_head_bwd_workspace(int64_t R, int K, int J, int P) {
    using namespace eh;
    (void)P;
    return (size_t)(R * JP + (int64_t)par_ctas(R) * slab_floats(K, J)) * sizeof(float);
}

extern "C" int stove_enc_head_bwd_data(int64_t R, int K, int J, int P, const float* w1, const float* w2,
                                       const float* hidden, const float* g_out, float* g_x, float* ws, void* stream) {
    using namespace eh;
    STOVE_CHECK_ARG(R >= 0 && w1 && w2 && hidden && g_out && g_x && ws, "bad argument");
    if (!dims_ok(K, J, P)) {
        stove_set_error("stove_enc_head_bwd_data: unsupported head %d -> %d -> %d", K, J, P);
        return STOVE_ERR_UNSUPPORTED;
    }
    STOVE_CHECK_ARG((((uintptr_t)w1 | (uintptr_t)ws) & 15) == 0, "w1, ws must be 16-byte aligned");
    if (R == 0) return STOVE_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)(JP * KMAX + PMAX * JP + JP * ROWS) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        STOVE_CUDA(cudaFuncSetAttribute(head_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    STOVE_KERNEL(K_ENC_HEAD_BWD_DATA, s, head_bwd_data_kernel<<<(unsigned)((R + ROWS - 1) / ROWS), THREADS, smem, s>>>(
        R, K, J, P, w1, w2, hidden, g_out, ws, g_x));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" int stove_enc_head_bwd_params(int64_t R, int K, int J, int P, const float* x, const float* hidden,
                                         const float* g_out, float* g_w1, float* g_b1, float* g_w2, float* g_b2,
                                         float* ws, void* stream) {
    using namespace eh;
    STOVE_CHECK_ARG(R >= 0 && x && hidden && g_out && g_w1 && g_b1 && g_w2 && g_b2 && ws, "bad argument");
    if (!dims_ok(K, J, P)) {
        stove_set_error("stove_enc_head_bwd_params: unsupported head %d -> %d -> %d", K, J, P);
        return STOVE_ERR_UNSUPPORTED;
    }
    STOVE_CHECK_ARG((((uintptr_t)x | (uintptr_t)ws) & 15) == 0, "x, ws must be 16-byte aligned");
    cudaStream_t ps = (cudaStream_t)stream;
    if (R == 0) {
        STOVE_CUDA(cudaMemsetAsync(g_w1, 0, sizeof(float) * J * K, ps));
        STOVE_CUDA(cudaMemsetAsync(g_b1, 0, sizeof(float) * J, ps));
        STOVE_CUDA(cudaMemsetAsync(g_w2, 0, sizeof(float) * P * J, ps));
        STOVE_CUDA(cudaMemsetAsync(g_b2, 0, sizeof(float) * P, ps));
        return STOVE_OK;
    }
    const float* g_pre = ws;
    float* slabs = ws + R * JP;
    const int ctas = par_ctas(R);
    const int64_t rows_per_cta = ((R + ctas - 1) / ctas + PAR_ROWS - 1) / PAR_ROWS * PAR_ROWS;
    const int used = (int)((R + rows_per_cta - 1) / rows_per_cta);
    const size_t psmem = (size_t)PAR_ROWS * (KMAX + 2 * JP + PMAX) * sizeof(float);
    static bool pattr = false;
    if (!pattr) {
        STOVE_CUDA(cudaFuncSetAttribute(head_bwd_par_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        pattr = true;
    }
    STOVE_KERNEL(K_ENC_HEAD_BWD_PAR, ps, head_bwd_par_kernel<<<used, THREADS, psmem, ps>>>(
        R, K, J, P, rows_per_cta, x, hidden, g_out, g_pre, slabs));
    STOVE_LAUNCH_CHECK();
    const int sf = slab_floats(K, J);
    STOVE_KERNEL(K_ENC_HEAD_BWD_PAR, ps, head_bwd_reduce_kernel<<<(sf * 8 + 255) / 256, 256, 0, ps>>>(
        K, J, P, used, slabs, g_w1, g_b1, g_w2, g_b2));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}
