// Recognition LSTM on the 5th-generation tensor cores (sm_100a) -- model/video_prediction/encoder.py:50-51
// (nn.LSTM fed the same frame num_obj times), forward AND backward, no library GEMM.
//
// Accuracy: every contraction is a 3xTF32 product.  Each fp32 operand is split ONCE into a TF32-exact part
// `hi` (low 13 mantissa bits cleared) and the remainder `lo = x - hi`, stored as two planes [2][rows][K] of
// one buffer; a k-block's hi and lo tiles are brought in once each by TMA (3-D tensor maps, plane = the
// outermost coordinate) and the tensor cores run hi*hi + hi*lo + lo*hi into ONE fp32 accumulator in TMEM:
// fp32-level accuracy (the dropped lo*lo term is 2^-22 relative) at 8 operand bytes per element -- round 1
// delivered K-concatenated [hi|hi|lo] copies, 12 bytes per element, and was bound by that traffic.
//
// Kernels
//   lstm_gemm_cell_fwd : one LSTM step = gate GEMM + cell in the epilogue (CTA = 128 frames x 32 hidden units
//                        x 4 gates: accumulator column = 32 gate + unit, so the cell update is thread-local
//                        after tcgen05.ld).  Emits h as fp32, as (hi, lo) planes for the next step's GEMM and
//                        TRANSPOSED (hi, lo) planes for the W_hh gradient.
//   tc3_gemm<BN>       : D[M][N] (+ split-K parts) = A B^T, A [2][M][K], B [2][N][K]: the hidden-state
//                        gradient (gate gradient x W_hh) and both weight gradients (contraction over the
//                        frames: operands are the transposed planes the cell kernels write).
//   lstm_cell_bwd_t    : gate/state backward; writes the gate gradient as (hi, lo) planes, row-major (left
//                        operand of the hidden-state GEMM) and transposed through a shared-memory tile
//                        (left operand of the weight-gradient GEMMs), the running sum over steps (what W_ih
//                        and the biases see) and per-block column sums for the bias gradient.
//   split_planes       : fp32 matrix -> (hi, lo) planes, optionally transposed.
//   sum_parts          : fixed-order sum of split-K partial products / bias partial sums.
// Warp roles of the tensor-core kernels: warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one
// elected lane: 12 x tcgen05.mma 128 x N x 8 per k-block, tcgen05.commit releases the stage), warps 2..9 =
// epilogue (TMEM -> registers -> 128-byte-swizzled shared-memory tiles in the drained operand ring -> TMA stores).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace lt {
constexpr int BM = 128, BH = 32, BK = 32, EPI_WARPS = 8, THREADS = 64 + 32 * EPI_WARPS;
constexpr int TILE_BYTES = BM * BK * 4;      // [128 rows][32 floats]: one operand plane of a k-block, one epilogue tile
constexpr int GATE_BYTES = BH * BK * 4;
// tcgen05 instruction descriptor, kind::tf32: D = f32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2),
// both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t idesc_for(int n_cols) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// shared-memory matrix descriptor: K-major tile, rows of 128 bytes, 128-byte swizzle, 8-row groups
// 1024 bytes apart (SBO), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// one k-block (32 floats of K) of the 3xTF32 product.  TWO accumulators: hi*hi goes to `tmem_big`, the cross
// terms lo*hi + hi*lo (2^-11 of the product) to `tmem_small`; the epilogue adds them in fp32.  The tensor cores
// align addends by truncation, a bias of ~0.5 ulp of the ACCUMULATOR per MMA step: keeping the small terms out
// of the big accumulator cuts the biased steps on it to a third and makes their own truncation 2^-11 smaller
// (measured on the nine-step recurrence of the multiball parity test: LSTM gradient error 5.7e-4 -> see
// profiles/r02_parity_*.jsonl).
__device__ __forceinline__ void mma_kblock_3x(uint32_t tmem_big, uint32_t tmem_small, uint32_t a_hi, uint32_t a_lo,
                                              uint32_t b_hi, uint32_t b_lo, uint32_t idesc, bool first) {
    const uint64_t ah = smem_desc(a_hi), al = smem_desc(a_lo), bh = smem_desc(b_hi), bl = smem_desc(b_lo);
#pragma unroll
    for (int k = 0; k < BK / 8; ++k) {         // 8 TF32 = 32 bytes per MMA: +2 in the (addr >> 4) field
        const uint32_t acc = (uint32_t)(!first || k != 0);
        mma_tf32(tmem_small, al + 2 * k, bh + 2 * k, idesc, acc);
        mma_tf32(tmem_small, ah + 2 * k, bl + 2 * k, idesc, 1u);
        mma_tf32(tmem_big, ah + 2 * k, bh + 2 * k, idesc, acc);
    }
}

__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// accumulator read-back: big + small
__device__ __forceinline__ void tmem_ld8_sum(uint32_t t_big, uint32_t t_small, float* v) {
    float w[8];
    tmem_ld8(t_big, v);
    tmem_ld8(t_small, w);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += w[i];
}
// the epilogue is instruction bound (one warp per scheduler): exp through MUFU.EX2 (__expf, ~2 ulp) and a fast
// reciprocal; tanh(x) = 2 sigmoid(2x) - 1 keeps the absolute error ~2e-7, far inside the 3e-5 parity bound
__device__ __forceinline__ float fsigmoid(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float ftanh(float v) { return 2.0f * fsigmoid(2.0f * v) - 1.0f; }
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// epilogue I/O goes through shared memory in the TMA 128-byte swizzle: thread = row reads / writes the
// 16-byte chunk c of its 128-byte row at chunk position c ^ (row & 7) -- conflict-free per quarter warp --
// and whole tiles move with TMA (coalesced, rows beyond n clipped / zero-filled by the hardware)
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }
__device__ __forceinline__ float4 lds4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// ------------------------------------------------------------------------------------
// forward step: gate GEMM + LSTM cell
// ------------------------------------------------------------------------------------
struct CellMaps {
    CUtensorMap A, B;          // GEMM operands, (hi, lo) planes
    CUtensorMap add, cprev;    // loads: gates of the input GEMM [n][4H], previous cell state [n][H]
    CUtensorMap act, gx, c, h, hpl;    // stores; hpl = (hi, lo) planes of h [2][n][H]
};
struct CellFwd {
    int H;
    int64_t n;
    const float* bias;        // [4H] | null: the addend then comes through maps.add
    float* hT;                // transposed (hi, lo) planes of h, [2][H][row stride ldT] | null; columns n .. spanT-1 are written as zero
    int64_t ldT, spanT, hT_plane;
    int has_cprev, has_gx, has_hpl;
};

constexpr int STAGE_BYTES = 4 * TILE_BYTES;     // A hi, A lo, B hi, B lo (B = four 32-row gate boxes per plane)
// Shared-memory plan:
//   PREFETCH = false (step 0: the addend is the bias, no previous cell state): 3 stages = 192 KB; all
//     epilogue tiles live in the ring once the accumulator is complete;
//   PREFETCH = true (later steps): 2 stages + 80 KB behind the ring that receive the gates of the input GEMM
//     and the previous cell state while the (short, K = H) mainloop runs.
template <bool PREFETCH>
struct Plan {
    static constexpr int STAGES = PREFETCH ? 2 : 3;
    static constexpr int RING = STAGES * STAGE_BYTES;
    static constexpr int OFF_HF = 0, OFF_HHI = TILE_BYTES, OFF_HLO = 2 * TILE_BYTES, OFF_GX = 3 * TILE_BYTES;
    static constexpr int OFF_G = PREFETCH ? RING : 7 * TILE_BYTES, OFF_C = OFF_G + 4 * TILE_BYTES;
    static constexpr int OFF_BAR = PREFETCH ? OFF_C + TILE_BYTES : RING;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static_assert(OFF_HLO + TILE_BYTES <= RING, "epilogue tiles must fit in the operand ring");
    static_assert(PREFETCH || OFF_C + TILE_BYTES <= RING, "epilogue tiles must fit in the operand ring");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

template <bool PREFETCH>
__global__ void __launch_bounds__(THREADS, 1)
lstm_gemm_cell_fwd_kernel(const __grid_constant__ CellMaps maps, int num_kb, CellFwd p) {
    using P_ = Plan<PREFETCH>;
    constexpr int STAGES = P_::STAGES, OFF_G = P_::OFF_G, OFF_C = P_::OFF_C, OFF_BAR = P_::OFF_BAR, OFF_GX = P_::OFF_GX,
                  OFF_HF = P_::OFF_HF, OFF_HHI = P_::OFF_HHI, OFF_HLO = P_::OFF_HLO;
    constexpr uint32_t TMEM_COLS = 256, IDESC = idesc_for(128);      // accumulators: big at column 0, small at 128
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);      // swizzle-128B tiles need 1024-byte alignment
    const uint32_t tiles = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, accum = empty0 + 8 * STAGES, ebar = accum + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, j0 = blockIdx.y * BH;
    const int H = p.H;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.A) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.B) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(accum, 1);
        mbar_init(ebar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const bool wait_e = PREFETCH && ((p.bias == nullptr) || p.has_cprev);

    if (warp == 0) {
        if (lane == 0) {
            // what the epilogue adds comes in behind the first operand tiles and lands long before it is needed
            if (wait_e) {
                mbar_expect_tx(ebar, (p.bias == nullptr ? 4 * TILE_BYTES : 0) + (p.has_cprev ? TILE_BYTES : 0));
                if (p.bias == nullptr) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) tma_load_2d(tiles + OFF_G + g * TILE_BYTES, &maps.add, g * H + j0, m0, ebar);
                }
                if (p.has_cprev) tma_load_2d(tiles + OFF_C, &maps.cprev, j0, m0, ebar);
            }
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
                const uint32_t sa = tiles + s * STAGE_BYTES, sb = sa + 2 * TILE_BYTES;
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
                    tma_load_3d(sa + pl * TILE_BYTES, &maps.A, kb * BK, m0, pl, full0 + 8 * s);
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        tma_load_3d(sb + pl * TILE_BYTES + g * GATE_BYTES, &maps.B, kb * BK, g * H + j0, pl, full0 + 8 * s);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
                mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = tiles + s * STAGE_BYTES;
                mma_kblock_3x(tmem, tmem + 128, sa, sa + TILE_BYTES, sa + 2 * TILE_BYTES, sa + 3 * TILE_BYTES, IDESC, kb == 0);
                mma_commit(empty0 + 8 * s);            // implies tcgen05.fence::before_thread_sync
            }
            mma_commit(accum);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;                        // the TMEM lane quarter this warp may read
        const int r = q * 32 + lane;                   // row of the tile
        const int jc0 = ((warp - 2) >> 2) * (BH / 8 / (EPI_WARPS / 4));   // two warps per quarter split the columns
        const int64_t grow = (int64_t)m0 + r;
        mbar_wait(accum, 0);                           // all MMAs done: accumulator complete, operand ring free
        if (wait_e) mbar_wait(ebar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int jc = jc0; jc < jc0 + BH / 8 / (EPI_WARPS / 4); ++jc) {
            float v[4][8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t t = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BH + jc * 8);
                tmem_ld8_sum(t, t + 128, v[g]);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int ch = jc * 2 + half;
                const uint32_t o = swz(r, ch);
                float4 pre[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 a = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + g * H + j0) + ch)
                                            : lds4(tiles + OFF_G + g * TILE_BYTES + o);
                    pre[g] = make_float4(v[g][half * 4 + 0] + a.x, v[g][half * 4 + 1] + a.y, v[g][half * 4 + 2] + a.z,
                                         v[g][half * 4 + 3] + a.w);
                    if (p.has_gx) sts4(tiles + OFF_GX + g * TILE_BYTES + o, pre[g]);
                }
                const float4 cp = p.has_cprev ? lds4(tiles + OFF_C + o) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 ig, fg, gg, og, c, h;
#define LT_CELL(X)                                                                     \
    ig.X = fsigmoid(pre[0].X); fg.X = fsigmoid(pre[1].X); gg.X = ftanh(pre[2].X);      \
    og.X = fsigmoid(pre[3].X); c.X = fg.X * cp.X + ig.X * gg.X; h.X = og.X * ftanh(c.X);
                LT_CELL(x) LT_CELL(y) LT_CELL(z) LT_CELL(w)
#undef LT_CELL
                sts4(tiles + OFF_G + 0 * TILE_BYTES + o, ig);
                sts4(tiles + OFF_G + 1 * TILE_BYTES + o, fg);
                sts4(tiles + OFF_G + 2 * TILE_BYTES + o, gg);
                sts4(tiles + OFF_G + 3 * TILE_BYTES + o, og);
                sts4(tiles + OFF_C + o, c);
                sts4(tiles + OFF_HF + o, h);
                if (p.has_hpl) {
                    const float4 hi = make_float4(tf32_hi(h.x), tf32_hi(h.y), tf32_hi(h.z), tf32_hi(h.w));
                    const float4 lo = make_float4(h.x - hi.x, h.y - hi.y, h.z - hi.z, h.w - hi.w);
                    sts4(tiles + OFF_HHI + o, hi);
                    sts4(tiles + OFF_HLO + o, lo);
                    if (p.hT && grow < p.spanT) {
                        // transposed planes [unit][frame]: lane = frame, so each store is one coalesced 128-byte row
                        const bool live = grow < p.n;
                        float* t0 = p.hT + (int64_t)(j0 + ch * 4) * p.ldT + grow;
                        float* t1 = t0 + p.hT_plane;
                        t0[0] = live ? hi.x : 0.f; t0[p.ldT] = live ? hi.y : 0.f;
                        t0[2 * p.ldT] = live ? hi.z : 0.f; t0[3 * p.ldT] = live ? hi.w : 0.f;
                        t1[0] = live ? lo.x : 0.f; t1[p.ldT] = live ? lo.y : 0.f;
                        t1[2 * p.ldT] = live ? lo.z : 0.f; t1[3 * p.ldT] = live ? lo.w : 0.f;
                    }
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to TMA
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");  // the epilogue warps
        if (warp == 2 && lane == 0) {
#pragma unroll
            for (int g = 0; g < 4; ++g) tma_store_2d(&maps.act, tiles + OFF_G + g * TILE_BYTES, g * H + j0, m0);
            tma_store_2d(&maps.c, tiles + OFF_C, j0, m0);
            tma_store_2d(&maps.h, tiles + OFF_HF, j0, m0);
            if (p.has_hpl) {
                tma_store_3d(&maps.hpl, tiles + OFF_HHI, j0, m0, 0);
                tma_store_3d(&maps.hpl, tiles + OFF_HLO, j0, m0, 1);
            }
            if (p.has_gx) {
#pragma unroll
                for (int g = 0; g < 4; ++g) tma_store_2d(&maps.gx, tiles + OFF_GX + g * TILE_BYTES, g * H + j0, m0);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the reads
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------
// D (+ split-K parts) = A B^T, 3xTF32
// ------------------------------------------------------------------------------------
struct GemmMaps {
    CUtensorMap A, B, D;       // A [2][M][K], B [2][N][K] (hi, lo) planes; D [parts][M][N]
};
template <int BN>
struct GemmPlan {
    static constexpr int B_BYTES = BN * BK * 4;
    static constexpr int STAGE = 2 * TILE_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (192 * 1024) / STAGE;       // BN = 64: 4, 128: 3, 256: 2
    static constexpr int RING = STAGES * STAGE;
    static constexpr int OFF_BAR = RING;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static_assert((BN / 32) * TILE_BYTES <= RING, "epilogue tiles must fit in the operand ring");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
tc3_gemm_kernel(const __grid_constant__ GemmMaps maps, int num_kb, int kb_per) {
    using P_ = GemmPlan<BN>;
    constexpr int STAGES = P_::STAGES, STAGE = P_::STAGE, B_BYTES = P_::B_BYTES, OFF_BAR = P_::OFF_BAR;
    constexpr uint32_t TMEM_COLS = 2 * BN, IDESC = idesc_for(BN);      // accumulators: big at column 0, small at BN
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const uint32_t tiles = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, accum = empty0 + 8 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN, z = blockIdx.z;
    const int kb0 = z * kb_per, kb1 = min(num_kb, kb0 + kb_per);     // host guarantees kb0 < kb1

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.A) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.B) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.D) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = kb0; kb < kb1; ++kb) {
                const int it = kb - kb0, s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                mbar_expect_tx(full0 + 8 * s, STAGE);
                const uint32_t sa = tiles + s * STAGE, sb = sa + 2 * TILE_BYTES;
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
                    tma_load_3d(sa + pl * TILE_BYTES, &maps.A, kb * BK, m0, pl, full0 + 8 * s);
                    tma_load_3d(sb + pl * B_BYTES, &maps.B, kb * BK, n0, pl, full0 + 8 * s);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            for (int kb = kb0; kb < kb1; ++kb) {
                const int it = kb - kb0, s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = tiles + s * STAGE, sb = sa + 2 * TILE_BYTES;
                mma_kblock_3x(tmem, tmem + BN, sa, sa + TILE_BYTES, sb, sb + B_BYTES, IDESC, it == 0);
                mma_commit(empty0 + 8 * s);
            }
            mma_commit(accum);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        constexpr int CH_PER_WARP = BN / 8 / (EPI_WARPS / 4);      // 8-column chunks per warp (two warps per quarter)
        const int c0 = ((warp - 2) >> 2) * CH_PER_WARP;
        mbar_wait(accum, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c8 = c0; c8 < c0 + CH_PER_WARP; c8 += 2) {
            float v[2][8];
            const uint32_t t = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c8 * 8);
            tmem_ld8_sum(t, t + BN, v[0]);
            if (CH_PER_WARP > 1) tmem_ld8_sum(t + 8, t + BN + 8, v[1]);
#pragma unroll
            for (int u = 0; u < (CH_PER_WARP > 1 ? 2 : 1); ++u) {
                const int col = (c8 + u) * 8;                      // column of the tile
                const uint32_t base = tiles + (uint32_t)(col >> 5) * TILE_BYTES;
                const int ch = (col & 31) >> 2;
                sts4(base + swz(r, ch), make_float4(v[u][0], v[u][1], v[u][2], v[u][3]));
                sts4(base + swz(r, ch + 1), make_float4(v[u][4], v[u][5], v[u][6], v[u][7]));
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        if (warp == 2 && lane == 0) {
#pragma unroll
            for (int t = 0; t < BN / 32; ++t) tma_store_3d(&maps.D, tiles + t * TILE_BYTES, n0 + 32 * t, m0, z);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------
// element-wise companions
// ------------------------------------------------------------------------------------
// x [rows][cols] (row stride ldx) -> pl [2][rows][cols] and / or plT [2][cols][ldT] (columns rows .. ldT-1 zero).
// Block (32, 8) owns a 32 x 32 tile; the transposed planes leave through a padded shared-memory tile.
__global__ void split_planes_kernel(int64_t rows, int cols, const float* __restrict__ x, int64_t ldx,
                                    float* __restrict__ pl, int64_t pl_plane, float* __restrict__ plT, int64_t ldT,
                                    int64_t spanT, int64_t plT_plane) {
    __shared__ float sh[2][32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t r0 = (int64_t)blockIdx.y * 32;
    const int c0 = blockIdx.x * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rr = ty + 8 * i;
        const int64_t r = r0 + rr;
        const int c = c0 + tx;
        float hi = 0.f, lo = 0.f;
        if (r < rows && c < cols) {
            const float v = x[r * ldx + c];
            hi = tf32_hi(v);
            lo = v - hi;
            if (pl) {
                pl[r * cols + c] = hi;
                pl[pl_plane + r * cols + c] = lo;
            }
        }
        sh[0][rr][tx] = hi;
        sh[1][rr][tx] = lo;
    }
    if (!plT) return;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int cc = ty + 8 * i;
        const int c = c0 + cc;
        const int64_t r = r0 + tx;
        if (c < cols && r < spanT) {
            plT[(int64_t)c * ldT + r] = sh[0][tx][cc];
            plT[plT_plane + (int64_t)c * ldT + r] = sh[1][tx][cc];
        }
    }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 sub4(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 hi4(float4 a) { return make_float4(tf32_hi(a.x), tf32_hi(a.y), tf32_hi(a.z), tf32_hi(a.w)); }

// Gate/state backward of one LSTM step.  A block of 256 threads owns 32 frames x 32 hidden units x 4 gates.
//   g_h = g_h_a (strided slice of the stacked gradient) + split-K parts of the hidden-state GEMM
//   g_pl  [2][n][4H]   : (hi, lo) planes of this step's gate gradient (left operand of the hidden-state GEMM) | null
//   gT    [2][4H][ldT] : transposed planes (left operand of a weight-gradient GEMM) of this step's gate gradient,
//                        or, with emit_acc, of the sum over steps; ldT = row stride, columns n .. spanT-1 are
//                        written as zero (a step's block of a buffer stacked along the columns) | null
//   g_acc [n][4H]      : running sum over steps (acc_mode 0: start, 1: add); not written back when emit_acc
//   bias_part [gridDim.y][4H] : column sums of the (accumulated) gate gradient over this block's frames | null
__global__ void __launch_bounds__(256)
lstm_cell_bwd_t_kernel(int64_t n, int H, const float* __restrict__ act, const float* __restrict__ c_prev,
                       const float* __restrict__ c_out, const float* __restrict__ g_h_a, int64_t g_h_a_ld,
                       const float* __restrict__ g_h_b, int g_h_b_parts, const float* __restrict__ g_c,
                       float* __restrict__ g_pl, float* __restrict__ gT, int64_t ldT, int64_t spanT, int64_t gT_plane,
                       float* __restrict__ g_acc, int acc_mode, int emit_acc, float* __restrict__ bias_part,
                       float* __restrict__ g_c_prev) {
    __shared__ float sh[2][4][32][33];
    const int tid = threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.y * 32;
    const int k0 = blockIdx.x * 32;
    const int64_t H4 = 4 * (int64_t)H;
    {
        // phase 1: thread = (frame, four consecutive hidden units): 16-byte loads / stores, a warp covers four
        // whole 128-byte rows of every operand
        const int quad = tid & 7, bb = tid >> 3;
        const int k = k0 + 4 * quad;
        const int64_t b = b0 + bb;
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < n) {
            const int64_t e = b * H + k;
            const float* pa = act + b * H4 + k;
            const float4 ig = ld4(pa), fg = ld4(pa + H), gg = ld4(pa + 2 * H), og = ld4(pa + 3 * H);
            const float4 co = ld4(c_out + e);
            float4 gh = ld4(g_h_a + b * g_h_a_ld + k);
            if (g_h_b)
                for (int part = 0; part < g_h_b_parts; ++part) gh = add4(gh, ld4(g_h_b + (int64_t)part * n * H + e));
            const float4 gcn = g_c ? ld4(g_c + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 cp = c_prev ? ld4(c_prev + e) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 g[4], gcp;
#define LT_BWD(X)                                                           \
    {                                                                       \
        const float tc = tanhf(co.X);                                       \
        const float gc = gcn.X + gh.X * og.X * (1.f - tc * tc);             \
        g[0].X = gc * gg.X * ig.X * (1.f - ig.X);                           \
        g[1].X = gc * cp.X * fg.X * (1.f - fg.X);                           \
        g[2].X = gc * ig.X * (1.f - gg.X * gg.X);                           \
        g[3].X = gh.X * tc * og.X * (1.f - og.X);                           \
        gcp.X = gc * fg.X;                                                  \
    }
            LT_BWD(x) LT_BWD(y) LT_BWD(z) LT_BWD(w)
#undef LT_BWD
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t o = b * H4 + q * H + k;
                const float4 a = acc_mode ? add4(ld4(g_acc + o), g[q]) : g[q];
                if (!emit_acc) st4(g_acc + o, a);
                if (g_pl) {
                    const float4 hi = hi4(g[q]);
                    st4(g_pl + o, hi);
                    st4(g_pl + n * H4 + o, sub4(g[q], hi));
                }
                v[q] = emit_acc ? a : g[q];
            }
            if (g_c_prev) st4(g_c_prev + e, gcp);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 hi = hi4(v[q]), lo = sub4(v[q], hi);
            float* s0 = &sh[0][q][bb][4 * quad];
            float* s1 = &sh[1][q][bb][4 * quad];
            s0[0] = hi.x; s0[1] = hi.y; s0[2] = hi.z; s0[3] = hi.w;
            s1[0] = lo.x; s1[1] = lo.y; s1[2] = lo.z; s1[3] = lo.w;
        }
    }
    if (!gT && !bias_part) return;
    __syncthreads();
    // phase 2: lane = frame, warp = hidden unit (+ 8 i): transposed planes as coalesced 128-byte rows
    const int lane = tid & 31, wp = tid >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int kk = wp + 8 * i;
        const int64_t b = b0 + lane;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float vh = sh[0][q][lane][kk], vl = sh[1][q][lane][kk];
            if (gT && b < spanT) {
                const int64_t o = (int64_t)(q * H + k0 + kk) * ldT + b;
                gT[o] = vh;
                gT[gT_plane + o] = vl;
            }
            if (bias_part) {
                const float s = warp_sum(vh + vl);         // hi + lo is the fp32 value again, exactly
                if (lane == 0) bias_part[(int64_t)blockIdx.y * H4 + q * H + k0 + kk] = s;
            }
        }
    }
}

// out[i] = sum_p parts[p * stride + i], fixed order
__global__ void sum_parts_kernel(int64_t n4, int parts, int64_t stride4, const float4* __restrict__ in,
                                 float4* __restrict__ out, float scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a = in[i];
#pragma unroll 8
    for (int p = 1; p < parts; ++p) {          // independent loads: unrolled so they are in flight together
        const float4 b = __ldg(in + (int64_t)p * stride4 + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;      // scale = 1 is exact
    out[i] = a;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}

// fp32 tensor [depth][rows][cols] (row stride ld floats, plane stride `plane` floats; 0 = rows * ld),
// box = 1 x box_rows x 32 floats, 128-byte swizzle, out-of-range elements read as zero / are not written.
// depth == 0: a 2-D map; depth >= 1: a 3-D map (the kernels address planes / split-K parts by the third coordinate)
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    int64_t depth = 0, int64_t plane = 0) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        stove_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return STOVE_ERR_CUDA;
    }
    if (plane == 0) plane = rows * ld;
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(depth > 0 ? depth : 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)plane * 4};
    const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, depth > 0 ? 3 : 2, const_cast<float*>(base), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        stove_set_error("cuTensorMapEncodeTiled failed (%d) for a %lld x %lld x %lld tensor, ld %lld, plane %lld", (int)r,
                        (long long)depth, (long long)rows, (long long)cols, (long long)ld, (long long)plane);
        return STOVE_ERR_CUDA;
    }
    return STOVE_OK;
}

template <int BN>
static int launch_gemm(const GemmMaps& m, int64_t M, int64_t Nn, int num_kb, int parts, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        STOVE_CUDA(cudaFuncSetAttribute(tc3_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmPlan<BN>::SMEM));
        attr_set = true;
    }
    const int kb_per = (num_kb + parts - 1) / parts;
    const dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((Nn + BN - 1) / BN), (unsigned)((num_kb + kb_per - 1) / kb_per));
    STOVE_KERNEL(K_TC3_GEMM, s, tc3_gemm_kernel<BN><<<grid, THREADS, GemmPlan<BN>::SMEM, s>>>(m, num_kb, kb_per));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}
}  // namespace lt

extern "C" int stove_lstm_gemm_cell_fwd(int64_t n, int H, int64_t K, const float* A_pl, const float* B_pl,
                                        const float* addend, int addend_is_bias, const float* c_prev, float* gx_out,
                                        float* h_out, int64_t h_ld, float* c_out, float* act, float* h_pl,
                                        float* hT_pl, int64_t ldT, int64_t spanT, int64_t hT_plane, void* stream) {
    using namespace lt;
    STOVE_CHECK_ARG(n >= 0 && H > 0 && H % BH == 0 && K > 0 && K % 4 == 0, "need H % 32 == 0 and K % 4 == 0");
    STOVE_CHECK_ARG(A_pl && B_pl && addend && h_out && c_out && act && h_ld >= H && h_ld % 4 == 0, "bad argument");
    STOVE_CHECK_ARG(!hT_pl || (h_pl && spanT >= n && ldT >= spanT && hT_plane >= (int64_t)H * ldT),
                    "hT_pl needs h_pl, ldT >= spanT >= n and room for a plane");
    STOVE_CHECK_ARG((((uintptr_t)A_pl | (uintptr_t)B_pl | (uintptr_t)addend | (uintptr_t)c_prev | (uintptr_t)gx_out |
                      (uintptr_t)h_out | (uintptr_t)c_out | (uintptr_t)act | (uintptr_t)h_pl) & 15) == 0,
                    "pointers must be 16-byte aligned");
    if (n == 0) return STOVE_OK;
    CellMaps m;
    memset(&m, 0, sizeof(m));
    const int64_t H4 = 4 * (int64_t)H;
    int rc = make_map(&m.A, A_pl, n, K, K, BM, 2);
    if (rc == STOVE_OK) rc = make_map(&m.B, B_pl, H4, K, K, BH, 2);
    if (rc == STOVE_OK) rc = make_map(&m.add, addend_is_bias ? act : addend, n, H4, H4, BM);
    if (rc == STOVE_OK) rc = make_map(&m.cprev, c_prev ? c_prev : c_out, n, H, H, BM);
    if (rc == STOVE_OK) rc = make_map(&m.act, act, n, H4, H4, BM);
    if (rc == STOVE_OK) rc = make_map(&m.gx, gx_out ? gx_out : act, n, H4, H4, BM);
    if (rc == STOVE_OK) rc = make_map(&m.c, c_out, n, H, H, BM);
    if (rc == STOVE_OK) rc = make_map(&m.h, h_out, n, H, h_ld, BM);
    if (rc == STOVE_OK) rc = h_pl ? make_map(&m.hpl, h_pl, n, H, H, BM, 2) : make_map(&m.hpl, c_out, n, H, H, BM);
    if (rc != STOVE_OK) return rc;
    const bool prefetch = !addend_is_bias || c_prev != nullptr;
    static bool attr_set = false;
    if (!attr_set) {
        STOVE_CUDA(cudaFuncSetAttribute(lstm_gemm_cell_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan<true>::SMEM));
        STOVE_CUDA(cudaFuncSetAttribute(lstm_gemm_cell_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan<false>::SMEM));
        attr_set = true;
    }
    CellFwd p;
    p.H = H; p.n = n; p.bias = addend_is_bias ? addend : nullptr;
    p.hT = hT_pl; p.ldT = ldT; p.spanT = spanT; p.hT_plane = hT_plane;
    p.has_cprev = c_prev != nullptr; p.has_gx = gx_out != nullptr; p.has_hpl = h_pl != nullptr;
    const int64_t span = hT_pl && spanT > n ? spanT : n;       // the pad columns of the transposed planes are zeroed too
    const dim3 grid((unsigned)((span + BM - 1) / BM), (unsigned)(H / BH));
    cudaStream_t s = (cudaStream_t)stream;
    const int num_kb = (int)((K + BK - 1) / BK);
    if (prefetch) {
        STOVE_KERNEL(K_LSTM_GEMM_CELL_FWD, s, lstm_gemm_cell_fwd_kernel<true><<<grid, THREADS, Plan<true>::SMEM, s>>>(m, num_kb, p));
    } else {
        STOVE_KERNEL(K_LSTM_GEMM_CELL_FWD, s, lstm_gemm_cell_fwd_kernel<false><<<grid, THREADS, Plan<false>::SMEM, s>>>(m, num_kb, p));
    }
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" int stove_tc3_gemm(int64_t M, int64_t Nn, int64_t K, const float* A_pl, int64_t lda, int64_t a_plane,
                              const float* B_pl, int64_t ldb, int64_t b_plane, float* D, int64_t ldd, int parts,
                              int64_t part_stride, int bn, void* stream) {
    using namespace lt;
    STOVE_CHECK_ARG(M > 0 && Nn > 0 && K > 0 && A_pl && B_pl && D && parts >= 1, "bad argument");
    STOVE_CHECK_ARG(lda >= K && ldb >= K && ldd >= Nn && lda % 4 == 0 && ldb % 4 == 0 && ldd % 4 == 0 &&
                        a_plane % 4 == 0 && b_plane % 4 == 0 && part_stride % 4 == 0,
                    "leading dimensions and plane strides must be multiples of 4 floats");
    STOVE_CHECK_ARG((((uintptr_t)A_pl | (uintptr_t)B_pl | (uintptr_t)D) & 15) == 0, "pointers must be 16-byte aligned");
    const int num_kb = (int)((K + BK - 1) / BK);
    if (parts > num_kb) parts = num_kb;
    const int kb_per = (num_kb + parts - 1) / parts;
    parts = (num_kb + kb_per - 1) / kb_per;                // no empty split
    STOVE_CHECK_ARG(parts == 1 || part_stride >= M * ldd, "part_stride too small");
    // tile width: 128 x 256 tiles take a quarter less operand traffic than 128 x 128 ones (these GEMMs are bound
    // by operand delivery from L2) but halve the CTA count; the caller asks for them where N >= 256 and split-K
    // still fills the machine (the weight gradients)
    STOVE_CHECK_ARG(bn == 0 || bn == 64 || bn == 128 || bn == 256, "bn must be 0 (auto), 64, 128 or 256");
    const int BN = bn ? bn : (Nn > 64 ? 128 : 64);
    GemmMaps m;
    memset(&m, 0, sizeof(m));
    int rc = make_map(&m.A, A_pl, M, K, lda, BM, 2, a_plane);
    if (rc == STOVE_OK) rc = make_map(&m.B, B_pl, Nn, K, ldb, BN, 2, b_plane);
    if (rc == STOVE_OK) rc = make_map(&m.D, D, M, Nn, ldd, BM, parts, parts > 1 ? part_stride : M * ldd);
    if (rc != STOVE_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (BN == 256) return launch_gemm<256>(m, M, Nn, num_kb, parts, s);
    return BN == 128 ? launch_gemm<128>(m, M, Nn, num_kb, parts, s) : launch_gemm<64>(m, M, Nn, num_kb, parts, s);
}

extern "C" int stove_tc3_gemm_parts(int64_t M, int64_t Nn, int64_t K, int want) {
    using namespace lt;
    if (M <= 0 || Nn <= 0 || K <= 0) return 1;
    const int num_kb = (int)((K + BK - 1) / BK);
    int parts = want < 1 ? 1 : want;
    if (parts > num_kb) parts = num_kb;
    const int kb_per = (num_kb + parts - 1) / parts;
    return (num_kb + kb_per - 1) / kb_per;
}

extern "C" int stove_split_planes(int64_t rows, int cols, const float* x, int64_t ldx, float* pl, float* plT,
                                  int64_t ldT, void* stream) {
    STOVE_CHECK_ARG(rows >= 0 && cols > 0 && x && ldx >= cols && (pl || plT), "bad argument");
    STOVE_CHECK_ARG(!plT || ldT >= rows, "ldT must cover the rows");
    if (rows == 0) return STOVE_OK;
    const int64_t span = plT && ldT > rows ? ldT : rows;
    const dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((span + 31) / 32));
    STOVE_KERNEL(K_SPLIT_TF32, (cudaStream_t)stream, lt::split_planes_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
        rows, cols, x, ldx, pl, rows * (int64_t)cols, plT, ldT, ldT, (int64_t)cols * ldT));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" int stove_lstm_cell_bwd_t(int64_t n, int H, const float* act, const float* c_prev, const float* c_out,
                                     const float* g_h_a, int64_t g_h_a_ld, const float* g_h_b, int g_h_b_parts,
                                     const float* g_c, float* g_pl, float* gT_pl, int64_t ldT, int64_t spanT,
                                     int64_t gT_plane, float* g_acc, int acc_mode, int emit_acc, float* bias_part, float* g_c_prev,
                                     void* stream) {
    STOVE_CHECK_ARG(n >= 0 && H > 0 && H % 32 == 0 && act && c_out && g_h_a && g_h_a_ld >= H, "bad argument");
    STOVE_CHECK_ARG(g_acc || (emit_acc && !acc_mode), "g_acc is required unless this is the only step");
    STOVE_CHECK_ARG(g_h_a_ld % 4 == 0 && (((uintptr_t)act | (uintptr_t)c_prev | (uintptr_t)c_out | (uintptr_t)g_h_a |
                                            (uintptr_t)g_h_b | (uintptr_t)g_c | (uintptr_t)g_pl | (uintptr_t)g_acc |
                                            (uintptr_t)g_c_prev) & 15) == 0,
                    "pointers must be 16-byte aligned and g_h_a_ld a multiple of 4");
    STOVE_CHECK_ARG(!g_h_b || g_h_b_parts >= 1, "g_h_b_parts must be >= 1");
    STOVE_CHECK_ARG(!gT_pl || (spanT >= n && ldT >= spanT && gT_plane >= 4 * (int64_t)H * ldT),
                    "gT_pl needs ldT >= spanT >= n and room for a plane");
    if (n == 0) return STOVE_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t span = gT_pl && spanT > n ? spanT : n;
    const dim3 grid((unsigned)(H / 32), (unsigned)((span + 31) / 32));
    STOVE_KERNEL(K_LSTM_CELL_BWD, s, lt::lstm_cell_bwd_t_kernel<<<grid, 256, 0, s>>>(
        n, H, act, c_prev, c_out, g_h_a, g_h_a_ld, g_h_b, g_h_b_parts, g_c, g_pl, gT_pl, ldT, spanT, gT_plane, g_acc, acc_mode,
        emit_acc, bias_part, g_c_prev));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" int stove_sum_parts(int64_t numel, int parts, int64_t stride, const float* in, float* out, float scale,
                               void* stream) {
    STOVE_CHECK_ARG(numel >= 0 && numel % 4 == 0 && parts >= 1 && stride % 4 == 0 && in && out, "need numel % 4 == 0 and stride % 4 == 0");
    STOVE_CHECK_ARG((((uintptr_t)in | (uintptr_t)out) & 15) == 0, "pointers must be 16-byte aligned");
    if (numel == 0) return STOVE_OK;
    cudaStream_t s = (cudaStream_t)stream;
    STOVE_KERNEL(K_SUM_PARTS, s, lt::sum_parts_kernel<<<(unsigned)((numel / 4 + 255) / 256), 256, 0, s>>>(
        numel / 4, parts, stride / 4, (const float4*)in, (float4*)out, scale));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}
