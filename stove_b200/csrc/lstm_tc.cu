// Recognition-LSTM step on the 5th-generation tensor cores (sm_100a): one kernel per LSTM step that
// does the gate GEMM with tcgen05.mma (kind::tf32, accumulator in TMEM, operands brought in by TMA
// with the 128-byte swizzle) and applies the LSTM cell in the epilogue, straight out of TMEM.
// Replaces, per step, a library GEMM + the cell kernel of csrc/glue.cu (model/video_prediction/
// encoder.py:50-51 = nn.LSTM fed the same frame num_obj times).
//
// Accuracy: the GEMM is a 3xTF32 product.  Both operands arrive K-concatenated -- A = [hi | hi | lo],
// B = [hi | lo | hi] along K (csrc/glue.cu split kernels, and this kernel's own epilogue for h) -- so
// ONE TF32 GEMM over 3K yields hi*hi + hi*lo + lo*hi with fp32 accumulation: fp32-level accuracy
// (the hi parts are TF32-exact, the dropped lo*lo term is 2^-22 relative).
//
// Tiling: CTA = 128 rows x 32 hidden units x all 4 gates.  The B tile is four TMA boxes of 32 weight
// rows (gate g, hidden j0 .. j0+31), so accumulator column c = 32 g + j: after tcgen05.ld one thread
// owns, for its row, i/f/g/o of the same hidden unit and the cell update is thread-local.
//   warp 0     : TMA producer (one elected lane), 6- or 4-stage ring of (A 16 KB + B 16 KB); also fetches the
//                epilogue's inputs (gates of the input GEMM, previous cell state) behind the first tiles
//   warp 1     : MMA issuer (one elected lane): 4 x tcgen05.mma 128x128x8 per stage, tcgen05.commit
//                releases the stage / signals the epilogue
//   warps 2..9 : epilogue, warp w reads TMEM lanes 32 (w % 4) .. (two warps per quarter, half the columns each); results are staged in swizzled shared-memory
//                tiles (the operand ring is free by then) and leave with TMA stores
// Grid = ceil(n / 128) x H / 32 (16 x 8 = 128 CTAs for 2048 frames, H = 256): one wave on 148 SMs.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace lt {
constexpr int BM = 128, BH = 32, BN = 4 * BH, BK = 32, EPI_WARPS = 8, THREADS = 64 + 32 * EPI_WARPS;
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int GATE_BYTES = BH * BK * 4;
constexpr uint32_t TMEM_COLS = 128;
// tcgen05 instruction descriptor, kind::tf32: D = f32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2),
// both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

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
// shared-memory matrix descriptor: K-major tile, rows of 128 bytes, 128-byte swizzle, 8-row groups
// 1024 bytes apart (SBO), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
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
// the epilogue is instruction bound (one warp per scheduler): exp through MUFU.EX2 (__expf, ~2 ulp) and a fast
// reciprocal; tanh(x) = 2 sigmoid(2x) - 1 keeps the absolute error ~2e-7, far inside the 3e-5 parity bound
__device__ __forceinline__ float fsigmoid(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
__device__ __forceinline__ float ftanh(float v) { return 2.0f * fsigmoid(2.0f * v) - 1.0f; }
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ void ld8(const float* p, float* v) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float* v) {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

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

struct CellMaps {
    CUtensorMap A, B;          // GEMM operands
    CUtensorMap add, cprev;    // loads: gates of the input GEMM [n][4H], previous cell state [n][H]
    CUtensorMap act, gx, c, h, hcol, hrow;    // stores
};
struct CellFwd {
    int H;
    int64_t n;
    const float* bias;        // [4H] | null: the addend then comes through maps.add
    int has_cprev, has_gx, has_hsplit;
    int debug;                // timing experiments (STOVE_LSTM_TC_DEBUG): 1 = no TMA stores, 2 = one k-block, 4 = no epilogue
};

constexpr int TILE_BYTES = BM * BH * 4;      // one [128 rows][32 floats] epilogue tile
// Shared-memory plan.  The mainloop is bound by the bytes it keeps in flight (ncu: tensor pipe 25 % busy,
// L2 22 %, 403 MB through the crossbar for the input GEMM), so the operand ring is as deep as the SM allows:
//   PREFETCH = false (step 0: the addend is the bias, no previous cell state): 6 stages = 192 KB; all
//     epilogue tiles live in the ring once the accumulator is complete;
//   PREFETCH = true (later steps): 4 stages + 80 KB behind the ring that receive the gates of the input GEMM
//     and the previous cell state while the (short, K = 3H) mainloop runs.
template <bool PREFETCH>
struct Plan {
    static constexpr int STAGES = PREFETCH ? 4 : 6;
    static constexpr int RING = STAGES * STAGE_BYTES;
    static constexpr int OFF_GX = 0, OFF_HF = 4 * TILE_BYTES, OFF_HHI = OFF_HF + TILE_BYTES, OFF_HLO = OFF_HHI + TILE_BYTES;
    static constexpr int OFF_G = PREFETCH ? RING : OFF_HLO + TILE_BYTES, OFF_C = OFF_G + 4 * TILE_BYTES;
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
                const uint32_t sa = tiles + s * STAGE_BYTES, sb = sa + A_BYTES;
                tma_load_2d(sa, &maps.A, kb * BK, m0, full0 + 8 * s);
#pragma unroll
                for (int g = 0; g < 4; ++g) tma_load_2d(sb + g * GATE_BYTES, &maps.B, kb * BK, g * H + j0, full0 + 8 * s);
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
                const uint32_t sa = tiles + s * STAGE_BYTES, sb = sa + A_BYTES;
                const uint64_t ad = smem_desc(sa), bd = smem_desc(sb);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k)       // 8 TF32 = 32 bytes per MMA: +2 in the (addr >> 4) field
                    mma_tf32(tmem, ad + 2 * k, bd + 2 * k, (uint32_t)((kb | k) != 0));
                mma_commit(empty0 + 8 * s);            // implies tcgen05.fence::before_thread_sync
            }
            mma_commit(accum);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;                        // the TMEM lane quarter this warp may read
        const int r = q * 32 + lane;                   // row of the tile
        const int jc0 = ((warp - 2) >> 2) * (BH / 8 / (EPI_WARPS / 4));   // two warps per quarter split the columns
        mbar_wait(accum, 0);                           // all MMAs done: accumulator complete, operand ring free
        if (!(p.debug & 4)) {
        if (wait_e) mbar_wait(ebar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int jc = jc0; jc < jc0 + BH / 8 / (EPI_WARPS / 4); ++jc) {
            float v[4][8];
#pragma unroll
            for (int g = 0; g < 4; ++g) tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BH + jc * 8), v[g]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
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
                if (p.has_hsplit) {
                    const float4 hi = make_float4(tf32_hi(h.x), tf32_hi(h.y), tf32_hi(h.z), tf32_hi(h.w));
                    sts4(tiles + OFF_HHI + o, hi);
                    sts4(tiles + OFF_HLO + o, make_float4(h.x - hi.x, h.y - hi.y, h.z - hi.z, h.w - hi.w));
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to TMA
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");  // the epilogue warps
        if (warp == 2 && lane == 0 && !(p.debug & 1)) {
#pragma unroll
            for (int g = 0; g < 4; ++g) tma_store_2d(&maps.act, tiles + OFF_G + g * TILE_BYTES, g * H + j0, m0);
            tma_store_2d(&maps.c, tiles + OFF_C, j0, m0);
            tma_store_2d(&maps.h, tiles + OFF_HF, j0, m0);
            if (p.has_hsplit) {
                tma_store_2d(&maps.hcol, tiles + OFF_HHI, j0, m0);
                tma_store_2d(&maps.hcol, tiles + OFF_HHI, H + j0, m0);
                tma_store_2d(&maps.hcol, tiles + OFF_HLO, 2 * H + j0, m0);
                tma_store_3d(&maps.hrow, tiles + OFF_HHI, j0, m0, 0);
                tma_store_3d(&maps.hrow, tiles + OFF_HLO, j0, m0, 1);
                tma_store_3d(&maps.hrow, tiles + OFF_HHI, j0, m0, 2);
            }
            if (p.has_gx) {
#pragma unroll
                for (int g = 0; g < 4; ++g) tma_store_2d(&maps.gx, tiles + OFF_GX + g * TILE_BYTES, g * H + j0, m0);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the reads
        }
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
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

// fp32 tensor [depth][rows][cols] (row stride ld floats, plane stride rows * ld), box = 1 x box_rows x 32 floats,
// 128-byte swizzle, out-of-range elements read as zero / are not written
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    int64_t depth = 1) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        stove_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return STOVE_ERR_CUDA;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)depth};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)rows * (cuuint64_t)ld * 4};
    const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, depth > 1 ? 3 : 2, const_cast<float*>(base), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        stove_set_error("cuTensorMapEncodeTiled failed (%d) for a %lld x %lld x %lld tensor, ld %lld", (int)r,
                        (long long)depth, (long long)rows, (long long)cols, (long long)ld);
        return STOVE_ERR_CUDA;
    }
    return STOVE_OK;
}
}  // namespace lt

extern "C" int stove_lstm_gemm_cell_fwd(int64_t n, int H, int64_t Kc, const float* A, const float* B,
                                        const float* addend, int addend_is_bias, const float* c_prev, float* gx_out,
                                        float* h_out, int64_t h_ld, float* c_out, float* act, float* h_col,
                                        float* h_row, void* stream) {
    using namespace lt;
    STOVE_CHECK_ARG(n >= 0 && H > 0 && H % BH == 0 && Kc > 0 && Kc % 4 == 0, "need H % 32 == 0 and Kc % 4 == 0");
    STOVE_CHECK_ARG(A && B && addend && h_out && c_out && act && h_ld >= H && h_ld % 4 == 0, "bad argument");
    STOVE_CHECK_ARG((h_col == nullptr) == (h_row == nullptr), "h_col and h_row go together");
    STOVE_CHECK_ARG((((uintptr_t)A | (uintptr_t)B | (uintptr_t)addend | (uintptr_t)c_prev | (uintptr_t)gx_out |
                      (uintptr_t)h_out | (uintptr_t)c_out | (uintptr_t)act | (uintptr_t)h_col | (uintptr_t)h_row) & 15) == 0,
                    "pointers must be 16-byte aligned");
    if (n == 0) return STOVE_OK;
    CellMaps m;
    memset(&m, 0, sizeof(m));
    const int64_t H4 = 4 * (int64_t)H;
    int rc = make_map(&m.A, A, n, Kc, Kc, BM);
    if (rc == STOVE_OK) rc = make_map(&m.B, B, H4, Kc, Kc, BH);
    if (rc == STOVE_OK) rc = make_map(&m.add, addend_is_bias ? act : addend, n, H4, H4, BM);
    if (rc == STOVE_OK) rc = make_map(&m.cprev, c_prev ? c_prev : c_out, n, H, H, BM);
    if (rc == STOVE_OK) rc = make_map(&m.act, act, n, H4, H4, BM);
    if (rc == STOVE_OK) rc = make_map(&m.gx, gx_out ? gx_out : act, n, H4, H4, BM);
    if (rc == STOVE_OK) rc = make_map(&m.c, c_out, n, H, H, BM);
    if (rc == STOVE_OK) rc = make_map(&m.h, h_out, n, H, h_ld, BM);
    if (rc == STOVE_OK) rc = make_map(&m.hcol, h_col ? h_col : c_out, n, h_col ? 3 * (int64_t)H : H, h_col ? 3 * (int64_t)H : H, BM);
    if (rc == STOVE_OK) rc = h_row ? make_map(&m.hrow, h_row, n, H, H, BM, 3) : make_map(&m.hrow, c_out, n, H, H, BM);
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
    p.has_cprev = c_prev != nullptr; p.has_gx = gx_out != nullptr; p.has_hsplit = h_col != nullptr;
    const dim3 grid((unsigned)((n + BM - 1) / BM), (unsigned)(H / BH));
    cudaStream_t s = (cudaStream_t)stream;
    static const int dbg = getenv("STOVE_LSTM_TC_DEBUG") ? atoi(getenv("STOVE_LSTM_TC_DEBUG")) : 0;
    p.debug = dbg;
    const int num_kb = (dbg & 2) ? 1 : (int)((Kc + BK - 1) / BK);
    if (prefetch) {
        STOVE_KERNEL(K_LSTM_GEMM_CELL_FWD, s, lstm_gemm_cell_fwd_kernel<true><<<grid, THREADS, Plan<true>::SMEM, s>>>(m, num_kb, p));
    } else {
        STOVE_KERNEL(K_LSTM_GEMM_CELL_FWD, s, lstm_gemm_cell_fwd_kernel<false><<<grid, THREADS, Plan<false>::SMEM, s>>>(m, num_kb, p));
    }
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}
