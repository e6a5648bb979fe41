// Glimpse extraction + sequential marginalisation masks, forward and backward (d/dz).
//
// Replaces Supair.patches_from_z (model/video_prediction/supair.py:241-276) and
// Supair.masks_from_z (:278-356): per object  F.affine_grid + F.grid_sample  (bilinear, zero
// padding) of the frame, of the inverted running background, and of a ones image pasted back
// with the inverse transform, followed by clamp.  The reference launches 4 grid ops per
// object and materialises an O-fold copy of every frame (supair.py:262-263); here one CTA
// owns one frame, keeps the frame and the running background in shared memory and walks the
// objects in order.  The paste of a ones image is separable -- tent(py) * tent(px) -- so it
// needs A + B evaluations instead of A * B bilinear samples.
//
// Coordinates: x walks the LAST image axis (length B), y the
// second-to-last (length A).  align_corners selects torch-1.0.1 semantics.
// Backward formulas: tests/kernel_spec.py (checked against autograd of the oracle).
#include "common.cuh"
#include "scene_math.cuh"

__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < nwarp; ++w) t += red[w];
    return t;
}

// sums of four values over the CTA in one pass (two barriers instead of eight); the totals are valid
// in thread 0 only
__device__ __forceinline__ float4 block_sum4_t0(float4 v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v.x = warp_sum(v.x); v.y = warp_sum(v.y); v.z = warp_sum(v.z); v.w = warp_sum(v.w);
    __syncthreads();
    if (lane == 0) reinterpret_cast<float4*>(red)[warp] = v;
    __syncthreads();
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0)
        for (int w = 0; w < nwarp; ++w) {
            const float4 r = reinterpret_cast<const float4*>(red)[w];
            t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
        }
    return t;
}

// tents of the paste of object (sx, sy, tx, ty): tX[v], tY[u] (+ derivatives if dX != null)
__device__ __forceinline__ void paste_tents(const SceneDims& d, float sx, float sy, float tx, float ty,
                                            float* tX, float* tY, float* dX, float* dY) {
    const float isx = 1.f / sx, isy = 1.f / sy, cx = -tx / sx, cy = -ty / sy;
    for (int k = threadIdx.x; k < d.A + d.B; k += blockDim.x) {
        float val, der;
        if (k < d.B) {
            tent(unnorm(isx * base_coord(k, d.B, d.align) + cx, d.B, d.align), d.B, val, der);
            tX[k] = val;
            if (dX) dX[k] = der;
        } else {
            const int u = k - d.B;
            tent(unnorm(isy * base_coord(u, d.A, d.align) + cy, d.A, d.align), d.A, val, der);
            tY[u] = val;
            if (dY) dY[u] = der;
        }
    }
}

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scene_fwd_kernel(SceneDims d, const float* __restrict__ img,
                                                        const float* __restrict__ z,
                                                        float* __restrict__ patches,
                                                        float* __restrict__ marg_patch,
                                                        float* __restrict__ marg_bg,
                                                        float* __restrict__ overlap) {
    extern __shared__ float smem[];
    const int AB = d.A * d.B, PP = d.pa * d.pb;
    float* ims = smem;                 // [C][AB]
    float* bg = ims + d.C * AB;        // [AB]
    float* tX = bg + AB;               // [B]
    float* tY = tX + d.B;              // [A]
    float* red = tY + d.A;             // [32]
    const int64_t f = blockIdx.x;
    const int tid = threadIdx.x;
    for (int i = tid; i < d.C * AB; i += blockDim.x) ims[i] = __ldg(img + f * d.C * AB + i);
    for (int i = tid; i < AB; i += blockDim.x) bg[i] = 0.f;
    __syncthreads();
    for (int o = 0; o < d.O; ++o) {
        const float* zo = z + (f * d.O + o) * 4;
        const float sx = __ldg(zo), sy = __ldg(zo + 1), tx = __ldg(zo + 2), ty = __ldg(zo + 3);
        float msum = 0.f;
        for (int idx = tid; idx < PP; idx += blockDim.x) {
            const int i = idx / d.pb, j = idx - i * d.pb;
            const float px = unnorm(sx * base_coord(j, d.pb, d.align) + tx, d.B, d.align);
            const float py = unnorm(sy * base_coord(i, d.pa, d.align) + ty, d.A, d.align);
            const Corner c = corners(py, px, d.A, d.B);
            float val, dy, dx;
            bilinear<true>(bg, d.B, c, val, dy, dx);
            const float mg = 1.f - val;
            msum += mg;
            const int64_t ob = (f * d.O + o) * d.C;
            for (int ch = 0; ch < d.C; ++ch) {
                bilinear<false>(ims + ch * AB, d.B, c, val, dy, dx);
                patches[(ob + ch) * PP + idx] = val;
                marg_patch[(ob + ch) * PP + idx] = mg;
            }
        }
        paste_tents(d, sx, sy, tx, ty, tX, tY, nullptr, nullptr);
        msum = block_sum(msum, red);      // also orders the bg reads / tent writes before the update
        if (tid == 0) overlap[f * d.O + o] = msum / (float)PP;
        for (int i = tid; i < AB; i += blockDim.x) {
            const int u = i / d.B, v = i - u * d.B;
            bg[i] = fminf(fmaxf(bg[i] + tY[u] * tX[v], 0.f), 1.f);
        }
        __syncthreads();
    }
    for (int i = tid; i < d.C * AB; i += blockDim.x) marg_bg[f * d.C * AB + i] = bg[i % AB];
}

// ------------------------------------------------------------------------------------
// backward (d/dz)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scene_bwd_kernel(SceneDims d, const float* __restrict__ img,
                                                        const float* __restrict__ z,
                                                        const float* __restrict__ g_patches,
                                                        const float* __restrict__ g_marg_patch,
                                                        const float* __restrict__ g_marg_bg,
                                                        const float* __restrict__ g_overlap,
                                                        float* __restrict__ g_z) {
    extern __shared__ __align__(16) float smem[];
    const int AB = d.A * d.B, PP = d.pa * d.pb;
    float* ims = smem;                  // [C][AB]
    float* bgs = ims + d.C * AB;        // [O][AB] background before object o
    float* Gb = bgs + d.O * AB;         // [AB] running gradient w.r.t. the background
    float* tX = Gb + AB;
    float* dX = tX + d.B;
    float* gpx = dX + d.B;
    float* tY = gpx + d.B;
    float* dY = tY + d.A;
    float* gpy = dY + d.A;
    float* colp = gpy + d.A;            // [nwarp][B] column sums per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    float* red = smem + ((((d.C + d.O + 1) * AB + 3 * (d.A + d.B) + nwarp * d.B) + 3) & ~3);   // [32], 16-byte aligned
    const int64_t f = blockIdx.x;
    const int tid = threadIdx.x;
    const int shB = (d.B & (d.B - 1)) == 0 ? __ffs(d.B) - 1 : -1;      // row index without a division when B = 2^k
    if ((AB & 3) == 0) {
        const float4* src = reinterpret_cast<const float4*>(img + f * d.C * AB);
        for (int i = tid; i < d.C * AB / 4; i += blockDim.x) reinterpret_cast<float4*>(ims)[i] = __ldg(src + i);
    } else {
        for (int i = tid; i < d.C * AB; i += blockDim.x) ims[i] = __ldg(img + f * d.C * AB + i);
    }
    for (int i = tid; i < AB; i += blockDim.x) bgs[i] = 0.f;
    __syncthreads();
    // replay the forward background states
    for (int o = 0; o + 1 < d.O; ++o) {
        const float* zo = z + (f * d.O + o) * 4;
        paste_tents(d, __ldg(zo), __ldg(zo + 1), __ldg(zo + 2), __ldg(zo + 3), tX, tY, nullptr, nullptr);
        __syncthreads();
        for (int i = tid; i < AB; i += blockDim.x) {
            const int u = shB >= 0 ? i >> shB : i / d.B, v = i - u * d.B;
            bgs[(o + 1) * AB + i] = fminf(fmaxf(bgs[o * AB + i] + tY[u] * tX[v], 0.f), 1.f);
        }
        __syncthreads();
    }
    for (int i = tid; i < AB; i += blockDim.x) {
        float g = 0.f;
        if (g_marg_bg)
            for (int ch = 0; ch < d.C; ++ch) g += __ldg(g_marg_bg + (f * d.C + ch) * AB + i);
        Gb[i] = g;
    }
    __syncthreads();
    const float kB = unnorm_slope(d.B, d.align), kA = unnorm_slope(d.A, d.align);
    for (int o = d.O - 1; o >= 0; --o) {
        const float* zo = z + (f * d.O + o) * 4;
        const float sx = __ldg(zo), sy = __ldg(zo + 1), tx = __ldg(zo + 2), ty = __ldg(zo + 3);
        const float* bgo = bgs + o * AB;
        paste_tents(d, sx, sy, tx, ty, tX, tY, dX, dY);
        __syncthreads();
        // One pass over the frame: clamp backward (pass where 0 <= bg + paste <= 1) and, with paste = tY[u] tX[v],
        // the row sums (-> d/d tY) and column sums (-> d/d tX) of the surviving gradient.  Warp = rows u, u + 8, ...;
        // lane = column: row sums by shuffle, column sums in registers, combined over the warps through `colp`.
        // (Round 1 ran three passes -- clamp, row sums, then column sums on 32 of the 256 threads -- and was
        // issue bound: 23 k warp instructions per frame, profiles/r01_ncu_full_spn_scene_v2.txt.)
        {
            float colacc[SCENE_MAXC];
#pragma unroll
            for (int c = 0; c < SCENE_MAXC; ++c) colacc[c] = 0.f;
            for (int u = warp; u < d.A; u += nwarp) {
                const float tyu = tY[u];
                float racc = 0.f;
#pragma unroll
                for (int c = 0; c < SCENE_MAXC; ++c) {
                    const int v = lane + 32 * c;
                    if (v < d.B) {
                        const int i = u * d.B + v;
                        const float txv = tX[v];
                        const float pre = bgo[i] + tyu * txv;
                        float g = Gb[i];
                        if (!(pre >= 0.f && pre <= 1.f)) {
                            g = 0.f;
                            Gb[i] = 0.f;
                        }
                        racc = fmaf(g, txv, racc);
                        colacc[c] = fmaf(g, tyu, colacc[c]);
                    }
                }
                racc = warp_sum(racc);
                if (lane == 0) gpy[u] = racc * dY[u];
            }
#pragma unroll
            for (int c = 0; c < SCENE_MAXC; ++c) {
                const int v = lane + 32 * c;
                if (v < d.B) colp[warp * d.B + v] = colacc[c];
            }
        }
        __syncthreads();
        for (int k = tid; k < d.B; k += blockDim.x) {
            float acc = 0.f;
            for (int w = 0; w < nwarp; ++w) acc += colp[w * d.B + k];
            gpx[k] = acc * dX[k];
        }
        __syncthreads();
        float gsx = 0.f, gsy = 0.f, gtx = 0.f, gty = 0.f;
        for (int k = tid; k < d.A + d.B; k += blockDim.x) {
            if (k < d.B) {
                const float xb = base_coord(k, d.B, d.align);
                gsx += gpx[k] * kB * (-(xb - tx) / (sx * sx));
                gtx += gpx[k] * kB * (-1.f / sx);
            } else {
                const int u = k - d.B;
                const float yb = base_coord(u, d.A, d.align);
                gsy += gpy[u] * kA * (-(yb - ty) / (sy * sy));
                gty += gpy[u] * kA * (-1.f / sy);
            }
        }
        // glimpse + mask sampling points
        const float gov = g_overlap ? __ldg(g_overlap + f * d.O + o) / (float)PP : 0.f;
        const int64_t ob = (f * d.O + o) * d.C;
        for (int idx = tid; idx < PP; idx += blockDim.x) {
            const int i = idx / d.pb, j = idx - i * d.pb;
            const float xb = base_coord(j, d.pb, d.align), yb = base_coord(i, d.pa, d.align);
            const float px = unnorm(sx * xb + tx, d.B, d.align), py = unnorm(sy * yb + ty, d.A, d.align);
            const Corner c = corners(py, px, d.A, d.B);
            float gm = gov;
            if (g_marg_patch)
                for (int ch = 0; ch < d.C; ++ch) gm += __ldg(g_marg_patch + (ob + ch) * PP + idx);
            float val, my, mx;
            bilinear<true>(bgo, d.B, c, val, my, mx);
            float dpx = -gm * mx, dpy = -gm * my;
            if (g_patches)
                for (int ch = 0; ch < d.C; ++ch) {
                    float qy, qx;
                    bilinear<false>(ims + ch * AB, d.B, c, val, qy, qx);
                    const float gp = __ldg(g_patches + (ob + ch) * PP + idx);
                    dpx = fmaf(gp, qx, dpx);
                    dpy = fmaf(gp, qy, dpy);
                }
            gsx += dpx * kB * xb;
            gtx += dpx * kB;
            gsy += dpy * kA * yb;
            gty += dpy * kA;
            // d marg / d bg_o = + bilinear weights
            if (gm != 0.f) {
                const float wy0 = 1.f - c.fy, wy1 = c.fy, wx0 = 1.f - c.fx, wx1 = c.fx;
                if (c.oky0 && c.okx0) atomicAdd(&Gb[c.y0 * d.B + c.x0], gm * wy0 * wx0);
                if (c.oky0 && c.okx1) atomicAdd(&Gb[c.y0 * d.B + c.x0 + 1], gm * wy0 * wx1);
                if (c.oky1 && c.okx0) atomicAdd(&Gb[(c.y0 + 1) * d.B + c.x0], gm * wy1 * wx0);
                if (c.oky1 && c.okx1) atomicAdd(&Gb[(c.y0 + 1) * d.B + c.x0 + 1], gm * wy1 * wx1);
            }
        }
        const float4 tot = block_sum4_t0(make_float4(gsx, gsy, gtx, gty), red);
        if (tid == 0) *reinterpret_cast<float4*>(g_z + (f * d.O + o) * 4) = tot;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
static int scene_check(int64_t F, int O, int C, int A, int B, int pa, int pb) {
    STOVE_CHECK_ARG(F >= 0 && O > 0 && C > 0 && A > 0 && B > 0 && pa > 0 && pb > 0, "bad sizes");
    return STOVE_OK;
}

extern "C" int stove_scene_fwd(int64_t F, int O, int C, int A, int B, int pa, int pb, int align_corners,
                               const float* img, const float* z, float* patches, float* marg_patch,
                               float* marg_bg, float* overlap, void* stream) {
    int rc = scene_check(F, O, C, A, B, pa, pb);
    if (rc) return rc;
    STOVE_CHECK_ARG(img && z && patches && marg_patch && marg_bg && overlap, "null pointer");
    if (F == 0) return STOVE_OK;
    SceneDims d{O, C, A, B, pa, pb, align_corners};
    const size_t smem = sizeof(float) * ((size_t)(C + 1) * A * B + A + B + 32);
    if (smem > 227 * 1024) {
        stove_set_error("stove_scene_fwd: frame too large for shared memory (%zu B)", smem);
        return STOVE_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        STOVE_CUDA(cudaFuncSetAttribute(scene_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    STOVE_KERNEL(K_SCENE_FWD, (cudaStream_t)stream, scene_fwd_kernel<<<(unsigned)F, 256, smem, (cudaStream_t)stream>>>(d, img, z, patches, marg_patch, marg_bg, overlap));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

extern "C" int stove_scene_bwd(int64_t F, int O, int C, int A, int B, int pa, int pb, int align_corners,
                               const float* img, const float* z, const float* g_patches,
                               const float* g_marg_patch, const float* g_marg_bg, const float* g_overlap,
                               float* g_z, void* stream) {
    int rc = scene_check(F, O, C, A, B, pa, pb);
    if (rc) return rc;
    STOVE_CHECK_ARG(img && z && g_z, "null pointer");
    if (F == 0) return STOVE_OK;
    SceneDims d{O, C, A, B, pa, pb, align_corners};
    const size_t smem = sizeof(float) * ((size_t)(C + O + 1) * A * B + 3 * (A + B) + 8 * B + 4 + 32);
    if (smem > 227 * 1024 || B > 32 * SCENE_MAXC) {
        stove_set_error("stove_scene_bwd: frame/objects too large (%zu B of shared memory, last axis %d > %d)", smem, B,
                        32 * SCENE_MAXC);
        return STOVE_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
        STOVE_CUDA(cudaFuncSetAttribute(scene_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    STOVE_KERNEL(K_SCENE_BWD, (cudaStream_t)stream, scene_bwd_kernel<<<(unsigned)F, 256, smem, (cudaStream_t)stream>>>(d, img, z, g_patches, g_marg_patch,
                                                                       g_marg_bg, g_overlap, g_z));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}

// ------------------------------------------------------------------------------------
// Renderer: frames from states (Supair.reconstruct_from_z, supair.py:425-501; the same paste that
// draw_balls of the simulator performs with analytic blobs, envs.py:310-339).
//   out[f] = clamp(bg + sum_o sample(patch[f, o], inverse affine of z[f, o]), 0, 1)
// where sample = F.grid_sample(patch, F.affine_grid(theta^-1, (A, B))), bilinear, zero padded.  The
// reference issues 2 grid ops + an add per object over an O-fold repeated canvas; here one thread owns one
// output pixel and walks the objects (patches are tiny and L1/L2 resident).  No gradient (visualisation,
// MCTS frame generation).
// ------------------------------------------------------------------------------------
__global__ void render_kernel(SceneDims d, int64_t F, const float* __restrict__ bg, int bg_per_frame,
                              const float* __restrict__ patches, int64_t patch_frame_stride,
                              const float* __restrict__ z, float* __restrict__ out) {
    const int64_t total = F * d.C * d.A * d.B;
    const int pp = d.pa * d.pb;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % d.B);
        const int u = (int)((i / d.B) % d.A);
        const int ch = (int)((i / ((int64_t)d.A * d.B)) % d.C);
        const int64_t f = i / ((int64_t)d.C * d.A * d.B);
        float acc = bg[(bg_per_frame ? f * d.C * d.A * d.B : 0) + (ch * d.A + u) * d.B + v];
        const float gx = base_coord(v, d.B, d.align), gy = base_coord(u, d.A, d.align);
        for (int o = 0; o < d.O; ++o) {
            const float4 zz = __ldg(reinterpret_cast<const float4*>(z + (f * d.O + o) * 4));
            // inverse transform (supair.py:218-239): [1/sx, 1/sy, -x/sx, -y/sy]
            const float px = unnorm(gx / zz.x - zz.z / zz.x, d.pb, d.align);
            const float py = unnorm(gy / zz.y - zz.w / zz.y, d.pa, d.align);
            const Corner c = corners(py, px, d.pa, d.pb);
            float val, dy, dx;
            bilinear<false>(patches + f * patch_frame_stride + (int64_t)(o * d.C + ch) * pp, d.pb, c, val, dy, dx);
            acc += val;
        }
        out[i] = fminf(fmaxf(acc, 0.f), 1.f);
    }
}

extern "C" int stove_render(int64_t F, int O, int C, int A, int B, int pa, int pb, int align_corners,
                            const float* bg, int bg_per_frame, const float* patches, int patches_per_frame,
                            const float* z, float* out, void* stream) {
    STOVE_CHECK_ARG(F >= 0 && O > 0 && C > 0 && A > 0 && B > 0 && pa > 0 && pb > 0 && bg && patches && z && out,
                    "bad argument");
    STOVE_CHECK_ARG(((uintptr_t)z & 15) == 0, "z must be 16-byte aligned");
    if (F == 0) return STOVE_OK;
    SceneDims d{O, C, A, B, pa, pb, align_corners};
    const int64_t total = F * C * A * B;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t s = (cudaStream_t)stream;
    STOVE_KERNEL(K_RENDER, s, render_kernel<<<blocks, 256, 0, s>>>(d, F, bg, bg_per_frame, patches,
                                                                  patches_per_frame ? (int64_t)O * C * pa * pb : 0, z, out));
    STOVE_LAUNCH_CHECK();
    return STOVE_OK;
}
