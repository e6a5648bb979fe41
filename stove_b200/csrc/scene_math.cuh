// Bilinear-sampling helpers shared by the glimpse / mask kernels (scene.cu) and the fused scene-likelihood
// kernels (scene_ll.cu).  Coordinates: x walks the LAST image axis (length B), y the
// second-to-last (length A); align_corners selects torch-1.0.1 semantics of F.affine_grid / F.grid_sample
// (model/video_prediction/supair.py:272-275, 321-341).
#pragma once
#include "common.cuh"

struct SceneDims {
    int O, C, A, B, pa, pb, align;
};
constexpr int SCENE_MAXC = 4;       // scene_bwd keeps one column-sum register per 32 columns: B <= 128

__device__ __forceinline__ float base_coord(int k, int n_out, int align) {
    if (align) return (n_out > 1) ? (2.f * k) / (float)(n_out - 1) - 1.f : -1.f;
    return (2.f * k + 1.f) / (float)n_out - 1.f;
}
__device__ __forceinline__ float unnorm(float g, int L, int align) {
    return align ? (g + 1.f) * 0.5f * (float)(L - 1) : ((g + 1.f) * (float)L - 1.f) * 0.5f;
}
__device__ __forceinline__ float unnorm_slope(int L, int align) {
    return align ? 0.5f * (float)(L - 1) : 0.5f * (float)L;
}
// bilinear sample of an all-ones row of length L with zero padding, and its derivative
__device__ __forceinline__ void tent(float p, int L, float& val, float& der) {
    const float x0 = floorf(p);
    const float f = p - x0;
    const float in0 = (x0 >= 0.f && x0 <= (float)(L - 1)) ? 1.f : 0.f;
    const float in1 = (x0 + 1.f >= 0.f && x0 + 1.f <= (float)(L - 1)) ? 1.f : 0.f;
    val = (1.f - f) * in0 + f * in1;
    der = in1 - in0;
}

struct Corner {
    int y0, x0;
    float fy, fx;
    bool oky0, oky1, okx0, okx1;
};
__device__ __forceinline__ Corner corners(float py, float px, int A, int B) {
    Corner c;
    const float fy0 = floorf(py), fx0 = floorf(px);
    c.fy = py - fy0;
    c.fx = px - fx0;
    // clamp before the int conversion so far-away boxes cannot overflow
    c.y0 = (int)fminf(fmaxf(fy0, -2.f), (float)A);
    c.x0 = (int)fminf(fmaxf(fx0, -2.f), (float)B);
    c.oky0 = c.y0 >= 0 && c.y0 < A;
    c.oky1 = c.y0 + 1 >= 0 && c.y0 + 1 < A;
    c.okx0 = c.x0 >= 0 && c.x0 < B;
    c.okx1 = c.x0 + 1 >= 0 && c.x0 + 1 < B;
    return c;
}
// value and d/dpy, d/dpx of the bilinear sample of `im` (A x B, zero padded); if INVERT the
// sampled image is (1 - im)
template <bool INVERT>
__device__ __forceinline__ void bilinear(const float* im, int B, const Corner& c, float& val, float& dy,
                                         float& dx) {
    float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
    if (c.oky0 && c.okx0) v00 = INVERT ? 1.f - im[c.y0 * B + c.x0] : im[c.y0 * B + c.x0];
    if (c.oky0 && c.okx1) v01 = INVERT ? 1.f - im[c.y0 * B + c.x0 + 1] : im[c.y0 * B + c.x0 + 1];
    if (c.oky1 && c.okx0) v10 = INVERT ? 1.f - im[(c.y0 + 1) * B + c.x0] : im[(c.y0 + 1) * B + c.x0];
    if (c.oky1 && c.okx1) v11 = INVERT ? 1.f - im[(c.y0 + 1) * B + c.x0 + 1] : im[(c.y0 + 1) * B + c.x0 + 1];
    const float top = v00 + c.fx * (v01 - v00), bot = v10 + c.fx * (v11 - v10);
    val = top + c.fy * (bot - top);
    dy = bot - top;
    dx = (1.f - c.fy) * (v01 - v00) + c.fy * (v11 - v10);
}

