"""Plain-loop fp64 statements of the hand-derived backward passes the CUDA kernels
implement (one frame at a time, the way a CTA walks it).  Checked against autograd of the
oracle in test_kernel_spec.py; the kernels in stove_b200/csrc/scene.cu follow these formulas."""
import math

import torch


def _coords(n_out, L, scale, shift, align):
    """sample position (pixel units) of output index k along an axis of length L."""
    k = torch.arange(n_out, dtype=torch.float64)
    base = 2 * k / max(n_out - 1, 1) - 1 if align else (2 * k + 1) / n_out - 1
    g = scale * base + shift
    p = (g + 1) * (L - 1) / 2 if align else ((g + 1) * L - 1) / 2
    return base, p


def _unnorm_slope(L, align):
    return (L - 1) / 2 if align else L / 2


def _tent(p, L):
    x0 = torch.floor(p)
    f = p - x0
    in0 = ((x0 >= 0) & (x0 <= L - 1)).double()
    in1 = ((x0 + 1 >= 0) & (x0 + 1 <= L - 1)).double()
    return (1 - f) * in0 + f * in1, in1 - in0


def _bilinear(img, py, px):
    """img (A,B); returns value (len(py), len(px)), d/dpy, d/dpx and the corner lists."""
    A, B = img.shape
    y0, x0 = torch.floor(py).long(), torch.floor(px).long()
    fy, fx = py - y0, px - x0
    val = torch.zeros(len(py), len(px), dtype=torch.float64)
    dpy = torch.zeros_like(val)
    dpx = torch.zeros_like(val)
    for dy in (0, 1):
        for dx in (0, 1):
            yy, xx = y0 + dy, x0 + dx
            ok = ((yy >= 0) & (yy < A))[:, None] & ((xx >= 0) & (xx < B))[None, :]
            v = img[yy.clamp(0, A - 1)][:, xx.clamp(0, B - 1)] * ok
            wy = fy if dy else 1 - fy
            wx = fx if dx else 1 - fx
            val += v * wy[:, None] * wx[None, :]
            dpy += v * (1.0 if dy else -1.0) * wx[None, :]
            dpx += v * wy[:, None] * (1.0 if dx else -1.0)
    return val, dpy, dpx


def scene_frame_fwd(img, z, pa, pb, align=False):
    """img (C,A,B), z (O,4) -> patches (O,C,pa,pb), marg (O,pa,pb), bg (A,B), states."""
    C, A, B = img.shape
    O = z.shape[0]
    bg = torch.zeros(A, B, dtype=torch.float64)
    patches, margs, bgs, pre = [], [], [], []
    for o in range(O):
        sx, sy, tx, ty = z[o]
        _, px = _coords(pb, B, sx, tx, align)
        _, py = _coords(pa, A, sy, ty, align)
        patches.append(torch.stack([_bilinear(img[c], py, px)[0] for c in range(C)]))
        margs.append(1 - _bilinear(1 - bg, py, px)[0])
        _, ppx = _coords(B, B, 1 / sx, -tx / sx, align)
        _, ppy = _coords(A, A, 1 / sy, -ty / sy, align)
        paste = _tent(ppy, A)[0][:, None] * _tent(ppx, B)[0][None, :]
        bgs.append(bg)
        pre.append(bg + paste)
        bg = torch.clamp(bg + paste, 0, 1)
    return torch.stack(patches), torch.stack(margs), bg, bgs, pre


def scene_frame_bwd(img, z, pa, pb, g_patch, g_marg, g_bg, align=False):
    """g_patch (O,C,pa,pb), g_marg (O,pa,pb) (already channel-summed, overlap folded in),
    g_bg (A,B) -> g_z (O,4)."""
    C, A, B = img.shape
    O = z.shape[0]
    _, _, _, bgs, pre = scene_frame_fwd(img, z, pa, pb, align)
    G = g_bg.clone()
    gz = torch.zeros(O, 4, dtype=torch.float64)
    kB, kA = _unnorm_slope(B, align), _unnorm_slope(A, align)
    for o in reversed(range(O)):
        sx, sy, tx, ty = z[o]
        Gv = G * ((pre[o] >= 0) & (pre[o] <= 1))
        # paste = tent(ppy) x tent(ppx)
        xbI, ppx = _coords(B, B, 1 / sx, -tx / sx, align)
        ybI, ppy = _coords(A, A, 1 / sy, -ty / sy, align)
        tX, dX = _tent(ppx, B)
        tY, dY = _tent(ppy, A)
        gppx = (Gv * tY[:, None]).sum(0) * dX
        gppy = (Gv * tX[None, :]).sum(1) * dY
        gsx = (gppx * kB * (-(xbI - tx) / sx ** 2)).sum()
        gtx = (gppx * kB * (-1 / sx)).sum()
        gsy = (gppy * kA * (-(ybI - ty) / sy ** 2)).sum()
        gty = (gppy * kA * (-1 / sy)).sum()
        # patch + marg sampling
        xb, px = _coords(pb, B, sx, tx, align)
        yb, py = _coords(pa, A, sy, ty, align)
        inv = 1 - bgs[o]
        _, my, mx = _bilinear(inv, py, px)
        dpx = -g_marg[o] * mx
        dpy = -g_marg[o] * my
        for c in range(C):
            _, qy, qx = _bilinear(img[c], py, px)
            dpx = dpx + g_patch[o, c] * qx
            dpy = dpy + g_patch[o, c] * qy
        gsx = gsx + (dpx * kB * xb[None, :]).sum()
        gtx = gtx + (dpx * kB).sum()
        gsy = gsy + (dpy * kA * yb[:, None]).sum()
        gty = gty + (dpy * kA).sum()
        gz[o] = torch.stack([gsx, gsy, gtx, gty])
        # d marg / d bg_o : + bilinear weights (scatter)
        Gp = Gv.clone()
        y0, x0 = torch.floor(py).long(), torch.floor(px).long()
        fy, fx = py - y0, px - x0
        for i in range(pa):
            for j in range(pb):
                for dy in (0, 1):
                    for dx in (0, 1):
                        yy, xx = int(y0[i]) + dy, int(x0[j]) + dx
                        if 0 <= yy < A and 0 <= xx < B:
                            wgt = (fy[i] if dy else 1 - fy[i]) * (fx[j] if dx else 1 - fx[j])
                            Gp[yy, xx] += g_marg[o, i, j] * wgt
        G = Gp
    return gz


# ------------------------------------------------------------------------------------------------
# Identities the fused scene-likelihood kernels (csrc/scene_ll.cu, scene_ll_bwd.cu) rely on
# ------------------------------------------------------------------------------------------------
def leaf_poly_table(mu, a, b):
    """(mu, a, b) -> (c1, c2, c0) with  a (x - mu)^2 + b = c2 x^2 - c1 x + c0  (scene_ll.cuh: stage_leaf_poly)."""
    return 2 * a * mu, a, a * mu * mu + b


def leaf_poly_backward(x, w, gl, c1, c2, c0):
    """Input pass of the object SPN in polynomial form: leaf value L_g = -w u_g(x); given gl_g = dloss/dL_g returns
    (dloss/dx, dloss/dmask) with w = 1 - mask: three sums over the Gaussians, x enters at the end."""
    A1, A2, A3 = (gl * c2).sum(-1), (gl * c1).sum(-1), (gl * c0).sum(-1)
    return -w * (2 * x * A1 - A2), x * x * A1 - x * A2 + A3


def sums_from_saved(sum_val, leaf0, leaf1):
    """Linear-domain value of a sum node from what the forward pass saved: sum = m0 + m1 + log T (max-shifted form of
    rat_torch.py:202-222)  =>  T = exp(sum - m0 - m1), m = max of the leaf vector."""
    return torch.exp(sum_val - leaf0.max(-1, keepdim=True)[0] - leaf1.max(-1, keepdim=True)[0])


def sequence_mode_backward(g_elbo, n, T, skip, t, z4, out_obj, g_state, beta):
    """What the backward kernel does in sequence mode for one object of frame t (1 <= t < T): z4 = (sx, q = sy / sx, x, y)
    as stored in z_sup / z_s, out_obj = raw object log-likelihood, g_state = dloss / d(sx, sy, x, y) through glimpse and
    masks computed with the unit-free frame weight.  Returns (frame weight, d loss / d raw object ll, d loss / d overlap,
    d loss / d (sx, q, x, y)) -- stove.py:731-748, supair.py:79, 84-85, 151-158."""
    S, nsup = T - skip, skip - 1
    w = g_elbo / (n * S) if t >= skip else g_elbo / (n * nsup)
    sx, q = z4[0], z4[1]
    sy = sx * q
    gsx = g_state[0] + w * out_obj * sy
    gsy = g_state[1] + w * out_obj * sx
    return w, w * sx * sy, -beta * w, torch.stack([gsx + gsy * q, gsy * sx, g_state[2], g_state[3]])
