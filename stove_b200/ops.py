"""torch.autograd wrappers around the C ABI (include/stove_b200.h).

PyTorch is plumbing here: it owns device memory, streams and the autograd tape; all
arithmetic of the hot path happens in the CUDA kernels of stove_b200/csrc.
"""
import ctypes as C
import os

import torch

from . import _native as N


def _c(t):
    return None if t is None else t.contiguous()


# ----------------------------------------------------------------------------------------
# bw_transform (model/utils/utils.py:10-15)
# ----------------------------------------------------------------------------------------
class IndirectFrames:
    """Frames addressed through a device cell: `cell` (int64[1], CUDA) holds the address of a contiguous
    (n, T, C, W, H) batch of `dtype` (float32 or uint8).  A captured CUDA graph whose first kernel
    (bw_transform) reads its input through the cell is re-pointed at the next batch with `point_at(x)` -- an
    8-byte write -- instead of a copy of the batch into a static buffer (dp.GraphedStep)."""

    def __init__(self, example):
        self.shape, self.dtype, self.device = tuple(example.shape), example.dtype, example.device
        self.cell = torch.zeros(1, dtype=torch.int64, device=example.device)
        self._keep = None
        self.point_at(example)

    def point_at(self, x):
        if tuple(x.shape) != self.shape or x.dtype != self.dtype or not x.is_contiguous() or x.data_ptr() % 16:
            raise ValueError('IndirectFrames: batch must be contiguous, 16-byte aligned, %s %s' % (self.shape, self.dtype))
        self._keep = x                         # the batch must stay alive until the consumer has run
        self.cell.fill_(x.data_ptr())

    @property
    def is_cuda(self):
        return True


def bw_transform(x, want_planes=False):
    """(n, T, C, W, H) -> (n, T, 1, W, H): sum colour channels, clamp to [0, 1] (utils.py:10-15).
    x: fp32 in [0, 1], or uint8 frames (scaled by 1/255 inside the kernel), or an IndirectFrames.  With
    `want_planes` also returns the (hi, lo) TF32 operand planes (2, n*T, W*H) of the result for the recognition
    LSTM's input GEMM."""
    ind = isinstance(x, IndirectFrames)
    if not x.is_cuda:
        raise RuntimeError('stove_b200 kernels need CUDA tensors (got a %s tensor); there is no CPU path' % x.device)
    if x.dtype not in (torch.float32, torch.uint8):
        raise RuntimeError('stove_b200: frames must be float32 in [0, 1] or uint8 (got %s)' % x.dtype)
    if not ind:
        x = x.contiguous()
    n, T, ch, w, h = x.shape
    y = torch.empty(n, T, 1, w, h, device=x.device, dtype=torch.float32)
    pl = torch.empty(2, n * T, w * h, device=x.device, dtype=torch.float32) if want_planes else None
    N.check(N.lib().stove_bw_transform_ex(N.ptr(x.cell) if ind else N.ptr(x), int(x.dtype == torch.uint8), int(ind),
                                          N.ptr(y), N.ptr(pl), n * T, ch, w * h, N.stream()))
    return (y, pl) if want_planes else y


def render(bg, patches, z, width, height, align_corners=False):
    """Frames from states: clamp(bg + sum_o paste(patches[:, o], z[:, o]), 0, 1) (supair.py:425-501, the paste loop).
    bg (C, A, B) or (F, C, A, B); patches (O, C, pa, pb) or (F, O, C, pa, pb); z (F, O, 4) = (sx, sy, x, y)
    -> (F, C, A, B).  No autograd."""
    z = z.detach().contiguous()
    bg, patches = bg.detach().contiguous(), patches.detach().contiguous()
    N.require_cuda_f32(z, bg, patches)
    F, O = z.shape[0], z.shape[1]
    Cc, pa, pb = patches.shape[-3], patches.shape[-2], patches.shape[-1]
    out = torch.empty(F, Cc, width, height, device=z.device, dtype=z.dtype)
    N.check(N.lib().stove_render(F, O, Cc, width, height, pa, pb, int(bool(align_corners)), N.ptr(bg), int(bg.dim() == 4),
                                 N.ptr(patches), int(patches.dim() == 5), N.ptr(z), N.ptr(out), N.stream()))
    return out


# ----------------------------------------------------------------------------------------
# SPN parameter packing
# ----------------------------------------------------------------------------------------
class PackLeaf(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, sigma, dst_row, GP, prow_total, vmin, vmax):
        means, sigma = means.contiguous(), sigma.contiguous()
        N.require_cuda_f32(means, sigma)
        rows, G = means.shape
        # when every row of the table is some leaf's row (dst_row is a permutation: the SuPAIR structures) nothing needs
        # clearing -- the kernel writes the padding Gaussians; structures with unequal regions leave unused rows
        alloc = torch.empty if rows == prow_total else torch.zeros
        packed = alloc(prow_total, 3, GP, device=means.device, dtype=means.dtype)
        N.check(N.lib().stove_spn_pack_leaf_fwd(N.ptr(means), N.ptr(sigma), N.ptr(dst_row), rows, G, GP,
                                                vmin, vmax, N.ptr(packed), N.stream()))
        ctx.save_for_backward(sigma, dst_row)
        ctx.meta = (rows, G, GP, vmin, vmax)
        return packed

    @staticmethod
    def backward(ctx, g_packed):
        sigma, dst_row = ctx.saved_tensors
        rows, G, GP, vmin, vmax = ctx.meta
        g_packed = g_packed.contiguous()
        g_means = torch.empty(rows, G, device=sigma.device, dtype=sigma.dtype)
        g_sigma = torch.empty_like(g_means)
        N.check(N.lib().stove_spn_pack_leaf_bwd(N.ptr(sigma), N.ptr(dst_row), rows, G, GP, vmin, vmax,
                                                N.ptr(g_packed), N.ptr(g_means), N.ptr(g_sigma), N.stream()))
        return g_means, g_sigma, None, None, None, None, None


class PackSum(torch.autograd.Function):
    """raw [nb, K, S] -> (wlog, wlin) [nb, K, SP]; only wlog carries gradient (w.r.t. log-weights)."""

    @staticmethod
    def forward(ctx, raw, SP):
        raw = raw.contiguous()
        N.require_cuda_f32(raw)
        nb, K, S = raw.shape
        wlog = torch.empty(nb, K, SP, device=raw.device, dtype=raw.dtype)       # padding columns written by the kernel
        wlin = torch.empty_like(wlog)
        N.check(N.lib().stove_spn_pack_sum_fwd(N.ptr(raw), nb, K, S, SP, N.ptr(wlog), N.ptr(wlin), N.stream()))
        ctx.save_for_backward(wlog)
        ctx.meta = (nb, K, S, SP)
        ctx.mark_non_differentiable(wlin)
        return wlog, wlin

    @staticmethod
    def backward(ctx, g_wlog, _g_wlin):
        (wlog,) = ctx.saved_tensors
        nb, K, S, SP = ctx.meta
        g_wlog = g_wlog.contiguous()
        g_raw = torch.empty(nb, K, S, device=wlog.device, dtype=wlog.dtype)
        N.check(N.lib().stove_spn_pack_sum_bwd(N.ptr(wlog), nb, K, S, SP, N.ptr(g_wlog), N.ptr(g_raw), N.stream()))
        return g_raw, None


def interleave_leaf(leaf, row_map):
    """Lane-interleaved copy of a packed leaf table with 8-float parameter rows ([row][3][8] -> blocks of 32 rows laid
    out [6 float4 parts][32 rows]): a warp whose lanes own consecutive rows reads it with fully coalesced 16-byte loads.
    row_map (int32): source row of every destination row, -1 = zero row.  Not differentiable -- the fused
    scene-likelihood kernels read it, parameter gradients are taken w.r.t. the plain table."""
    N.require_cuda_f32(leaf)
    n = row_map.numel()
    out = torch.empty((n + 31) // 32 * 32 * 24, device=leaf.device, dtype=leaf.dtype)
    N.check(N.lib().stove_spn_interleave_leaf(N.ptr(leaf.detach()), N.ptr(row_map), n, N.ptr(out), N.stream()))
    return out


# ----------------------------------------------------------------------------------------
# fused SPNs
# ----------------------------------------------------------------------------------------
_FORK = True


def set_fork(enabled):
    """Side streams on / off, here and in the library (option "fork"): off = every kernel of a step runs on
    the caller's stream, which is what a per-kernel timing pass needs.  Returns the previous setting."""
    global _FORK
    prev, _FORK = _FORK, bool(enabled)
    N.set_option('fork', int(_FORK))
    return prev


def fork_enabled():
    return _FORK


_DYN_STREAM = True


def set_dyn_stream(enabled):
    """The dynamics weights are packed on a side stream of their own (True) or on the stream that packs the SPN
    parameters (False).  autograd runs a node's backward on the stream of its forward, and DynamicsLoop.backward
    joins its weight-gradient kernels into the stream that packed the weights: with one stream for both, the
    backward of the SPN packing (ready as soon as the SPN parameter-gradient kernels are) queues behind the
    weight-gradient kernels of the dynamics loop and ends up as the tail of the step.  Returns the previous
    setting."""
    global _DYN_STREAM
    prev, _DYN_STREAM = _DYN_STREAM, bool(enabled)
    return prev


def dyn_stream_enabled():
    return _DYN_STREAM


def _zeros_views(ref, *shapes):
    """Zero tensors of the given shapes carved out of ONE buffer (one fill launch instead of one each);
    every view starts on a 256-byte boundary."""
    numels = [int(torch.Size(sh).numel()) for sh in shapes]
    sizes = [(m + 63) // 64 * 64 for m in numels]
    buf = torch.zeros(sum(sizes), device=ref.device, dtype=ref.dtype)
    out, at = [], 0
    for sh, m, sz in zip(shapes, numels, sizes):
        out.append(buf[at:at + m].view(sh))
        at += sz
    return out


def _npad(n):
    return (n + 31) // 32 * 32


def _join_handle(stream):
    """cudaStream_t of the stream that receives the parameter-gradient kernels (None: the current one)."""
    if stream is None or not _FORK:
        return None
    return stream.cuda_stream


def _keep_for(stream, *tensors):
    """Kernels joined into `stream` still read / write these after the calling node returns."""
    if stream is None or not _FORK:
        return
    for t in tensors:
        if t is not None:
            t.record_stream(stream)


class Spn2(torch.autograd.Function):
    """Object SPN (D2 structure).  `tables` keeps the int32 structure tensors alive."""

    @staticmethod
    def forward(ctx, x, marg, leaf, wlog, wlin, rlog, rlin, tables, param_stream=None):
        """`param_stream`: the stream the packed parameters were produced on (their backward runs there):
        the parameter-gradient kernels of the backward pass are joined into it instead of the caller's
        stream, which then only carries what the rest of the chain waits for."""
        x, marg = _c(x), _c(marg)
        N.require_cuda_f32(x, marg, leaf, wlog, wlin, rlog, rlin)
        st = tables.cstruct
        ctx.param_stream = param_stream
        n = x.shape[0]
        npad = _npad(max(n, 1))
        Q, G, S = 2 * st.R, st.G, st.S
        leaf_val = torch.empty(Q * 2 * G, npad, device=x.device, dtype=x.dtype)
        sum_val = torch.empty(Q * S, npad, device=x.device, dtype=x.dtype)
        out = torch.empty(n, device=x.device, dtype=x.dtype)
        N.check(N.lib().stove_spn2_fwd(C.byref(st), n, N.ptr(x), N.ptr(marg), N.ptr(leaf), N.ptr(wlin),
                                       N.ptr(wlog), N.ptr(rlin), N.ptr(rlog), N.ptr(leaf_val),
                                       N.ptr(sum_val), N.ptr(out), N.stream()))
        ctx.save_for_backward(x, marg, leaf, wlog, wlin, rlog, rlin, leaf_val, sum_val, out)
        ctx.tables = tables
        ctx.has_marg = marg is not None
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, marg, leaf, wlog, wlin, rlog, rlin, leaf_val, sum_val, out = ctx.saved_tensors
        st = ctx.tables.cstruct
        n = x.shape[0]
        g_out = g_out.contiguous()
        need_x, need_m = ctx.needs_input_grad[0], ctx.has_marg and ctx.needs_input_grad[1]
        g_x = torch.empty_like(x) if need_x else None
        g_m = torch.empty_like(marg) if need_m else None
        g_leaf, g_wlog, g_rlog = _zeros_views(leaf, leaf.shape, wlog.shape, rlog.shape)
        ws = torch.empty(max(N.lib().stove_spn2_bwd_workspace(C.byref(st), n), 4) // 4, device=x.device,
                         dtype=torch.float32)
        N.check(N.lib().stove_spn2_bwd(C.byref(st), n, N.ptr(x), N.ptr(marg), N.ptr(leaf), N.ptr(wlin),
                                       N.ptr(wlog), N.ptr(rlin), N.ptr(rlog), N.ptr(leaf_val), N.ptr(sum_val),
                                       N.ptr(out), N.ptr(g_out), N.ptr(g_x), N.ptr(g_m), N.ptr(g_leaf),
                                       N.ptr(g_wlog), N.ptr(g_rlog), N.ptr(ws), N.stream(), _join_handle(ctx.param_stream)))
        _keep_for(ctx.param_stream, x, marg, leaf, wlin, rlin, ws, g_leaf, g_wlog, g_rlog)
        return g_x, g_m, g_leaf, g_wlog, None, g_rlog, None, None, None


class Spn1(torch.autograd.Function):
    """Background SPN (D1 structure)."""

    @staticmethod
    def forward(ctx, x, marg, leaf, rlog, rlin, tables, param_stream=None):
        x, marg = _c(x), _c(marg)
        N.require_cuda_f32(x, marg, leaf, rlog, rlin)
        st = tables.cstruct
        ctx.param_stream = param_stream
        n = x.shape[0]
        npad = _npad(max(n, 1))
        leaf_val = torch.empty(st.R * 2 * st.G, npad, device=x.device, dtype=x.dtype)
        out = torch.empty(n, device=x.device, dtype=x.dtype)
        ws = torch.empty(max(N.lib().stove_spn1_fwd_workspace(C.byref(st), n), 4) // 4, device=x.device,
                         dtype=torch.float32)
        N.check(N.lib().stove_spn1_fwd(C.byref(st), n, N.ptr(x), N.ptr(marg), N.ptr(leaf), N.ptr(rlin),
                                       N.ptr(rlog), N.ptr(leaf_val), N.ptr(out), N.ptr(ws), N.stream()))
        ctx.save_for_backward(x, marg, leaf, rlog, rlin, leaf_val, out)
        ctx.tables = tables
        ctx.has_marg = marg is not None
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, marg, leaf, rlog, rlin, leaf_val, out = ctx.saved_tensors
        st = ctx.tables.cstruct
        n = x.shape[0]
        g_out = g_out.contiguous()
        need_x, need_m = ctx.needs_input_grad[0], ctx.has_marg and ctx.needs_input_grad[1]
        g_x = torch.empty_like(x) if need_x else None
        g_m = torch.empty_like(marg) if need_m else None
        g_leaf, g_rlog = _zeros_views(leaf, leaf.shape, rlog.shape)
        ws = torch.empty(max(N.lib().stove_spn1_bwd_workspace(C.byref(st), n), 4) // 4, device=x.device,
                         dtype=torch.float32)
        N.check(N.lib().stove_spn1_bwd(C.byref(st), n, N.ptr(x), N.ptr(marg), N.ptr(leaf), N.ptr(rlin),
                                       N.ptr(rlog), N.ptr(leaf_val), N.ptr(out), N.ptr(g_out), N.ptr(g_x),
                                       N.ptr(g_m), N.ptr(g_leaf), N.ptr(g_rlog), N.ptr(ws), N.stream(),
                                       _join_handle(ctx.param_stream)))
        _keep_for(ctx.param_stream, x, marg, leaf, rlin, ws, g_leaf, g_rlog)
        return g_x, g_m, g_leaf, g_rlog, None, None, None


# ----------------------------------------------------------------------------------------
# glimpse + masks
# ----------------------------------------------------------------------------------------
class Scene(torch.autograd.Function):
    """img (F, C, A, B), z (F, O, 4) -> patches (F*O, C, pa, pb), marg_patch (same),
    marg_bg (F, C, A, B), overlap (F, O).  Differentiable w.r.t. z only."""

    @staticmethod
    def forward(ctx, img, z, pa, pb, align_corners):
        img, z = img.contiguous(), z.contiguous()
        N.require_cuda_f32(img, z)
        F_, Cc, A, B = img.shape
        O = z.shape[1]
        dev, dt = img.device, img.dtype
        patches = torch.empty(F_ * O, Cc, pa, pb, device=dev, dtype=dt)
        marg_patch = torch.empty_like(patches)
        marg_bg = torch.empty(F_, Cc, A, B, device=dev, dtype=dt)
        overlap = torch.empty(F_, O, device=dev, dtype=dt)
        N.check(N.lib().stove_scene_fwd(F_, O, Cc, A, B, pa, pb, int(align_corners), N.ptr(img), N.ptr(z),
                                        N.ptr(patches), N.ptr(marg_patch), N.ptr(marg_bg), N.ptr(overlap),
                                        N.stream()))
        ctx.save_for_backward(img, z)
        ctx.meta = (F_, O, Cc, A, B, pa, pb, int(align_corners))
        return patches, marg_patch, marg_bg, overlap

    @staticmethod
    def backward(ctx, g_patches, g_marg_patch, g_marg_bg, g_overlap):
        img, z = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise NotImplementedError('stove_b200: gradient w.r.t. the frames is not implemented '
                                      '(frames are data on the STOVE hot path)')
        F_, O, Cc, A, B, pa, pb, ac = ctx.meta
        g_z = torch.empty_like(z)
        N.check(N.lib().stove_scene_bwd(F_, O, Cc, A, B, pa, pb, ac, N.ptr(img), N.ptr(z),
                                        N.ptr(_c(g_patches)), N.ptr(_c(g_marg_patch)), N.ptr(_c(g_marg_bg)),
                                        N.ptr(_c(g_overlap)), N.ptr(g_z), N.stream()))
        return None, g_z, None, None, None


# ----------------------------------------------------------------------------------------
# fused scene likelihood: glimpses + masks + object SPN + background SPN (csrc/scene_ll.cu)
# ----------------------------------------------------------------------------------------
_SCENE_LL = True
_SCENE_LL_BWD = True          # False: the fused forward is differentiated by the unfused backward kernels (tests)


def set_scene_ll(enabled):
    """Fused scene-likelihood kernels on / off (off = Scene -> Spn2 / Spn1, the unfused launch sequence that the
    parity tests compare against).  Returns the previous setting."""
    global _SCENE_LL
    prev, _SCENE_LL = _SCENE_LL, bool(enabled)
    return prev


_SCENE_SEQ = True             # False: Stove.stove_forward goes ZAll -> SceneLL -> ElboAssemble instead of SceneElbo


def set_scene_seq(enabled):
    global _SCENE_SEQ
    prev, _SCENE_SEQ = _SCENE_SEQ, bool(enabled)
    return prev


def scene_seq_enabled():
    return _SCENE_SEQ and _SCENE_LL and _SCENE_LL_BWD


def set_scene_ll_bwd(enabled):
    global _SCENE_LL_BWD
    prev, _SCENE_LL_BWD = _SCENE_LL_BWD, bool(enabled)
    return prev


def scene_ll_supported(img, z, pa, pb, obj_tables, bg_tables):
    if not _SCENE_LL or obj_tables is None or bg_tables is None or not img.is_cuda:
        return False
    if obj_tables.kind != 'D2' or bg_tables.kind != 'D1':
        return False
    F_, Cc, A, B = img.shape
    return bool(N.lib().stove_scene_ll_supported(F_, z.shape[1], Cc, A, B, pa, pb, C.byref(obj_tables.cstruct),
                                                 C.byref(bg_tables.cstruct)))


class SceneLL(torch.autograd.Function):
    """img (F, 1, A, B), z (F, O, 4), packed object / background SPN parameters ->
    (bg_ll (F,), obj_ll (F*O,), overlap (F, O), patches, marg_patch (F*O, 1, pa, pb), marg_bg (F, 1, A, B)).
    One forward launch; the last three outputs are by-products (not differentiable)."""

    @staticmethod
    def forward(ctx, img, z, leaf, wlog, wlin, rlog, rlin, bleaf, brlog, brlin, obj_tables, bg_tables, pa, pb,
                align_corners, obj_stream=None, bg_stream=None, bleaf_il_f=None, bleaf_il_b=None):
        # no zero tensors for the gradients of the three by-product outputs (autograd would fill 11 MB per step)
        ctx.set_materialize_grads(False)
        img, z = img.contiguous(), z.contiguous()
        N.require_cuda_f32(img, z, leaf, wlog, wlin, rlog, rlin, bleaf, brlog, brlin)
        F_, Cc, A, B = img.shape
        O = z.shape[1]
        st2, st1 = obj_tables.cstruct, bg_tables.cstruct
        dev, dt = img.device, img.dtype
        n = F_ * O
        npad, npad_f = _npad(max(n, 1)), _npad(max(F_, 1))
        Q, G, S = 2 * st2.R, st2.G, st2.S
        patches = torch.empty(n, Cc, pa, pb, device=dev, dtype=dt)
        marg_patch = torch.empty_like(patches)
        marg_bg = torch.empty(F_, Cc, A, B, device=dev, dtype=dt)
        overlap = torch.empty(F_, O, device=dev, dtype=dt)
        leaf_val = torch.empty(Q * 2 * G, npad, device=dev, dtype=dt)
        sum_val = torch.empty(Q * S, npad, device=dev, dtype=dt)
        out_obj = torch.empty(n, device=dev, dtype=dt)
        bleaf_val = torch.empty(st1.R * 2 * st1.G, npad_f, device=dev, dtype=dt)
        out_bg = torch.empty(F_, device=dev, dtype=dt)
        N.check(N.lib().stove_scene_ll_fwd(
            F_, O, A, B, pa, pb, int(align_corners), N.ptr(img), N.ptr(z),
            C.byref(st2), N.ptr(leaf), N.ptr(wlin), N.ptr(wlog), N.ptr(rlin), N.ptr(rlog),
            C.byref(st1), N.ptr(bg_tables.dev['bg_scope']), N.ptr(bg_tables.dev['bg_cnt']),
            N.ptr(bleaf), N.ptr(brlin), N.ptr(brlog), N.ptr(bleaf_il_f), bg_tables.il_stride_f,
            N.ptr(patches), N.ptr(marg_patch), N.ptr(marg_bg), N.ptr(overlap),
            N.ptr(leaf_val), N.ptr(sum_val), N.ptr(out_obj), N.ptr(bleaf_val), N.ptr(out_bg), None, N.stream()))
        ctx.save_for_backward(img, z, leaf, wlog, wlin, rlog, rlin, bleaf, brlog, brlin, patches, marg_patch, marg_bg,
                              leaf_val, sum_val, out_obj, bleaf_val, out_bg)
        ctx.tables = (obj_tables, bg_tables)
        ctx.meta = (F_, O, Cc, A, B, pa, pb, int(align_corners))
        ctx.streams = (obj_stream, bg_stream)
        ctx.bleaf_il_b = bleaf_il_b
        ctx.mark_non_differentiable(patches, marg_patch, marg_bg)
        return out_bg, out_obj, overlap, patches, marg_patch, marg_bg

    @staticmethod
    def backward(ctx, g_bg, g_obj, g_overlap, _gp, _gmp, _gmb):
        (img, z, leaf, wlog, wlin, rlog, rlin, bleaf, brlog, brlin, patches, marg_patch, marg_bg, leaf_val, sum_val,
         out_obj, bleaf_val, out_bg) = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise NotImplementedError('stove_b200: gradient w.r.t. the frames is not implemented '
                                      '(frames are data on the STOVE hot path)')
        obj_tables, bg_tables = ctx.tables
        st2, st1 = obj_tables.cstruct, bg_tables.cstruct
        F_, O, Cc, A, B, pa, pb, ac = ctx.meta
        obj_stream, bg_stream = ctx.streams
        n = F_ * O
        dev = img.device
        g_bg = torch.zeros_like(out_bg) if g_bg is None else g_bg.contiguous()
        g_obj = torch.zeros_like(out_obj) if g_obj is None else g_obj.contiguous()
        g_overlap = None if g_overlap is None else g_overlap.contiguous()
        g_leaf, g_wlog, g_rlog, g_bleaf, g_brlog = _zeros_views(leaf, leaf.shape, wlog.shape, rlog.shape, bleaf.shape,
                                                               brlog.shape)
        g_z = torch.empty_like(z)
        x2, m2 = patches.view(n, -1), marg_patch.view(n, -1)
        xb, mb = img.view(F_, -1), marg_bg.view(F_, -1)
        ws2 = torch.empty(max(N.lib().stove_spn2_bwd_workspace(C.byref(st2), n), 4) // 4, device=dev, dtype=torch.float32)
        ws1 = torch.empty(max(N.lib().stove_spn1_bwd_workspace(C.byref(st1), F_), 4) // 4, device=dev, dtype=torch.float32)
        if _SCENE_LL_BWD:
            N.check(N.lib().stove_scene_ll_bwd(
                F_, O, A, B, pa, pb, ac, N.ptr(img), N.ptr(z),
                C.byref(st2), N.ptr(leaf), N.ptr(wlin), N.ptr(wlog), N.ptr(rlin), N.ptr(rlog),
                C.byref(st1), N.ptr(bg_tables.dev['bg_scope']), N.ptr(bg_tables.dev['bg_cnt']),
                N.ptr(bleaf), N.ptr(brlin), N.ptr(brlog), N.ptr(ctx.bleaf_il_b), bg_tables.il_stride_b,
                N.ptr(x2), N.ptr(m2), N.ptr(mb), N.ptr(leaf_val), N.ptr(sum_val), N.ptr(out_obj), N.ptr(bleaf_val),
                N.ptr(out_bg), N.ptr(g_obj), N.ptr(g_bg), N.ptr(g_overlap),
                N.ptr(g_z), N.ptr(g_leaf), N.ptr(g_wlog), N.ptr(g_rlog), N.ptr(g_bleaf), N.ptr(g_brlog),
                N.ptr(ws2), N.ptr(ws1), None, N.stream(), _join_handle(obj_stream), _join_handle(bg_stream)))
            _keep_for(obj_stream, x2, m2, leaf, wlin, rlin, ws2, g_leaf, g_wlog, g_rlog)
            _keep_for(bg_stream, xb, mb, bleaf, brlin, ws1, g_bleaf, g_brlog)
            return (None, g_z, g_leaf, g_wlog, None, g_rlog, None, g_bleaf, g_brlog, None, None, None, None, None, None,
                    None, None, None, None)
        g_x, g_m = torch.empty_like(x2), torch.empty_like(m2)
        g_mb = torch.empty_like(mb)
        N.check(N.lib().stove_spn2_bwd(C.byref(st2), n, N.ptr(x2), N.ptr(m2), N.ptr(leaf), N.ptr(wlin), N.ptr(wlog),
                                       N.ptr(rlin), N.ptr(rlog), N.ptr(leaf_val), N.ptr(sum_val), N.ptr(out_obj),
                                       N.ptr(g_obj), N.ptr(g_x), N.ptr(g_m), N.ptr(g_leaf), N.ptr(g_wlog),
                                       N.ptr(g_rlog), N.ptr(ws2), N.stream(), _join_handle(obj_stream)))
        N.check(N.lib().stove_spn1_bwd(C.byref(st1), F_, N.ptr(xb), N.ptr(mb), N.ptr(bleaf), N.ptr(brlin),
                                       N.ptr(brlog), N.ptr(bleaf_val), N.ptr(out_bg), N.ptr(g_bg), None,
                                       N.ptr(g_mb), N.ptr(g_bleaf), N.ptr(g_brlog), N.ptr(ws1), N.stream(),
                                       _join_handle(bg_stream)))
        N.check(N.lib().stove_scene_bwd(F_, O, Cc, A, B, pa, pb, ac, N.ptr(img), N.ptr(z), N.ptr(g_x), N.ptr(g_m),
                                        N.ptr(g_mb), N.ptr(g_overlap), N.ptr(g_z), N.stream()))
        _keep_for(obj_stream, x2, m2, leaf, wlin, rlin, ws2, g_leaf, g_wlog, g_rlog)
        _keep_for(bg_stream, xb, mb, bleaf, brlin, ws1, g_bleaf, g_brlog)
        return (None, g_z, g_leaf, g_wlog, None, g_rlog, None, g_bleaf, g_brlog, None, None, None, None, None, None,
                None, None, None, None)


class SceneElbo(torch.autograd.Function):
    """The sequence ELBO's likelihood half (sequence mode of csrc/scene_ll*.cu + stove_elbo_fwd): the fused kernels read
    the states from z_sup / z_s and weight the object terms themselves; the backward kernel derives the per-frame
    weights from d loss / d elbo and writes the gradients of z_sup, z_s, log q and the transition likelihood -- what
    ZAll -> SceneLL -> ElboAssemble do with three launches forward and five on the backward chain (+ 2 fills, 1 add).
    img (n*(T-1), 1, A, B) = x[:, 1:]; z_sup (n, T, O, 4), z_s (n, S, O, Z) as [sx, sy/sx, x, y, ...]; logq, trans (n, S).
    Returns (elbo (), stats (8,), bg (F,), obj (F*O,) raw, overlap (F, O), patches, marg_patch, marg_bg); everything
    but the ELBO is a by-product (not differentiable)."""

    @staticmethod
    def forward(ctx, img, z_sup, z_s, logq, trans, skip, beta, leaf, wlog, wlin, rlog, rlin, bleaf, brlog, brlin,
                obj_tables, bg_tables, pa, pb, align_corners, obj_stream, bg_stream, bleaf_il_f, bleaf_il_b):
        ctx.set_materialize_grads(False)
        img, z_sup, z_s, logq, trans = (t.contiguous() for t in (img, z_sup, z_s, logq, trans))
        N.require_cuda_f32(img, z_sup, z_s, logq, trans, leaf, wlog, wlin, rlog, rlin, bleaf, brlog, brlin)
        F_, Cc, A, B = img.shape
        n, T, O, _ = z_sup.shape
        Z = z_s.shape[-1]
        if F_ != n * (T - 1) or z_s.shape[1] != T - skip:
            raise ValueError('SceneElbo: frames must be x[:, 1:] of the sequences of z_sup / z_s')
        st2, st1 = obj_tables.cstruct, bg_tables.cstruct
        dev, dt = img.device, img.dtype
        npt = F_ * O
        npad, npad_f = _npad(max(npt, 1)), _npad(max(F_, 1))
        Q, G, S = 2 * st2.R, st2.G, st2.S
        patches = torch.empty(npt, Cc, pa, pb, device=dev, dtype=dt)
        marg_patch = torch.empty_like(patches)
        marg_bg = torch.empty(F_, Cc, A, B, device=dev, dtype=dt)
        overlap = torch.empty(F_, O, device=dev, dtype=dt)
        leaf_val = torch.empty(Q * 2 * G, npad, device=dev, dtype=dt)
        sum_val = torch.empty(Q * S, npad, device=dev, dtype=dt)
        out_obj = torch.empty(npt, device=dev, dtype=dt)
        bleaf_val = torch.empty(st1.R * 2 * st1.G, npad_f, device=dev, dtype=dt)
        out_bg = torch.empty(F_, device=dev, dtype=dt)
        patch_w = torch.empty(npt, device=dev, dtype=dt)
        stats = torch.empty(8, device=dev, dtype=dt)
        elbo = torch.empty((), device=dev, dtype=dt)
        seq = N.SceneSeq(n, T, skip, Z, float(beta), N.ptr(z_sup), N.ptr(z_s), N.ptr(patch_w), None, None, None, None, None)
        N.check(N.lib().stove_scene_ll_fwd(
            F_, O, A, B, pa, pb, int(align_corners), N.ptr(img), None,
            C.byref(st2), N.ptr(leaf), N.ptr(wlin), N.ptr(wlog), N.ptr(rlin), N.ptr(rlog),
            C.byref(st1), N.ptr(bg_tables.dev['bg_scope']), N.ptr(bg_tables.dev['bg_cnt']),
            N.ptr(bleaf), N.ptr(brlin), N.ptr(brlog), N.ptr(bleaf_il_f), bg_tables.il_stride_f,
            N.ptr(patches), N.ptr(marg_patch), N.ptr(marg_bg), N.ptr(overlap),
            N.ptr(leaf_val), N.ptr(sum_val), N.ptr(out_obj), N.ptr(bleaf_val), N.ptr(out_bg), C.byref(seq), N.stream()))
        N.check(N.lib().stove_elbo_fwd(n, T, skip, O, float(beta), N.ptr(out_bg), N.ptr(patch_w), None, N.ptr(overlap),
                                       N.ptr(logq), N.ptr(trans), N.ptr(stats), N.ptr(elbo), N.stream()))
        ctx.save_for_backward(img, z_sup, z_s, leaf, wlog, wlin, rlog, rlin, bleaf, brlog, brlin, patches, marg_patch,
                              marg_bg, leaf_val, sum_val, out_obj, bleaf_val, out_bg)
        ctx.tables = (obj_tables, bg_tables)
        ctx.meta = (F_, O, Cc, A, B, pa, pb, int(align_corners), n, T, skip, Z, float(beta), tuple(logq.shape))
        ctx.streams = (obj_stream, bg_stream)
        ctx.bleaf_il_b = bleaf_il_b
        ctx.mark_non_differentiable(stats, out_bg, out_obj, overlap, patches, marg_patch, marg_bg)
        return elbo, stats, out_bg, out_obj, overlap, patches, marg_patch, marg_bg

    @staticmethod
    def backward(ctx, g_elbo, *_unused):
        (img, z_sup, z_s, leaf, wlog, wlin, rlog, rlin, bleaf, brlog, brlin, patches, marg_patch, marg_bg, leaf_val,
         sum_val, out_obj, bleaf_val, out_bg) = ctx.saved_tensors
        if g_elbo is None:
            return (None,) * 24
        if ctx.needs_input_grad[0]:
            raise NotImplementedError('stove_b200: gradient w.r.t. the frames is not implemented '
                                      '(frames are data on the STOVE hot path)')
        obj_tables, bg_tables = ctx.tables
        st2, st1 = obj_tables.cstruct, bg_tables.cstruct
        F_, O, Cc, A, B, pa, pb, ac, n, T, skip, Z, beta, lq_shape = ctx.meta
        obj_stream, bg_stream = ctx.streams
        npt = F_ * O
        dev = img.device
        g_elbo = g_elbo.contiguous()
        g_leaf, g_wlog, g_rlog, g_bleaf, g_brlog = _zeros_views(leaf, leaf.shape, wlog.shape, rlog.shape, bleaf.shape,
                                                               brlog.shape)
        g_z_sup, g_z_s = torch.empty_like(z_sup), torch.empty_like(z_s)
        g_lq = torch.empty(lq_shape, device=dev, dtype=img.dtype)
        g_tr = torch.empty(lq_shape, device=dev, dtype=img.dtype)
        x2, m2 = patches.view(npt, -1), marg_patch.view(npt, -1)
        xb, mb = img.view(F_, -1), marg_bg.view(F_, -1)
        ws2 = torch.empty(max(N.lib().stove_spn2_bwd_workspace(C.byref(st2), npt), 4) // 4, device=dev, dtype=torch.float32)
        ws1 = torch.empty(max(N.lib().stove_spn1_bwd_workspace(C.byref(st1), F_), 4) // 4, device=dev, dtype=torch.float32)
        seq = N.SceneSeq(n, T, skip, Z, beta, N.ptr(z_sup), N.ptr(z_s), None, N.ptr(g_elbo), N.ptr(g_z_sup), N.ptr(g_z_s),
                         N.ptr(g_lq), N.ptr(g_tr))
        N.check(N.lib().stove_scene_ll_bwd(
            F_, O, A, B, pa, pb, ac, N.ptr(img), None,
            C.byref(st2), N.ptr(leaf), N.ptr(wlin), N.ptr(wlog), N.ptr(rlin), N.ptr(rlog),
            C.byref(st1), N.ptr(bg_tables.dev['bg_scope']), N.ptr(bg_tables.dev['bg_cnt']),
            N.ptr(bleaf), N.ptr(brlin), N.ptr(brlog), N.ptr(ctx.bleaf_il_b), bg_tables.il_stride_b,
            N.ptr(x2), N.ptr(m2), N.ptr(mb), N.ptr(leaf_val), N.ptr(sum_val), N.ptr(out_obj), N.ptr(bleaf_val),
            N.ptr(out_bg), None, None, None,
            None, N.ptr(g_leaf), N.ptr(g_wlog), N.ptr(g_rlog), N.ptr(g_bleaf), N.ptr(g_brlog),
            N.ptr(ws2), N.ptr(ws1), C.byref(seq), N.stream(), _join_handle(obj_stream), _join_handle(bg_stream)))
        _keep_for(obj_stream, x2, m2, leaf, wlin, rlin, ws2, g_leaf, g_wlog, g_rlog)
        _keep_for(bg_stream, xb, mb, bleaf, brlin, ws1, g_bleaf, g_brlog)
        return (None, g_z_sup, g_z_s, g_lq, g_tr, None, None, g_leaf, g_wlog, None, g_rlog, None, g_bleaf, g_brlog, None,
                None, None, None, None, None, None, None, None, None)


# ----------------------------------------------------------------------------------------
# fused sequence glue (constrain_zp + matching + fix_supair + velocities)
# ----------------------------------------------------------------------------------------
MATCH_KINDS = {'3_only': 0, 'greedy': 1, 'volatile': 2}


class SupPrepare(torch.autograd.Function):
    """zp (n, T, O, 8) [, app (n, T, O, 3)] -> z_sup (n,T,O,4), z_full (n,T,O,6), std_full (n,T,O,6),
    app_matched (n,T,O,3) | None."""

    @staticmethod
    def forward(ctx, zp, app, cfg):
        ctx.set_materialize_grads(False)
        zp, app = zp.contiguous(), _c(app)
        N.require_cuda_f32(zp, app)
        n, T, O, _ = zp.shape
        dev, dt = zp.device, zp.dtype
        z_sup = torch.empty(n, T, O, 4, device=dev, dtype=dt)
        z_full = torch.empty(n, T, O, 6, device=dev, dtype=dt)
        std_full = torch.empty(n, T, O, 6, device=dev, dtype=dt)
        app_out = torch.empty(n, T, O, 3, device=dev, dtype=dt) if app is not None else None
        idx = torch.empty(n, T, O, device=dev, dtype=torch.int32)
        flag = torch.empty(n, T, O, device=dev, dtype=torch.int32)
        N.check(N.lib().stove_sup_prepare_fwd(C.byref(cfg), n, N.ptr(zp), N.ptr(app), N.ptr(z_sup), N.ptr(z_full),
                                              N.ptr(std_full), N.ptr(app_out), N.ptr(idx), N.ptr(flag), N.stream()))
        ctx.save_for_backward(zp, idx, flag, std_full)
        ctx.cfg = cfg
        if app_out is None:
            app_out = zp.new_zeros(())
        ctx.mark_non_differentiable(app_out)
        return z_sup, z_full, std_full, app_out

    @staticmethod
    def backward(ctx, g_z_sup, g_z_full, g_std_full, _g_app):
        zp, idx, flag, std_full = ctx.saved_tensors
        g_zp = torch.empty_like(zp)
        N.check(N.lib().stove_sup_prepare_bwd(C.byref(ctx.cfg), zp.shape[0], N.ptr(zp), N.ptr(idx), N.ptr(flag),
                                              N.ptr(std_full), N.ptr(_c(g_z_sup)), N.ptr(_c(g_z_full)),
                                              N.ptr(_c(g_std_full)), N.ptr(g_zp), N.stream()))
        return g_zp, None, None


# ----------------------------------------------------------------------------------------
# 3xTF32 operands and GEMM of the recognition LSTM (csrc/lstm_tc.cu)
# ----------------------------------------------------------------------------------------
def _round4(v):
    return (v + 3) // 4 * 4


def split_planes(x, planes=True, transposed=False):
    """x (rows, cols) -> (pl (2, rows, cols) | None, plT (2, cols, ldT) | None): the TF32-exact part `hi` of
    every element and the remainder `lo = x - hi` as two planes; plT holds the transposed planes with
    ldT = rows rounded up to 4 (pad columns zero).  No autograd."""
    N.require_cuda_f32(x)
    if x.dim() != 2 or x.stride(1) != 1:
        x = x.contiguous()
    rows, cols = x.shape
    pl = torch.empty(2, rows, cols, device=x.device, dtype=x.dtype) if planes else None
    ldT = _round4(rows)
    plT = torch.empty(2, cols, ldT, device=x.device, dtype=x.dtype) if transposed else None
    N.check(N.lib().stove_split_planes(rows, cols, N.ptr(x), x.stride(0), N.ptr(pl), N.ptr(plT), ldT, N.stream()))
    return pl, plT


def tc3_gemm(a_pl, b_pl, K=None, parts=1, out=None, bn=0):
    """a_pl (2, M, lda), b_pl (2, N, ldb): (hi, lo) planes -> a @ b.t() over the first K columns as
    `parts` split-K partial products (parts, M, N) (parts may come back smaller), 3xTF32 on tcgen05.
    bn: tile width (0 = automatic, 256 = fewest operand bytes)."""
    M, lda = a_pl.shape[1], a_pl.shape[2]
    Nn, ldb = b_pl.shape[1], b_pl.shape[2]
    K = min(lda, ldb) if K is None else K
    parts = N.lib().stove_tc3_gemm_parts(M, Nn, K, parts)
    if out is None:
        out = torch.empty(parts, M, Nn, device=a_pl.device, dtype=a_pl.dtype)
    N.check(N.lib().stove_tc3_gemm(M, Nn, K, N.ptr(a_pl), lda, a_pl.stride(0), N.ptr(b_pl), ldb, b_pl.stride(0),
                                   N.ptr(out), Nn, parts, M * Nn, bn, N.stream()))
    return out


def sum_parts(parts_t, out=None, scale=1.0):
    """(P, ...) -> scale * sum over the first dimension in a fixed order (split-K partials, bias partial sums)."""
    P = parts_t.shape[0]
    numel = parts_t[0].numel()
    if P == 1 and out is None and scale == 1.0:
        return parts_t[0]
    if out is None:
        out = torch.empty(parts_t.shape[1:], device=parts_t.device, dtype=parts_t.dtype)
    N.check(N.lib().stove_sum_parts(numel, P, numel, N.ptr(parts_t), N.ptr(out), float(scale), N.stream()))
    return out


def _split_k(tiles, num_kb, sms=148, cap=8):
    """split-K factor that fills the machine: about one CTA per SM, at least 4 k-blocks per part, at most `cap`
    parts (every part is another partial product to write and sum)"""
    return max(1, min(sms // max(tiles, 1), num_kb // 4, cap))


_AUX_STREAMS = {}


def _aux_stream(device, which=0):
    """Library-wide side streams (per device) for work that is off the critical chain of a backward pass."""
    if not _FORK:                                    # serial execution (per-kernel timing passes)
        return torch.cuda.current_stream(device)
    key = (device.type, device.index, which)
    if key not in _AUX_STREAMS:
        _AUX_STREAMS[key] = torch.cuda.Stream(device=device)
    return _AUX_STREAMS[key]


class LstmEncoder(torch.autograd.Function):
    """h_1 .. h_steps of a one-layer LSTM that is fed the SAME input x at every step from a zero state
    (the recognition network, encoder.py:50-51; nn.LSTM gate order and parameter shapes).
    x (n, K), w_ih (4H, K), w_hh (4H, H), b_ih / b_hh (4H,) -> (n, steps, H).

    One autograd node, every contraction a hand-written tcgen05 kernel (csrc/lstm_tc.cu), 3xTF32 over (hi, lo)
    operand planes that are written once and read once per tile: the input GEMM is done once (not per
    step) with the LSTM cell as its epilogue; the backward pass runs cell kernel -> hidden-state GEMM per
    step, then the two weight-gradient GEMMs (contractions over the frames, fed by the transposed planes the
    forward / cell kernels emit) as split-K launches with a fixed-order reduction."""

    # data-parallel engines set this: an object with __call__(name, tensor, row_lo=None, row_hi=None) and flush();
    # names: w_ih (possibly in row blocks), w_hh, b_ih, b_hh, w1, b1, w2, b2 (stove_b200.dp.DataParallel)
    grad_sink = None

    @staticmethod
    def prepare(w_ih, w_hh, b_ih, b_hh):
        """Operand planes of the weights and the summed bias, on the library's side stream.  They depend on the
        parameters only: issued at the very start of a step (before the frames are even transformed) they are
        off the chain; pass the result as `prepared`."""
        dev = w_ih.device
        cur, side = torch.cuda.current_stream(dev), _aux_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            # what step 0 needs comes first and has its own event: the chain starts ~5 us earlier
            wih_pl, _ = split_planes(w_ih.detach())
            bias = (b_ih.detach() + b_hh.detach()).contiguous()
            ready0 = torch.cuda.Event()
            ready0.record(side)
            whh_pl, whhT_pl = split_planes(w_hh.detach(), True, True)
            done = torch.cuda.Event()
            done.record(side)
        return wih_pl, whh_pl, whhT_pl, bias, (ready0, done)

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, steps, w1=None, b1=None, w2=None, b2=None, prepared=None, x_pl=None):
        """With the head parameters (fc1 / fc2 of encoder.py:53-56) the node returns fc2(sigmoid(fc1(h_t)))
        (n, steps, P) instead of h_t: one autograd node for the whole recognition network, whose backward
        runs every parameter-gradient kernel (head, W_hh, biases) beside the chain h_t -> h_{t-1}."""
        N.require_cuda_f32(x, w_ih, w_hh, b_ih, b_hh)
        n, H = x.shape[0], w_hh.shape[1]
        dev, dt = x.device, x.dtype
        lib, st = N.lib(), N.stream()
        cur = torch.cuda.current_stream(dev)
        if prepared is None:
            prepared = LstmEncoder.prepare(w_ih, w_hh, b_ih, b_hh)
        wih_pl, whh_pl, whhT_pl, bias, (ready0, done) = prepared
        x = x.contiguous()
        if x_pl is None:                                   # (bw_transform can emit the planes in its own pass)
            x_pl, _ = split_planes(x)
        # the transposed planes (right operand of g^T x) are only needed by the backward pass, which builds them
        # off the chain
        cur.wait_event(ready0)                             # W_ih planes and the bias; W_hh's are awaited before step 1
        for t_ in (wih_pl, whh_pl, whhT_pl, bias):
            t_.record_stream(cur)
        out = torch.empty(n, steps, H, device=dev, dtype=dt)
        gx = torch.empty(n, 4 * H, device=dev, dtype=dt) if steps > 1 else None
        acts, cs = [], []
        # transposed (hi, lo) planes of h_0 .. h_{steps-2}, stacked along the columns: the backward contracts them
        # with the equally stacked gate gradients in ONE GEMM
        ldT = _round4(n)
        ldS = max(steps - 1, 1) * ldT
        hT_all = torch.empty(2, H, ldS, device=dev, dtype=dt) if steps > 1 else None
        h_pl = c_prev = None
        for t in range(steps):
            # one tcgen05 kernel per step: gate GEMM with the cell as its epilogue; step 0 contracts the frame
            # with W_ih and leaves gx = x W_ih^T + b for the later steps, which contract h_{t-1} with W_hh and add gx
            c = torch.empty(n, H, device=dev, dtype=dt)
            act = torch.empty(n, 4 * H, device=dev, dtype=dt)
            more = t + 1 < steps
            h_pl_next = torch.empty(2, n, H, device=dev, dtype=dt) if more else None
            hT = hT_all.data_ptr() + 4 * t * ldT if more else None
            if t == 0:
                N.check(lib.stove_lstm_gemm_cell_fwd(n, H, x_pl.shape[2], N.ptr(x_pl), N.ptr(wih_pl), N.ptr(bias), 1,
                                                     None, N.ptr(gx), out[:, t].data_ptr(), steps * H, N.ptr(c),
                                                     N.ptr(act), N.ptr(h_pl_next), hT, ldS, ldT, H * ldS, st))
            else:
                if t == 1:
                    cur.wait_event(done)
                N.check(lib.stove_lstm_gemm_cell_fwd(n, H, H, N.ptr(h_pl), N.ptr(whh_pl), N.ptr(gx), 0,
                                                     N.ptr(c_prev), None, out[:, t].data_ptr(), steps * H, N.ptr(c),
                                                     N.ptr(act), N.ptr(h_pl_next), hT, ldS, ldT, H * ldS, st))
            acts.append(act)
            cs.append(c)
            h_pl, c_prev = h_pl_next, c
        ctx.stash = (x, whhT_pl, acts, cs, hT_all, steps, H)
        ctx.head = None
        if w1 is not None:
            w1, b1, w2, b2 = w1.contiguous(), b1.contiguous(), w2.contiguous(), b2.contiguous()
            N.require_cuda_f32(w1, b1, w2, b2)
            hs = out.view(n * steps, H)
            hidden, zp = _head_fwd(hs, w1, b1, w2, b2)
            ctx.head = (hs, w1, w2, hidden)
            return zp.view(n, steps, w2.shape[0])
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, whhT_pl, acts, cs, hT_all, steps, H = ctx.stash
        g_out = g_out.contiguous()
        n = g_out.shape[0]
        dev, dt = g_out.device, g_out.dtype
        lib, st = N.lib(), N.stream()
        cur, side = torch.cuda.current_stream(dev), _aux_stream(dev)
        xT_pl = None
        if ctx.needs_input_grad[1]:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                _, xT_pl = split_planes(x, False, True)    # consumed by the W_ih gradient GEMM at the very end
                xT_ready = torch.cuda.Event()
                xT_ready.record(side)
        g_head = (None, None, None, None)
        if ctx.head is not None:
            hs, w1, w2, hidden = ctx.head
            g_zp = g_out.view(n * steps, w2.shape[0])
            g_hs, ws = _head_bwd_data(hs, w1, w2, hidden, g_zp)       # on the chain: the LSTM backward needs it
            # off the chain, on a stream of its own (the W_hh gradient GEMM must not queue behind it), joined at the end
            side2 = _aux_stream(dev, 1)
            side2.wait_stream(cur)
            with torch.cuda.stream(side2):
                g_head = _head_bwd_params(hs, w1, w2, hidden, g_zp, ws)
            for t_ in (ws, g_zp, hs, hidden) + g_head:
                t_.record_stream(side2)
            g_out = g_hs.view(n, steps, H)
        if ctx.needs_input_grad[0]:
            raise NotImplementedError('stove_b200: gradient w.r.t. the frames is not implemented '
                                      '(frames are data on the STOVE hot path)')
        H4 = 4 * H
        ldT = _round4(n)
        ldS = max(steps - 1, 1) * ldT
        g_acc = torch.empty(n, H4, device=dev, dtype=dt) if steps > 1 else None   # gate gradients summed over steps
        gT_all = torch.empty(2, H4, ldS, device=dev, dtype=dt) if steps > 1 else None   # steps 1 .. steps-1, stacked
        gsumT = torch.empty(2, H4, ldT, device=dev, dtype=dt)        # transposed planes of the summed gradient
        bias_part = torch.empty((ldT + 31) // 32, H4, device=dev, dtype=dt)
        dh = g_c = None
        for t in reversed(range(steps)):
            g_c_prev = torch.empty(n, H, device=dev, dtype=dt) if t > 0 else None
            g_pl = torch.empty(2, n, H4, device=dev, dtype=dt) if t > 0 else None
            if t > 0:
                gT, ld, plane = gT_all.data_ptr() + 4 * (t - 1) * ldT, ldS, H4 * ldS
            else:
                gT, ld, plane = gsumT.data_ptr(), ldT, H4 * ldT
            # the cell kernel writes this step's gate gradient as (hi, lo) planes, row-major for the hidden-state
            # GEMM below and transposed for the W_hh gradient; the last step processed (t = 0) emits the sum over
            # the steps instead (what W_ih sees) and the per-block column sums for the bias gradient
            N.check(lib.stove_lstm_cell_bwd_t(n, H, N.ptr(acts[t]), N.ptr(cs[t - 1]) if t > 0 else None,
                                              N.ptr(cs[t]), g_out[:, t].data_ptr(), steps * H, N.ptr(dh),
                                              dh.shape[0] if dh is not None else 0, N.ptr(g_c), N.ptr(g_pl), gT, ld,
                                              ldT, plane, N.ptr(g_acc), 0 if t == steps - 1 else 1,
                                              1 if t == 0 else 0, N.ptr(bias_part) if t == 0 else None,
                                              N.ptr(g_c_prev), st))
            if t == 1:
                # every gate gradient that W_hh sees exists now (steps 1 .. steps-1): g_W_hh = sum_t g_t^T h_{t-1}
                # is ONE GEMM over the stacked transposed planes, on the side stream, under the last hidden-state
                # GEMM and the last cell kernel -- not at the tail of the node, where it was the longest branch
                # (profiles/r02_timeline_v1_lstm_planes.txt: 40 us + 17 us of reductions behind the W_ih GEMM)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    # (128 x 256 tiles -- a quarter less operand traffic, two pipeline stages instead of three --
                    # measured slower here: 40 vs 29 us, profiles/r02_timeline_v3_bn256.txt)
                    tiles = (H4 // 128) * ((H + 127) // 128)
                    g_whh = sum_parts(tc3_gemm(gT_all, hT_all, parts=_split_k(tiles, (ldS + 31) // 32, cap=4)))
                for t_ in (gT_all, hT_all):
                    t_.record_stream(side)
            if t > 0:
                # step t read h_{t-1}: the gradient flowing back into h_{t-1} stays on the chain; split-K fills the
                # machine and the next cell kernel sums the parts
                dh = tc3_gemm(g_pl, whhT_pl, parts=_split_k(((n + 127) // 128) * ((H + 127) // 128), H4 // 32))
                g_c = g_c_prev
        if steps == 1:
            g_whh = torch.zeros(H4, H, device=dev, dtype=dt)
        sink = LstmEncoder.grad_sink
        K = x.shape[1]
        side.wait_stream(cur)
        with torch.cuda.stream(side):              # bias gradient: fixed-order sum of the per-block column sums
            g_b = sum_parts(bias_part)
        bias_part.record_stream(side)
        g_wih = None
        if sink is None:
            if ctx.needs_input_grad[1]:
                cur.wait_event(xT_ready)
                xT_pl.record_stream(cur)
                tiles = (H4 // 128) * ((K + 127) // 128)
                g_wih = sum_parts(tc3_gemm(gsumT, xT_pl, parts=_split_k(tiles, (ldT + 31) // 32)))
            cur.wait_stream(side)                  # W_hh and bias gradients
        else:
            cur.wait_stream(side)                  # W_hh and bias gradients
            if ctx.head is not None:
                cur.wait_stream(_aux_stream(dev, 1))   # head gradients
            # data-parallel run: gradients are handed to the engine's sink the moment they exist, so their
            # all-reduce overlaps what is still being computed: W_hh / biases / head are gathered into the bucket
            # now, W_ih follows in two row blocks -- the first block (with everything before it) is on the wire
            # while the second is computed; only the last 2 MB are exposed.
            sink('w_hh', g_whh)
            sink('b_ih', g_b)
            sink('b_hh', g_b)
            for name, t_ in zip(('w1', 'b1', 'w2', 'b2'), g_head):
                if t_ is not None:
                    sink(name, t_)
            sink.flush()                           # gathered into the bucket; nothing on the wire yet
            if ctx.needs_input_grad[1]:
                cur.wait_event(xT_ready)
                xT_pl.record_stream(cur)
                half = (H4 // 2 + 127) // 128 * 128
                for lo, hi in ((half, H4), (0, half)):          # upper rows first: they touch the rest of the bucket
                    if lo >= hi:
                        continue
                    tiles = ((hi - lo + 127) // 128) * ((K + 127) // 128)
                    parts = tc3_gemm(gsumT[:, lo:hi], xT_pl, parts=_split_k(tiles, (ldT + 31) // 32, cap=4))
                    # the last reduction of the split-K parts writes these rows straight into the engine's bucket,
                    # already scaled by 1 / world size: no gather launch between the GEMM and the wire.  (The node then
                    # returns no W_ih gradient of its own; the engine hands out views of the bucket.)
                    dst = sink.slot('w_ih', lo, hi)
                    if dst is None:
                        if g_wih is None:
                            g_wih = torch.empty(H4, K, device=dev, dtype=dt)
                        sum_parts(parts, out=g_wih[lo:hi])
                        sink('w_ih', g_wih, lo, hi)
                    else:
                        sum_parts(parts, out=dst, scale=sink.scale)
                        sink.delivered('w_ih', lo, hi)
                    sink.flush(send=True)          # all-reduce of everything in the bucket so far
        if ctx.head is not None:
            cur.wait_stream(_aux_stream(dev, 1))   # head gradients
        for t_ in (g_whh, g_b) + g_head:
            if t_ is not None:
                t_.record_stream(cur)
        return (None, g_wih, g_whh, g_b, g_b, None) + g_head + (None, None)


def _head_fwd(x2, w1, b1, w2, b2):
    R, K = x2.shape
    J, P = w1.shape[0], w2.shape[0]
    hidden = torch.empty(R, J, device=x2.device, dtype=x2.dtype)
    out = torch.empty(R, P, device=x2.device, dtype=x2.dtype)
    N.check(N.lib().stove_enc_head_fwd(R, K, J, P, N.ptr(x2), N.ptr(w1), N.ptr(b1), N.ptr(w2), N.ptr(b2),
                                       N.ptr(hidden), N.ptr(out), N.stream()))
    return hidden, out


def _head_bwd_data(x2, w1, w2, hidden, g_out):
    """-> (g_x, ws); ws carries the pre-activation gradients to _head_bwd_params."""
    R, K = x2.shape
    J, P = w1.shape[0], w2.shape[0]
    g_x = torch.empty_like(x2)
    ws = torch.empty(max(N.lib().stove_enc_head_bwd_workspace(R, K, J, P), 16) // 4, device=x2.device,
                     dtype=torch.float32)
    N.check(N.lib().stove_enc_head_bwd_data(R, K, J, P, N.ptr(w1), N.ptr(w2), N.ptr(hidden), N.ptr(g_out),
                                            N.ptr(g_x), N.ptr(ws), N.stream()))
    return g_x, ws


def _head_bwd_params(x2, w1, w2, hidden, g_out, ws):
    """Parameter gradients of the head on the CURRENT stream (must be ordered behind _head_bwd_data)."""
    R, K = x2.shape
    J, P = w1.shape[0], w2.shape[0]
    dev, dt = x2.device, x2.dtype
    g_w1, g_b1 = torch.empty_like(w1), torch.empty(J, device=dev, dtype=dt)
    g_w2, g_b2 = torch.empty_like(w2), torch.empty(P, device=dev, dtype=dt)
    N.check(N.lib().stove_enc_head_bwd_params(R, K, J, P, N.ptr(x2), N.ptr(hidden), N.ptr(g_out), N.ptr(g_w1),
                                              N.ptr(g_b1), N.ptr(g_w2), N.ptr(g_b2), N.ptr(ws), N.stream()))
    return g_w1, g_b1, g_w2, g_b2


class EncHead(torch.autograd.Function):
    """fc2(sigmoid(fc1(x))) of the recognition network (encoder.py:53-56) in one kernel forward and three
    backward (csrc/enc_head.cu).  x (..., K); w1 (J, K), b1 (J,), w2 (P, J), b2 (P,) -> (..., P).
    Inside the fused recognition network (LstmEncoder with head parameters) the same kernels are used and the
    parameter-gradient part runs beside the LSTM backward chain."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        w1, b1, w2, b2 = w1.contiguous(), b1.contiguous(), w2.contiguous(), b2.contiguous()
        N.require_cuda_f32(x2, w1, b1, w2, b2)
        hidden, out = _head_fwd(x2, w1, b1, w2, b2)
        ctx.save_for_backward(x2, w1, w2, hidden)
        ctx.lead = lead
        return out.view(*lead, w2.shape[0])

    @staticmethod
    def backward(ctx, g_out):
        x2, w1, w2, hidden = ctx.saved_tensors
        g_out = g_out.reshape(x2.shape[0], w2.shape[0]).contiguous()
        g_x, ws = _head_bwd_data(x2, w1, w2, hidden, g_out)
        g_w1, g_b1, g_w2, g_b2 = _head_bwd_params(x2, w1, w2, hidden, g_out, ws)
        return g_x.view(*ctx.lead, x2.shape[1]), g_w1, g_b1, g_w2, g_b2


def gather_flat(tensors, out=None, offsets=None, scale=1.0):
    """Copy the flattened fp32 CUDA tensors into one flat buffer with one or two launches of the library's gather
    kernel (the data-parallel gradient bucket), multiplying by `scale` on the way.  `offsets` (floats, one per
    tensor) places them inside `out`; default: back to back from 0.  Returns the flat tensor."""
    tensors = [t.contiguous() for t in tensors]
    N.require_cuda_f32(*tensors)
    numels = [t.numel() for t in tensors]
    if offsets is None:
        offsets, at = [], 0
        for m in numels:
            offsets.append(at)
            at += m
    if out is None:
        out = torch.empty(sum(numels), device=tensors[0].device, dtype=tensors[0].dtype)
    n = len(tensors)
    srcs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
    offs = (C.c_int64 * n)(*offsets)
    nums = (C.c_int64 * n)(*numels)
    N.check(N.lib().stove_gather_flat(C.cast(srcs, C.c_void_p), C.cast(offs, C.c_void_p), C.cast(nums, C.c_void_p),
                                      n, N.ptr(out), float(scale), N.stream()))
    return out


# ----------------------------------------------------------------------------------------
# GNN dynamics
# ----------------------------------------------------------------------------------------
GNN_SEGMENTS = ['act', 'enc', 'self0', 'self1', 'ra0', 'rel1', 'att1', 'rel2', 'att2', 'aff0', 'aff1',
                'aff2', 'out0', 'out1', 'rew00', 'rew02', 'rew10', 'rew12', 'rew14']


def gnn_weight_offsets(cfg):
    """{'enc_w': off, 'enc_b': off, ..., 'total': n} from the library (single source of truth)."""
    buf = (C.c_int32 * 64)()
    cnt = N.lib().stove_gnn_weight_offsets(C.byref(cfg), C.cast(buf, C.c_void_p), 64)
    if cnt < 0:
        N.check(cnt)
    names = [s + sfx for s in GNN_SEGMENTS for sfx in ('_w', '_b')] + ['total']
    assert cnt == len(names), (cnt, len(names))
    return {k: int(buf[i]) for i, k in enumerate(names)}


class GnnStep(torch.autograd.Function):
    """s (n, O, cl/2) [, actions (n, A), app (n, O, 3)], flat weights -> out (n, O, cl), reward (n, 1)."""

    @staticmethod
    def forward(ctx, s, actions, app, weights, cfg):
        s, actions, app = _c(s), _c(actions), _c(app)
        N.require_cuda_f32(s, actions, app, weights)
        n, O, _ = s.shape
        out = torch.empty(n, O, cfg.cl, device=s.device, dtype=s.dtype)
        reward = torch.empty(n, 1, device=s.device, dtype=s.dtype) if cfg.reward else None
        N.check(N.lib().stove_gnn_fwd(C.byref(cfg), n, N.ptr(s), N.ptr(actions), N.ptr(app), N.ptr(weights),
                                      N.ptr(out), N.ptr(reward), N.stream()))
        ctx.save_for_backward(s, actions, app, weights)
        ctx.cfg = cfg
        if reward is None:
            reward = s.new_zeros(())
            ctx.mark_non_differentiable(reward)
        return out, reward

    @staticmethod
    def backward(ctx, g_out, g_reward):
        s, actions, app, weights = ctx.saved_tensors
        cfg = ctx.cfg
        n = s.shape[0]
        g_out = g_out.contiguous()
        g_reward = _c(g_reward) if cfg.reward and g_reward is not None else None
        g_s = torch.empty_like(s)
        g_w = torch.empty_like(weights)
        ws = torch.empty(max(N.lib().stove_gnn_bwd_workspace(C.byref(cfg), n), 4) // 4, device=s.device,
                         dtype=torch.float32)
        N.check(N.lib().stove_gnn_bwd(C.byref(cfg), n, N.ptr(s), N.ptr(actions), N.ptr(app), N.ptr(weights),
                                      N.ptr(g_out), N.ptr(g_reward), N.ptr(g_s), N.ptr(g_w), N.ptr(ws),
                                      N.stream()))
        return g_s, None, None, g_w, None


class DynamicsLoop(torch.autograd.Function):
    """The whole dynamics loop of Stove.stove_forward (stove.py:696-713) in one library call: for
    every t >= skip the fused step (GNN + constrain + fusion with SuPAIR + sample + log q +
    transition lik), chained through z_t on the device -- one persistent kernel forward, three
    kernels backward (csrc/dynloop.cu).

    sup, sup_std (n, T, O, 6); lat0 (n, O, Z-6) initial latents (noise, no gradient) -- the initial
    state is [sup[:, skip-1], lat0] (stove.py:672-676); eps (S, n, O, Z); actions (n, T, A) | None;
    app (n, T, O, 3) | None; weights flat.  Returns z (n, S, O, Z), z_dyn, z_dyn_std (n, S, O, Z-2),
    z_std (n, S, O, Z), logq (n, S), trans (n, S), rewards (n, S, 1)."""

    @staticmethod
    def _io(T, skip, z_init, sup, sup_std, eps, actions, app):
        io = N.DynloopIO()
        io.T, io.skip = T, skip
        io.z_init, io.sup, io.sup_std, io.eps = N.ptr(z_init), N.ptr(sup), N.ptr(sup_std), N.ptr(eps)
        io.actions, io.app = N.ptr(actions), N.ptr(app)
        return io

    @staticmethod
    def forward(ctx, sup, sup_std, lat0, eps, actions, app, weights, cfg, fuse, skip, wgrad_stream=None):
        ctx.set_materialize_grads(False)       # undefined gradients arrive as None, not as zero-filled tensors
        sup, sup_std, eps = sup.contiguous(), sup_std.contiguous(), eps.contiguous()
        z_init = torch.cat([sup[:, skip - 1], lat0], -1)
        actions, app = _c(actions), _c(app)
        N.require_cuda_f32(z_init, sup, sup_std, eps, actions, app, weights)
        n, O, Z = z_init.shape
        T = sup.shape[1]
        S = T - skip
        dev, dt = z_init.device, z_init.dtype
        z = torch.empty(n, S, O, Z, device=dev, dtype=dt)
        z_dyn = torch.empty(n, S, O, Z - 2, device=dev, dtype=dt)
        z_dyn_std = torch.empty_like(z_dyn)
        z_std = torch.empty_like(z)
        logq = torch.empty(n, S, device=dev, dtype=dt)
        trans = torch.empty(n, S, device=dev, dtype=dt)
        rewards = torch.empty(n, S, 1, device=dev, dtype=dt) if cfg.reward else None
        io = DynamicsLoop._io(T, skip, z_init, sup, sup_std, eps, actions, app)
        io.z, io.z_dyn, io.z_dyn_std, io.z_std = N.ptr(z), N.ptr(z_dyn), N.ptr(z_dyn_std), N.ptr(z_std)
        io.logq, io.trans, io.reward = N.ptr(logq), N.ptr(trans), N.ptr(rewards)
        # training: the forward kernel keeps each step's activations so the backward chain does not recompute
        xrec = None
        if any(ctx.needs_input_grad):
            nx = N.lib().stove_dynloop_xrec_floats(C.byref(cfg), n, T, skip)
            if nx > 0:
                xrec = torch.empty(nx, device=dev, dtype=torch.float32)
        io.xrec = N.ptr(xrec)
        N.check(N.lib().stove_dynloop_fwd(C.byref(cfg), C.byref(fuse), n, C.byref(io), N.ptr(weights), N.stream()))
        ctx.xrec = xrec
        ctx.save_for_backward(z_init, sup, sup_std, eps, actions, app, weights, z)
        ctx.meta = (cfg, fuse, skip)
        # stream on which `weights` was packed: its backward runs there too, so the weight-gradient
        # kernels can leave the main chain (ops.DynamicsLoop.backward)
        ctx.wgrad_stream = wgrad_stream
        ctx.mark_non_differentiable(z_dyn, z_dyn_std, z_std)
        if rewards is None:
            rewards = z_init.new_zeros(())
            ctx.mark_non_differentiable(rewards)
        return z, z_dyn, z_dyn_std, z_std, logq, trans, rewards

    @staticmethod
    def backward(ctx, g_z, _g1, _g2, _g3, g_logq, g_trans, g_rewards):
        z_init, sup, sup_std, eps, actions, app, weights, z = ctx.saved_tensors
        cfg, fuse, skip = ctx.meta
        n, O, Z = z_init.shape
        T = sup.shape[1]
        dev, dt = z_init.device, z_init.dtype
        g_z, g_logq, g_trans = _c(g_z), _c(g_logq), _c(g_trans)
        g_rewards = _c(g_rewards) if (cfg.reward and g_rewards is not None and g_rewards.dim() > 0) else None
        g_z_init = torch.empty_like(z_init)
        g_sup = torch.empty_like(sup)
        g_sup_std = torch.empty_like(sup_std)
        g_w = torch.empty_like(weights)
        nbytes = N.lib().stove_dynloop_bwd_workspace(C.byref(cfg), n, T, skip)
        ws = torch.empty(max(nbytes, 16) // 4, device=dev, dtype=torch.float32)
        io = DynamicsLoop._io(T, skip, z_init, sup, sup_std, eps, actions, app)
        io.z = N.ptr(z)
        io.g_z, io.g_logq, io.g_trans, io.g_reward = N.ptr(g_z), N.ptr(g_logq), N.ptr(g_trans), N.ptr(g_rewards)
        io.g_z_init, io.g_sup, io.g_sup_std = N.ptr(g_z_init), N.ptr(g_sup), N.ptr(g_sup_std)
        io.xrec = N.ptr(ctx.xrec)
        cur = torch.cuda.current_stream(dev)
        side = ctx.wgrad_stream if ctx.wgrad_stream is not None else cur
        N.check(N.lib().stove_dynloop_bwd2(C.byref(cfg), C.byref(fuse), n, C.byref(io), N.ptr(weights), N.ptr(g_w),
                                           N.ptr(ws), cur.cuda_stream, side.cuda_stream))
        if side is not cur:
            # g_w is consumed by the backward of the weight packing, which autograd runs on `side` (the
            # stream of its forward); everything issued so far on `cur` is ordered before it
            side.wait_stream(cur)
            for t in (g_w, ws) + ((ctx.xrec,) if ctx.xrec is not None else ()):
                t.record_stream(side)
        # (the library also routes g_z_init[..., :6] into g_sup[:, skip-1]: the initial state is that slice)
        return g_sup, g_sup_std, None, None, None, None, g_w, None, None, None, None


class ZAll(torch.autograd.Function):
    """z of every scored frame (stove.py:731-736): z_sup (n, T, O, 4) for 1 <= t < skip, the sampled
    z_s (n, S, O, Z)[..., :4] for t >= skip, converted [sx, sy/sx, x, y] -> [sx, sy, x, y]
    (supair.py:151-158).  Returns (n, T-1, O, 4)."""

    @staticmethod
    def forward(ctx, z_sup, z_s, skip):
        z_sup, z_s = z_sup.contiguous(), z_s.contiguous()
        N.require_cuda_f32(z_sup, z_s)
        n, T, O, _ = z_sup.shape
        Z = z_s.shape[-1]
        z_all = torch.empty(n, T - 1, O, 4, device=z_sup.device, dtype=z_sup.dtype)
        N.check(N.lib().stove_zall_fwd(n, T, skip, O, Z, N.ptr(z_sup), N.ptr(z_s), N.ptr(z_all), N.stream()))
        ctx.save_for_backward(z_sup, z_s)
        ctx.skip = skip
        return z_all

    @staticmethod
    def backward(ctx, g_z_all):
        z_sup, z_s = ctx.saved_tensors
        n, T, O, _ = z_sup.shape
        g_z_sup, g_z_s = torch.empty_like(z_sup), torch.empty_like(z_s)
        N.check(N.lib().stove_zall_bwd(n, T, ctx.skip, O, z_s.shape[-1], N.ptr(z_sup), N.ptr(z_s),
                                       N.ptr(g_z_all.contiguous()), N.ptr(g_z_sup), N.ptr(g_z_s), N.stream()))
        return g_z_sup, g_z_s, None


class ElboAssemble(torch.autograd.Function):
    """Sequence ELBO from the per-frame terms (supair.py:84-110, stove.py:737-748), one launch each way.
    bg (F,), patch (F*O,) raw object-SPN log-likelihoods, z_all (n, T-1, O, 4), overlap (F, O),
    logq / trans (n, S).  Returns (average_elbo (), stats (8,)): stats = [elbo, mean bg, mean patch,
    mean overlap prior, mean log q, mean transition lik, mean SuPAIR-frame lik, 0] for logging."""

    @staticmethod
    def forward(ctx, bg, patch, z_all, overlap, logq, trans, skip, beta):
        ctx.set_materialize_grads(False)          # `stats` is a by-product: no zero gradient for it
        bg, patch, z_all, overlap = bg.contiguous(), patch.contiguous(), z_all.contiguous(), overlap.contiguous()
        logq, trans = logq.contiguous(), trans.contiguous()
        N.require_cuda_f32(bg, patch, z_all, overlap, logq, trans)
        n, Tm1, O, _ = z_all.shape
        stats = torch.empty(8, device=bg.device, dtype=bg.dtype)
        elbo = torch.empty((), device=bg.device, dtype=bg.dtype)
        N.check(N.lib().stove_elbo_fwd(n, Tm1 + 1, skip, O, beta, N.ptr(bg), N.ptr(patch), N.ptr(z_all),
                                       N.ptr(overlap), N.ptr(logq), N.ptr(trans), N.ptr(stats), N.ptr(elbo),
                                       N.stream()))
        ctx.save_for_backward(patch, z_all)
        ctx.meta = (skip, beta, bg.shape, overlap.shape, logq.shape)
        ctx.mark_non_differentiable(stats)
        return elbo, stats

    @staticmethod
    def backward(ctx, g_elbo, _g_stats):
        if g_elbo is None:
            return (None,) * 8
        patch, z_all = ctx.saved_tensors
        skip, beta, bg_shape, ov_shape, lq_shape = ctx.meta
        n, Tm1, O, _ = z_all.shape
        dev, dt = patch.device, patch.dtype
        g_bg = torch.empty(bg_shape, device=dev, dtype=dt)
        g_patch = torch.empty_like(patch)
        g_z_all = torch.empty_like(z_all)
        g_ov = torch.empty(ov_shape, device=dev, dtype=dt)
        g_lq = torch.empty(lq_shape, device=dev, dtype=dt)
        g_tr = torch.empty(lq_shape, device=dev, dtype=dt)
        N.check(N.lib().stove_elbo_bwd(n, Tm1 + 1, skip, O, beta, N.ptr(g_elbo.contiguous()), N.ptr(patch),
                                       N.ptr(z_all), N.ptr(g_bg), N.ptr(g_patch), N.ptr(g_z_all), N.ptr(g_ov),
                                       N.ptr(g_lq), N.ptr(g_tr), N.stream()))
        return g_bg, g_patch, g_z_all, g_ov, g_lq, g_tr, None, None


def gnn_rollout(cfg, z_last, num, weights, actions=None, app=None, noise=None, pos_var=0.3, vel_std=0.04,
                latent_std=0.04, want_std=False):
    """Persistent rollout kernel (stove.py:777-861).  Returns z (n, num, O, cl/2+2), std, logq, rewards."""
    z_last, actions, app, noise = _c(z_last), _c(actions), _c(app), _c(noise)
    N.require_cuda_f32(z_last, actions, app, noise, weights)
    n, O, zd = z_last.shape
    dev, dt = z_last.device, z_last.dtype
    z = torch.empty(n, num, O, zd, device=dev, dtype=dt)
    std = torch.empty(n, num, O, zd - 2, device=dev, dtype=dt) if (want_std or noise is not None) else None
    logq = torch.empty(n, num, O, zd - 2, device=dev, dtype=dt) if noise is not None else None
    rewards = torch.empty(n, num, 1, device=dev, dtype=dt) if cfg.reward else None
    alen = actions.shape[1] if actions is not None else 1
    N.check(N.lib().stove_gnn_rollout(C.byref(cfg), n, num, N.ptr(z_last), N.ptr(actions), alen, N.ptr(app),
                                      N.ptr(weights), N.ptr(noise), pos_var, vel_std, latent_std, N.ptr(z),
                                      N.ptr(std), N.ptr(logq), N.ptr(rewards), N.stream()))
    return z, std, logq, rewards
