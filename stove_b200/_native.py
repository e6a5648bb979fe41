"""ctypes binding of the C ABI declared in include/stove_b200.h.

There is deliberately no fallback: if the library is missing or a call fails, a
RuntimeError is raised.  Nothing here imports `oracle/`.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'csrc', 'libstove_b200.so')
_lib = None

c_f = C.c_void_p          # device pointers travel as integers
i32, i64, f32, vp, sz = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_size_t


class Spn2Struct(C.Structure):
    _fields_ = [('D', i32), ('R', i32), ('G', i32), ('S', i32), ('pmax', i32),
                ('region_scope', vp), ('region_n0', vp), ('region_n', vp), ('pix_slot', vp)]


class Spn1Struct(C.Structure):
    _fields_ = [('D', i32), ('R', i32), ('G', i32), ('side', vp)]


class SceneSeq(C.Structure):                      # stove_scene_seq
    _fields_ = [('n', i64), ('T', i32), ('skip', i32), ('Z', i32), ('beta', f32),
                ('z_sup', vp), ('z_s', vp), ('patch_w', vp), ('g_elbo', vp), ('g_z_sup', vp), ('g_z_s', vp),
                ('g_logq', vp), ('g_trans', vp)]


class SupCfg(C.Structure):
    _fields_ = [('T', i32), ('num_obj', i32), ('match_kind', i32), ('app_dim', i32),
                ('match_appearance', i32), ('fix_supair', i32), ('min_obj_scale', f32),
                ('max_obj_scale', f32), ('min_y_scale', f32), ('max_y_scale', f32), ('obj_pos_bound', f32),
                ('scale_var', f32), ('pos_var', f32), ('fix_eps', f32)]


class GnnCfg(C.Structure):
    _fields_ = [('num_obj', i32), ('cl', i32), ('action_dim', i32), ('app_dim', i32),
                ('reward', i32), ('lim_enc', i32), ('nonlin', i32), ('state_dim', i32)]


class FuseCfg(C.Structure):
    _fields_ = [('pos_var', f32), ('vel_std', f32), ('latent_std', f32), ('trans_std', f32 * 32)]


class DynstepIO(C.Structure):
    _fields_ = [('z_prev', vp), ('z_prev_ss', i64),
                ('sup', vp), ('sup_std', vp), ('sup_ss', i64),
                ('eps', vp), ('eps_ss', i64),
                ('actions', vp), ('act_ss', i64),
                ('app', vp), ('app_ss', i64),
                ('z_out', vp), ('z_out_ss', i64),
                ('z_dyn', vp), ('z_dyn_std', vp), ('zdyn_ss', i64),
                ('z_std', vp), ('z_std_ss', i64),
                ('logq', vp), ('trans', vp), ('reward', vp), ('sc_ss', i64),
                ('g_z_a', vp), ('g_z_a_ss', i64),
                ('g_z_b', vp), ('g_z_b_ss', i64),
                ('g_logq', vp), ('g_trans', vp), ('g_reward', vp), ('g_sc_ss', i64),
                ('g_z_prev', vp), ('g_z_prev_ss', i64),
                ('g_sup', vp), ('g_sup_std', vp), ('g_sup_ss', i64)]


class DynloopIO(C.Structure):
    _fields_ = [('T', i32), ('skip', i32),
                ('z_init', vp), ('sup', vp), ('sup_std', vp), ('eps', vp), ('actions', vp), ('app', vp),
                ('z', vp), ('z_dyn', vp), ('z_dyn_std', vp), ('z_std', vp),
                ('logq', vp), ('trans', vp), ('reward', vp),
                ('g_z', vp), ('g_logq', vp), ('g_trans', vp), ('g_reward', vp),
                ('g_z_init', vp), ('g_sup', vp), ('g_sup_std', vp), ('xrec', vp)]


P2, P1, PG, PS = C.POINTER(Spn2Struct), C.POINTER(Spn1Struct), C.POINTER(GnnCfg), C.POINTER(SupCfg)

# name -> (restype, argtypes); must list every symbol of include/stove_b200.h
SIGNATURES = {
    'stove_last_error': (C.c_char_p, []),
    'stove_abi_version': (C.c_int, []),
    'stove_launch_count': (i64, [C.c_int]),
    'stove_profile_enable': (C.c_int, [C.c_int]),
    'stove_profile_read': (C.c_int, [vp, vp, C.c_int]),
    'stove_kernel_name': (C.c_char_p, [C.c_int]),
    'stove_kernel_count': (C.c_int, []),
    'stove_microbench_ffma': (C.c_int, [C.c_int, C.c_int, vp, C.POINTER(C.c_double), vp]),
    'stove_microbench_tf32': (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), vp]),
    'stove_render': (C.c_int, [i64] + [C.c_int] * 7 + [vp, C.c_int, vp, C.c_int, vp, vp, vp]),
    'stove_set_option': (C.c_int, [C.c_char_p, C.c_int]),
    'stove_get_option': (C.c_int, [C.c_char_p]),
    'stove_bw_transform': (C.c_int, [vp, vp, i64, C.c_int, i64, vp]),
    'stove_bw_transform_ex': (C.c_int, [vp, C.c_int, C.c_int, vp, vp, i64, C.c_int, i64, vp]),
    'stove_spn_pack_leaf_fwd': (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, f32, f32, vp, vp]),
    'stove_spn_pack_leaf_bwd': (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, f32, f32, vp, vp, vp, vp]),
    'stove_spn_pack_sum_fwd': (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    'stove_spn_pack_sum_bwd': (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    'stove_spn2_fwd': (C.c_int, [P2, i64] + [vp] * 10 + [vp]),
    'stove_spn2_bwd_workspace': (sz, [P2, i64]),
    'stove_spn2_bwd': (C.c_int, [P2, i64] + [vp] * 17 + [vp, vp]),
    'stove_spn1_fwd_workspace': (sz, [P1, i64]),
    'stove_spn1_fwd': (C.c_int, [P1, i64] + [vp] * 8 + [vp]),
    'stove_spn1_bwd_workspace': (sz, [P1, i64]),
    'stove_spn1_bwd': (C.c_int, [P1, i64] + [vp] * 13 + [vp, vp]),
    'stove_scene_fwd': (C.c_int, [i64] + [C.c_int] * 7 + [vp] * 6 + [vp]),
    'stove_scene_bwd': (C.c_int, [i64] + [C.c_int] * 7 + [vp] * 7 + [vp]),
    'stove_scene_ll_supported': (C.c_int, [i64] + [C.c_int] * 6 + [P2, P1]),
    'stove_spn_interleave_leaf': (C.c_int, [vp, vp, i64, vp, vp]),
    'stove_scene_ll_fwd': (C.c_int, [i64] + [C.c_int] * 6 + [vp, vp, P2] + [vp] * 5 + [P1] + [vp] * 5 + [vp, C.c_int] + [vp] * 9 + [C.POINTER(SceneSeq)] + [vp]),
    'stove_scene_ll_bwd': (C.c_int, [i64] + [C.c_int] * 6 + [vp, vp, P2] + [vp] * 5 + [P1] + [vp] * 5 + [vp, C.c_int] + [vp] * 8 + [vp] * 3
                           + [vp] * 6 + [vp, vp] + [C.POINTER(SceneSeq)] + [vp, vp, vp]),
    'stove_sup_prepare_fwd': (C.c_int, [PS, i64] + [vp] * 8 + [vp]),
    'stove_sup_prepare_bwd': (C.c_int, [PS, i64] + [vp] * 8 + [vp]),
    'stove_lstm_gemm_cell_fwd': (C.c_int, [i64, C.c_int, i64, vp, vp, vp, C.c_int, vp, vp, vp, i64, vp, vp, vp, vp, i64, i64, i64, vp]),
    'stove_split_planes': (C.c_int, [i64, C.c_int, vp, i64, vp, vp, i64, vp]),
    'stove_tc3_gemm': (C.c_int, [i64, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, C.c_int, i64, C.c_int, vp]),
    'stove_tc3_gemm_parts': (C.c_int, [i64, i64, i64, C.c_int]),
    'stove_lstm_cell_bwd_t': (C.c_int, [i64, C.c_int, vp, vp, vp, vp, i64, vp, C.c_int, vp, vp, vp, i64, i64, i64, vp, C.c_int, C.c_int, vp, vp, vp]),
    'stove_sum_parts': (C.c_int, [i64, C.c_int, i64, vp, vp, C.c_float, vp]),
    'stove_enc_head_fwd': (C.c_int, [i64, C.c_int, C.c_int, C.c_int] + [vp] * 7 + [vp]),
    'stove_enc_head_bwd_workspace': (sz, [i64, C.c_int, C.c_int, C.c_int]),
    'stove_enc_head_bwd_data': (C.c_int, [i64, C.c_int, C.c_int, C.c_int] + [vp] * 6 + [vp]),
    'stove_enc_head_bwd_params': (C.c_int, [i64, C.c_int, C.c_int, C.c_int] + [vp] * 8 + [vp]),
    'stove_gather_flat': (C.c_int, [vp, vp, vp, C.c_int, vp, C.c_float, vp]),
    'stove_adam_workspace_floats': (C.c_int, []),
    'stove_adam_step': (C.c_int, [vp, vp, vp, C.c_int, i64, vp, vp, vp, vp, vp, vp, vp, f32, f32, f32, f32, vp]),
    'stove_gnn_weight_count': (i64, [PG]),
    'stove_gnn_weight_offsets': (C.c_int, [PG, vp, C.c_int]),
    'stove_gnn_bwd_workspace': (sz, [PG, i64]),
    'stove_gnn_fwd': (C.c_int, [PG, i64] + [vp] * 6 + [vp]),
    'stove_gnn_bwd': (C.c_int, [PG, i64] + [vp] * 9 + [vp]),
    'stove_dynstep_fwd': (C.c_int, [PG, C.POINTER(FuseCfg), i64, C.POINTER(DynstepIO), vp, vp]),
    'stove_dynstep_bwd': (C.c_int, [PG, C.POINTER(FuseCfg), i64, C.POINTER(DynstepIO), vp, vp, C.c_int, C.c_int, vp, vp]),
    'stove_dynloop_fwd': (C.c_int, [PG, C.POINTER(FuseCfg), i64, C.POINTER(DynloopIO), vp, vp]),
    'stove_dynloop_bwd_workspace': (sz, [PG, i64, C.c_int, C.c_int]),
    'stove_dynloop_bwd2': (C.c_int, [PG, C.POINTER(FuseCfg), i64, C.POINTER(DynloopIO), vp, vp, vp, vp, vp]),
    'stove_dynloop_xrec_floats': (i64, [PG, i64, C.c_int, C.c_int]),
    'stove_zall_fwd': (C.c_int, [i64, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]),
    'stove_zall_bwd': (C.c_int, [i64, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]),
    'stove_elbo_fwd': (C.c_int, [i64, C.c_int, C.c_int, C.c_int, f32] + [vp] * 8 + [vp]),
    'stove_elbo_bwd': (C.c_int, [i64, C.c_int, C.c_int, C.c_int, f32] + [vp] * 9 + [vp]),
    'stove_gnn_rollout': (C.c_int, [PG, i64, C.c_int, vp, vp, C.c_int, vp, vp, vp, f32, f32, f32,
                                    vp, vp, vp, vp, vp]),
}


def lib():
    """Load (once) and return the native library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'stove_b200: native library %s is missing -- run `python -m stove_b200.build` '
                '(there is no CPU / PyTorch fallback for the hot path)' % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError('stove_b200 native call failed (%d): %s'
                           % (rc, lib().stove_last_error().decode()))


def set_option(name, value):
    """Library option (include/stove_b200.h: stove_set_option); returns the previous value."""
    prev = lib().stove_set_option(name.encode(), int(value))
    if prev < 0:
        check(prev)
    return prev


def profile_read(max_n=1 << 16):
    """[(kernel name, device ms), ...] recorded since the last read (profiling must be enabled)."""
    ids = (C.c_int32 * max_n)()
    ms = (C.c_float * max_n)()
    n = lib().stove_profile_read(C.cast(ids, C.c_void_p), C.cast(ms, C.c_void_p), max_n)
    return [(lib().stove_kernel_name(ids[i]).decode(), float(ms[i])) for i in range(n)]


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda_f32(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError('stove_b200 kernels need CUDA tensors (got a %s tensor); there is no '
                               'CPU path' % t.device)
        if t.dtype != torch.float32:
            raise RuntimeError('stove_b200 kernels compute in fp32 (got %s)' % t.dtype)
