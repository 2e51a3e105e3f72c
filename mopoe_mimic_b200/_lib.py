"""ctypes binding of libmopoe_b200.so (the C ABI declared in include/mopoe_b200.h).

The product has NO fallback: if the library is missing, or a call returns non-zero, a RuntimeError
is raised.  Nothing here imports the CPU oracle.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MOPOE_LIB_PATH') or os.path.join(_HERE, 'libmopoe_b200.so')     # (override: developer A/B builds)

F32, BF16 = 0, 1
MASK_NONE, MASK_BC, MASK_ELEM = 0, 1, 2
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2


class View(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('dtype', C.c_int32), ('B', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('C', C.c_int32), ('ph', C.c_int32), ('pw', C.c_int32), ('_pad', C.c_int32),
                ('sB', C.c_int64), ('sH', C.c_int64), ('sW', C.c_int64)]


class Window(C.Structure):
    _fields_ = [('a', C.c_void_p), ('a_dtype', C.c_int32), ('E0', C.c_int32), ('E1', C.c_int32), ('E2', C.c_int32),
                ('R', C.c_int32), ('KW', C.c_int32), ('_pad', C.c_int32),
                ('a_off', C.c_int64), ('sA0', C.c_int64), ('sA1', C.c_int64), ('sA2', C.c_int64), ('sAr', C.c_int64)]


class Rows(C.Structure):
    _fields_ = [('d', C.c_void_p), ('d_dtype', C.c_int32), ('N', C.c_int32),
                ('d_off', C.c_int64), ('s0', C.c_int64), ('s1', C.c_int64), ('s2', C.c_int64)]


class FusionCfg(C.Structure):
    _fields_ = [('M', C.c_int32), ('B', C.c_int32), ('D', C.c_int32), ('nsub', C.c_int32), ('S', C.c_int32),
                ('fuse_mode', C.c_int32), ('prior_expert', C.c_int32), ('kl_chunks', C.c_int32),
                ('members', C.c_int32 * 16), ('stacked', C.c_int32 * 16), ('sel_end', C.c_int32 * 16),
                ('mem_cnt', C.c_int32 * 16), ('mem_idx', (C.c_int32 * 4) * 16), ('mem_end', (C.c_int32 * 4) * 16), ('norm', C.c_float), ('_pad', C.c_float)]


class BnReq(C.Structure):
    _fields_ = [('out', View), ('mask', C.c_void_p), ('mask_mode', C.c_int32), ('nchunk', C.c_int32), ('ws', C.c_void_p),
                ('ws_doubles', C.c_int64), ('eps', C.c_float), ('momentum', C.c_float), ('mean', C.c_void_p),
                ('invstd', C.c_void_p), ('running_mean', C.c_void_p), ('running_var', C.c_void_p)]


class ResReq(C.Structure):
    _fields_ = [('r', C.POINTER(Rows)), ('mean', C.c_void_p), ('invstd', C.c_void_p), ('gamma', C.c_void_p),
                ('beta', C.c_void_p), ('a', C.c_float), ('b', C.c_float), ('mask', C.c_void_p), ('mask_mode', C.c_int32),
                ('out', C.POINTER(View))]


class BnBwdReq(C.Structure):
    _fields_ = [('x', C.POINTER(Rows)), ('mask', C.c_void_p), ('mask_mode', C.c_int32), ('accumulate', C.c_int32),
                ('mean', C.c_void_p), ('invstd', C.c_void_p), ('gamma', C.c_void_p), ('beta', C.c_void_p),
                ('ws', C.c_void_p), ('ws_doubles', C.c_int64), ('dgamma', C.c_void_p), ('dbeta', C.c_void_p),
                ('sums', C.c_void_p)]


class PackJob(C.Structure):
    _fields_ = [('W', C.c_void_p), ('dst', C.c_void_p * 4), ('A', C.c_int32), ('B', C.c_int32), ('KH', C.c_int32),
                ('KW', C.c_int32), ('form', C.c_int32), ('bpad', C.c_int32), ('tile0', C.c_int32), ('nx', C.c_int32)]


class DpPeers(C.Structure):
    _fields_ = [('grad', C.c_void_p * 16), ('param', C.c_void_p * 16), ('flags', C.c_void_p * 16)]


_P, _I, _F, _L, _S = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t
_V, _W, _R = C.POINTER(View), C.POINTER(Window), C.POINTER(Rows)

# name -> (restype, argtypes); every symbol include/mopoe_b200.h declares
SIGNATURES = {
    'mopoe_last_error': (C.c_char_p, []),
    'mopoe_version': (_I, []),
    'mopoe_tc_available': (_I, []),
    'mopoe_tc_wgrad_built': (_I, []),
    'mopoe_conv_gemm': (_I, [_W, _P, _P, _R, _I, _P]),
    'mopoe_conv_gemm_batched': (_I, [_I, _W, _P, _P, _R, _I, _P]),
    'mopoe_conv_gemm_bn': (_I, [_I, _W, _P, _P, _R, _I, C.POINTER(BnReq), _P]),
    'mopoe_conv_gemm_res_eligible': (_I, [_I, _W, _P, _R, _I, C.POINTER(ResReq), C.POINTER(BnReq)]),
    'mopoe_conv_gemm_res': (_I, [_I, _W, _P, _P, _R, _I, C.POINTER(ResReq), C.POINTER(BnReq), _P]),
    'mopoe_conv_gemm_bnbwd_eligible': (_I, [_I, _W, _R, _I, C.POINTER(BnBwdReq)]),
    'mopoe_conv_gemm_bnbwd': (_I, [_I, _W, _P, _R, _I, C.POINTER(BnBwdReq), _P]),
    'mopoe_conv_gemm_splitk_ws': (_S, [_W, _R, _I]),
    'mopoe_conv_gemm_splitk': (_I, [_W, _P, _P, _R, _P, _S, _I, _P]),
    'mopoe_conv_wgrad_ws': (_S, [_W, _R, _I]),
    'mopoe_conv_wgrad': (_I, [_W, _R, _P, _I, _P, _S, _I, _P]),
    'mopoe_colsum': (_I, [_V, _P, _I, _P, _I, _P, _P]),
    'mopoe_bn_stats': (_I, [_V, _P, _I, _P, _I, _F, _F, _P, _P, _P, _P, _P, _P]),
    'mopoe_bn_apply': (_I, [_V, _P, _I, _P, _P, _P, _P, _I, _V, _P]),
    'mopoe_combine': (_I, [_V, _P, _P, _P, _P, _V, _P, _I, _F, _F, _V, _P]),
    'mopoe_combine_bn': (_I, [_V, _P, _P, _P, _P, _V, _P, _I, _F, _F, _V, _P, _I, _F, _F, _P, _P, _P, _P, _P]),
    'mopoe_bn_bwd_reduce': (_I, [_V, _V, _F, _V, _P, _I, _P, _P, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P]),
    'mopoe_bn_bwd_apply': (_I, [_V, _V, _F, _V, _P, _I, _P, _P, _P, _P, _V, _V, _P, _P]),
    'mopoe_combine_bwd_apply': (_I, [_V, _F, _V, _P, _P, _P, _P, _P, _I, _F, _V, _V, _P]),
    'mopoe_scale_mask': (_I, [_V, _P, _I, _F, _V, _P]),
    'mopoe_convert': (_I, [_V, _I, _V, _P]),
    'mopoe_dropout_mask': (_I, [_P, _L, C.c_uint64, C.c_uint64, _P, _P]),
    'mopoe_pack_weight_tiled': (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _I, _P]),
    'mopoe_conv_wgrad_param_ws': (_S, [_W, _R, _I]),
    'mopoe_conv_wgrad_param': (_I, [_W, _R, _P, _I, _I, _I, _I, _I, _P, _S, _I, _P]),
    'mopoe_pack_weight': (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P]),
    'mopoe_step_advance': (_I, [_P, _P, _P, _F, _F, _F, _P]),
    'mopoe_adam_flat_dev': (_I, [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _P]),
    'mopoe_conv3x3s2_c1_fwd': (_I, [_P, _P, _I, _I, _I, _V, _P]),
    'mopoe_conv3x3s2_c1_wgrad': (_I, [_P, _V, _I, _I, _I, _P, _I, _P, _I, _P]),
    'mopoe_deconv3x3s2_c1_fwd_ws': (_S, [_V]),
    'mopoe_deconv3x3s2_c1_fwd': (_I, [_V, _P, _P, _P, _P, _S, _P]),
    'mopoe_deconv3x3s2_c1_bwd': (_I, [_V, _P, _P, _V, _P, _P, _I, _P, _I, _P]),
    'mopoe_im2col3x3s2': (_I, [_P, _I, _I, _I, _I, _P, _P]),
    'mopoe_deconv3x3s2_c1_assemble': (_I, [_P, _I, _P, _P, _I, _I, _I, _P]),
    'mopoe_zero_border': (_I, [_V, _P]),
    'mopoe_fusion_fwd': (_I, [C.POINTER(FusionCfg), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'mopoe_fusion_bwd': (_I, [C.POINTER(FusionCfg), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'mopoe_laplace_logprob_sum': (_I, [_P, _P, _L, _F, _P, _P, _I, _P]),
    'mopoe_laplace_logprob_bwd': (_I, [_P, _P, _L, _F, _P, _P, _P]),
    'mopoe_laplace_logprob_elem': (_I, [_P, _P, _L, _F, _P, _P]),
    'mopoe_categorical_logprob_sum': (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P, _P, _I, _P]),
    'mopoe_categorical_has_lse': (_I, [_P, _P, _I]),
    'mopoe_categorical_logprob_bwd': (_I, [_P, _P, _P, _L, _I, _P, _P, _P]),
    'mopoe_jsd_divergence_fwd': (_I, [_I, _I, _I, _P, _P, _P, _F, _P, _P, _P, _P, _P]),
    'mopoe_jsd_divergence_bwd': (_I, [_I, _I, _I, _P, _P, _P, _F, _P, _P, _P, _P]),
    'mopoe_onehot': (_I, [_P, _L, _I, _I, _P, _I, _P, _P]),
    'mopoe_onehot_u8': (_I, [_P, _L, _I, _P, _P]),
    'mopoe_text_stem_gather_fwd': (_I, [_P, _I, _I, _I, _P, _I, _P, _V, _P]),
    'mopoe_text_onehot_act': (_I, [_P, _I, _I, _I, _V, _P]),
    'mopoe_u8_to_unit': (_I, [_P, _L, _P, _P]),
    'mopoe_pack_job_tiles': (_I, [_I, _I, _I, _I, _I, _I, C.POINTER(C.c_int)]),
    'mopoe_pack_weights_batched': (_I, [_P, _I, _I, _I, _P]),
    'mopoe_dp_adam_exchange': (_I, [C.POINTER(DpPeers), _P, _P, _P, _P, _L, _I, _I, _P, _P, _F, _F, _F, _F, _P]),
    'mopoe_dp_adam_exchange_ex': (_I, [C.POINTER(DpPeers), _P, _P, _P, _P, _L, _I, _I, _P, _P, _F, _F, _F, _F, _I, _P]),
    'mopoe_adam_flat': (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _I, _F, _P]),
}

_lib = None
LAUNCHES = 0     # kernels-launching C calls issued (bench.py's gpu_launches evidence)


def load():
    """Load the shared library (no device needed) and bind every declared symbol."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('libmopoe_b200.so is missing (%s): build it with `python -m mopoe_mimic_b200.build` '
                               'or __graft_entry__.build(); there is no CPU fallback' % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


# bench.py's per-call device timing: when PROFILE is a list, every call is bracketed by a CUDA event pair on the current
# stream and appended as a dict(name, a, b, + the annotation set by annotate() just before the call: kind / flops / bytes /
# tag).  PROFILE_EXTERNAL: create the events as `external` so that they become event-record NODES of a graph being captured.
PROFILE = None
PROFILE_EXTERNAL = False
_ANNOT = None


def annotate(**kw):
    """algorithmic work of the NEXT call (consumed by it): kind=..., flops=..., bytes=..., tag=..."""
    global _ANNOT
    if PROFILE is not None:
        _ANNOT = kw


def call(name, *args):
    """Invoke an int-returning entry point; raise RuntimeError(mopoe_last_error()) on failure."""
    global LAUNCHES, _ANNOT
    lib = load()
    if PROFILE is not None:
        ext = {'external': True} if PROFILE_EXTERNAL else {}
        a, b = torch.cuda.Event(enable_timing=True, **ext), torch.cuda.Event(enable_timing=True, **ext)
        a.record()
        rc = getattr(lib, name)(*args)
        b.record()
        rec = {'name': name, 'a': a, 'b': b}
        if _ANNOT:
            rec.update(_ANNOT)
        _ANNOT = None
        PROFILE.append(rec)
    else:
        rc = getattr(lib, name)(*args)
    LAUNCHES += 1
    if rc != 0:
        raise RuntimeError('%s failed: %s' % (name, lib.mopoe_last_error().decode()))


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise TypeError('unsupported dtype %s' % dt)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('mopoe_mimic_b200 runs on CUDA only (sm_100a); got a %s tensor — there is no CPU '
                               'fallback' % t.device)
