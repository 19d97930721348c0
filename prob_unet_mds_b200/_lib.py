"""ctypes binding of libprobunet_b200.so (the C ABI declared in include/probunet_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libprobunet_b200.so')

PU_F32, PU_BF16 = 0, 1
RS_NONE, RS_UP, RS_DOWN = 0, 1, 2
CONV_RELU, CONV_FORCE_SIMPLE, CONV_FORCE_TC = 1, 2, 4

c_void_p, c_int, c_float, c_ll, c_ull = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_ulonglong


class PuConvArgs(C.Structure):
    _fields_ = [('N', c_int), ('H', c_int), ('W', c_int), ('C0', c_int), ('C1', c_int), ('Cout', c_int),
                ('ksize', c_int), ('dtype', c_int), ('flags', c_int), ('bias_per_sample', c_int),
                ('src0', c_void_p), ('src1', c_void_p), ('weight', c_void_p), ('bias', c_void_p),
                ('residual', c_void_p), ('out', c_void_p), ('gn_stats', c_void_p), ('gn_groups', c_int)]


class PuWgradArgs(C.Structure):
    _fields_ = [('N', c_int), ('H', c_int), ('W', c_int), ('C0', c_int), ('C1', c_int), ('Cout', c_int),
                ('ksize', c_int), ('dtype', c_int), ('flags', c_int),
                ('src0', c_void_p), ('src1', c_void_p), ('dy', c_void_p), ('dw', c_void_p), ('accumulate', c_int)]


class PuGnArgs(C.Structure):
    _fields_ = [('N', c_int), ('H', c_int), ('W', c_int), ('C0', c_int), ('C1', c_int), ('G', c_int),
                ('dtype', c_int), ('silu', c_int), ('resample', c_int), ('eps', c_float), ('dropout_p', c_float),
                ('seed', c_ull), ('src0', c_void_p), ('src1', c_void_p), ('stats', c_void_p), ('gamma', c_void_p),
                ('beta', c_void_p), ('ada', c_void_p), ('y', c_void_p)]


class PuGnBwdArgs(C.Structure):
    _fields_ = [('f', PuGnArgs), ('dy', c_void_p), ('dres', c_void_p), ('dres_resample', c_int),
                ('sums', c_void_p), ('dx0', c_void_p), ('dx1', c_void_p), ('acc0', c_int), ('acc1', c_int),
                ('dgamma', c_void_p), ('dbeta', c_void_p), ('dada', c_void_p), ('acc_params', c_int)]


class PuFcombArgs(C.Structure):
    _fields_ = [('N', c_int), ('HW', c_int), ('L', c_int), ('dtype', c_int), ('S', c_int), ('num_classes', c_int),
                ('feat', c_void_p), ('z', c_void_p), ('w0', c_void_p), ('b0', c_void_p), ('w1', c_void_p),
                ('b1', c_void_p), ('w2', c_void_p), ('b2', c_void_p), ('out_nchw', c_void_p),
                ('h1_out', c_void_p), ('h2_out', c_void_p)]


_SIGNATURES = {
    'pu_last_error': (C.c_char_p, []),
    'pu_version': (c_int, []),
    'pu_device_supports_tc': (c_int, []),
    'pu_launch_count': (c_ll, [c_int]),
    'pu_nchw_to_nhwc': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_nhwc_to_nchw': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_pack_conv_weight': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_void_p]),
    'pu_unpack_conv_wgrad': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_void_p]),
    'pu_gather_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'pu_scatter_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'pu_conv2d': (c_int, [C.POINTER(PuConvArgs), c_void_p]),
    'pu_conv2d_wgrad': (c_int, [C.POINTER(PuWgradArgs), c_void_p]),
    'pu_bias_grad': (c_int, [c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p]),
    'pu_gn_stats': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'pu_gn_apply': (c_int, [C.POINTER(PuGnArgs), c_void_p]),
    'pu_gn_bwd': (c_int, [C.POINTER(PuGnBwdArgs), c_void_p]),
    'pu_attention_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_attention_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_int, c_int, c_void_p]),
    'pu_upsample2': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_avgpool2': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_relu_pool_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_global_mean': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_relu_mean_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_relu_mask': (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    'pu_heads_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'pu_heads_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                             c_int, c_void_p]),
    'pu_rsample': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'pu_rsample_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'pu_kl_fwd_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_int, c_void_p]),
    'pu_mse_fwd_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_loss_finalize': (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    'pu_loss_bwd_scales': (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    'pu_fcomb_fwd': (c_int, [C.POINTER(PuFcombArgs), c_void_p]),
    'pu_fcomb_z_bwd': (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                               c_int, c_void_p]),
    'pu_adamw': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float, c_float, c_float,
                         c_int, c_void_p]),
}

_lib = None


def lib():
    """Loads the shared library (building it first if the sources are newer and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from . import build as _build
            _build.build()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(
                f'probunet_b200: {LIB_PATH} is missing and could not be built ({e}). '
                'Run `python -m prob_unet_mds_b200.build`; there is no CPU or PyTorch fallback.') from e
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)   # AttributeError if the ABI is incomplete: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc, what=''):
    if rc != 0:
        msg = lib().pu_last_error().decode(errors='replace')
        if rc == -1:
            raise ValueError(f'probunet_b200 {what}: {msg}')
        raise RuntimeError(f'probunet_b200 {what} failed (rc={rc}): {msg}')


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def dtype_code(t):
    if t == torch.float32:
        return PU_F32
    if t == torch.bfloat16:
        return PU_BF16
    raise ValueError(f'unsupported dtype {t}')
