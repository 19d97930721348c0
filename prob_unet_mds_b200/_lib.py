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


class PuConvGnBwd(C.Structure):
    _fields_ = [('x0', c_void_p), ('x1', c_void_p), ('C0', c_int), ('C1', c_int), ('consts', c_void_p),
                ('sums', c_void_p), ('silu', c_int), ('dropout_p', c_float), ('seed', c_ull), ('keep_mask', c_void_p)]


class PuConvArgs(C.Structure):
    _fields_ = [('N', c_int), ('H', c_int), ('W', c_int), ('C0', c_int), ('C1', c_int), ('Cout', c_int),
                ('ksize', c_int), ('dtype', c_int), ('flags', c_int), ('bias_per_sample', c_int),
                ('src0', c_void_p), ('src1', c_void_p), ('weight', c_void_p), ('bias', c_void_p),
                ('residual', c_void_p), ('out', c_void_p), ('qstats', c_void_p), ('reserved', c_int),
                ('gn_bwd', C.POINTER(PuConvGnBwd))]


class PuWgradArgs(C.Structure):
    _fields_ = [('N', c_int), ('H', c_int), ('W', c_int), ('C0', c_int), ('C1', c_int), ('Cout', c_int),
                ('ksize', c_int), ('dtype', c_int), ('flags', c_int),
                ('src0', c_void_p), ('src1', c_void_p), ('dy', c_void_p), ('dw', c_void_p), ('accumulate', c_int)]


class PuGnArgs(C.Structure):
    _fields_ = [('N', c_int), ('H', c_int), ('W', c_int), ('C0', c_int), ('C1', c_int), ('G', c_int),
                ('dtype', c_int), ('silu', c_int), ('resample', c_int), ('eps', c_float), ('dropout_p', c_float),
                ('seed', c_ull), ('src0', c_void_p), ('src1', c_void_p), ('stats', c_void_p), ('gamma', c_void_p),
                ('beta', c_void_p), ('ada', c_void_p), ('y', c_void_p), ('keep_mask', c_void_p)]


class PuGnBwdArgs(C.Structure):
    _fields_ = [('f', PuGnArgs), ('dy', c_void_p), ('dres', c_void_p), ('dres_resample', c_int),
                ('sums', c_void_p), ('dx0', c_void_p), ('dx1', c_void_p), ('acc0', c_int), ('acc1', c_int),
                ('dgamma', c_void_p), ('dbeta', c_void_p), ('dada', c_void_p), ('acc_params', c_int),
                ('colsum0', c_void_p), ('colsum1', c_void_p), ('du_ready', c_int)]


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
    'pu_zero': (c_int, [c_void_p, c_ll, c_void_p]),
    'pu_copy': (c_int, [c_void_p, c_void_p, c_ll, c_void_p]),
    'pu_nchw_to_nhwc': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_nhwc_to_nchw': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_pack_conv_weight': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_void_p]),
    'pu_pack_conv_weights_multi': (c_int, [c_void_p, c_int, c_int, c_void_p]),
    'pu_unpack_conv_wgrad': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_void_p]),
    'pu_gather_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'pu_scatter_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'pu_conv2d': (c_int, [C.POINTER(PuConvArgs), c_void_p]),
    'pu_conv2d_wgrad': (c_int, [C.POINTER(PuWgradArgs), c_void_p]),
    'pu_bias_grad': (c_int, [c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p]),
    'pu_gn_stats': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'pu_gn_stats_from_quads': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'pu_gn_apply': (c_int, [C.POINTER(PuGnArgs), c_void_p]),
    'pu_gn_bwd': (c_int, [C.POINTER(PuGnBwdArgs), c_void_p]),
    'pu_gn_bwd_consts': (c_int, [C.POINTER(PuGnArgs), c_void_p, c_void_p]),
    'pu_attention_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_attention_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                 c_int, c_int, c_int, c_int, c_void_p]),
    'pu_upsample2': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_resample_grad': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
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
    'pu_mse_fwd_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'pu_loss_finalize': (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    'pu_loss_bwd_scales': (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    'pu_fcomb_fwd': (c_int, [C.POINTER(PuFcombArgs), c_void_p]),
    'pu_fcomb_z_bwd': (c_int, [c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                               c_int, c_void_p]),
    'pu_adamw': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float, c_float, c_float,
                         c_int, c_void_p]),
    'pu_climex_prepare': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    'pu_climex_residual_to_hr': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_int, c_int, c_int,
                                         c_void_p, c_void_p]),
    'pu_crps_empirical': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_ll, c_ll, c_ll, c_ll, c_void_p]),
    'pu_adamw_multi': (c_int, [c_void_p, c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, c_int,
                               c_void_p]),
}

_lib = None


def lib():
    """The loaded C ABI (or its event-recording proxy while profiling is active)."""
    return _profiled if _profiled is not None else _raw_lib()


def _raw_lib():
    """Loads the shared library (building it first if it is missing and nvcc is available)."""
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


# ---------------------------------------------------------------------------------------------------------------------
# optional per-call timing (CUDA events on the launching stream) -- used by bench.py for the roofline numbers
# ---------------------------------------------------------------------------------------------------------------------
class CallProfiler:
    """While active, every pu_* call is bracketed by two CUDA events on the current stream.  For the convolution
    entry points the algorithmic FLOPs (2*M*N*K of the implicit GEMM) are recorded from the argument struct."""

    def __init__(self):
        self.records = []   # (name, kind, flops, start_event, end_event)

    @staticmethod
    def _conv_info(name, args):
        if name not in ('pu_conv2d', 'pu_conv2d_wgrad'):
            return name, 0.0
        a = args[0]._obj
        ctot = a.C0 + a.C1
        flops = 2.0 * a.N * a.H * a.W * a.Cout * a.ksize * a.ksize * ctot
        tc = (a.dtype == PU_BF16 and a.C0 % 64 == 0 and a.C1 % 64 == 0 and a.Cout % 64 == 0 and a.W >= 16
              and a.H >= (8 if name == 'pu_conv2d' else 4) and not (a.flags & CONV_FORCE_SIMPLE))
        kind = ('conv_tc' if name == 'pu_conv2d' else 'wgrad_tc') if tc else \
               ('conv_simple' if name == 'pu_conv2d' else 'wgrad_simple')
        return kind, flops

    def wrap(self, name, fn):
        def call(*args):
            kind, flops = self._conv_info(name, args)
            s = torch.cuda.Event(enable_timing=True)
            e = torch.cuda.Event(enable_timing=True)
            s.record()
            rc = fn(*args)
            e.record()
            self.records.append((name, kind, flops, s, e))
            return rc
        return call

    def summary(self):
        """{kind: dict(calls, ms, flops)} -- call after torch.cuda.synchronize()."""
        out = {}
        for name, kind, flops, s, e in self.records:
            d = out.setdefault(kind, dict(calls=0, ms=0.0, flops=0.0))
            d['calls'] += 1
            d['ms'] += s.elapsed_time(e)
            d['flops'] += flops
        return out


class _ProfiledLib:
    def __init__(self, handle, prof):
        self._h = handle
        self._p = prof
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw = getattr(self._h, name)
            fn = self._p.wrap(name, raw) if name in _SIGNATURES and name not in (
                'pu_last_error', 'pu_version', 'pu_launch_count', 'pu_device_supports_tc') else raw
            self._cache[name] = fn
        return fn


_profiled = None


def start_profiling():
    global _profiled
    prof = CallProfiler()
    _profiled = _ProfiledLib(_raw_lib(), prof)
    return prof


def stop_profiling():
    global _profiled
    _profiled = None


def check(rc, what=''):
    if rc != 0:
        msg = _raw_lib().pu_last_error().decode(errors='replace')
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
