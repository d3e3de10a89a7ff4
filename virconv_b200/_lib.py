"""ctypes binding of libvirconv_sm100.so (include/virconv_b200.h).

The library is the product: there is no CPU / PyTorch fallback.  If the shared object is missing the
first op that needs it raises — loudly — instead of computing anything elsewhere.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('VIRCONV_LIB') or os.path.join(_HERE, 'lib', 'libvirconv_sm100.so')

_P = c_void_p
_I = c_int
_F = c_float
_Z = c_size_t
_HOST = ctypes.POINTER(c_int32)
_HOSTF = ctypes.POINTER(c_float)

# name -> (restype, argtypes).  Every symbol include/virconv_b200.h declares.
SIGNATURES = {
    'vc_version': (_I, []),
    'vc_last_error': (c_char_p, []),
    'vc_launch_count': (ctypes.c_longlong, []),
    'vc_set_pdl': (_I, [_I]),
    'vc_set_bn_fused': (_I, [_I]),
    'vc_set_tc_variant': (_I, [_I]),
    'vc_conv_tc2_config': (_I, [_I]),
    'vc_conv_wgrad_tc3_config': (_I, [_I, _I]),
    'vc_conv_wgrad_tc2_config': (_I, [_I]),
    'vc_subm_rulebook_ws_bytes': (_Z, [_I]),
    'vc_subm_rulebook': (_I, [_P, _I, _I, _I, _HOST, _HOST, _HOST, _P, _P, _P, _Z, _P]),
    'vc_conv_rulebook_ws_bytes': (_Z, [_I, _I, _HOST]),
    'vc_conv_out_shape': (_I, [_I, _HOST, _HOST, _HOST, _HOST, _HOST, _HOST]),
    'vc_conv_rulebook_count': (_I, [_P, _I, _I, _I, _HOST, _HOST, _HOST, _HOST, _HOST, _P, _P, _Z, _P]),
    'vc_conv_rulebook_fill': (_I, [_P, _I, _I, _I, _HOST, _HOST, _HOST, _HOST, _HOST, _I, _P, _P, _P, _P, _P, _Z, _P]),
    'vc_pairs_from_nbr': (_I, [_P, _I, _I, _P, _P, _P]),
    'vc_conv_ws_bytes': (_Z, [_I, _I, _I]),
    'vc_conv_fwd_f32': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _Z, _P]),
    'vc_conv_dgrad_f32': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _Z, _P]),
    'vc_conv_dgrad_scatter_f32': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    'vc_conv_wgrad_ws_bytes': (_Z, [_I, _I, _I, _I]),
    'vc_conv_wgrad_f32': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _Z, _P]),
    'vc_cast_f32_bf16': (_I, [_P, _P, ctypes.c_longlong, _P]),
    'vc_conv_tc_ws_bytes': (_Z, [_I, _I, _I]),
    'vc_conv_fwd_tc': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _Z, _P, _P]),
    'vc_conv_dgrad_tc': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _Z, _P, _P]),
    'vc_conv_dgrad_scatter_tc': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _Z, _P, _P]),
    'vc_conv_wgrad_tc_ws_bytes': (_Z, [_I, _I, _I, _I]),
    'vc_conv_wgrad_tc': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _Z, _P, _P]),
    'vc_bn_apply_relu_f32': (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P, _F, _F, _I, _P, _P, _P, _I, _P]),
    'vc_bn_relu_bwd_f32': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P]),
    'vc_index2uv': (_I, [_P, _I, _I, _P, _HOSTF, _I, _I, _I, _P, _P]),
    'vc_dense_f32': (_I, [_P, _P, _I, _I, _I, _I, _HOST, _P, _P]),
    'vc_dense_bwd_f32': (_I, [_P, _P, _I, _I, _I, _I, _HOST, _P, _P]),
    'vc_voxelize_ws_bytes': (_Z, [_I, _I, _I]),
    'vc_voxelize_mean': (_I, [_P, _I, _I, _I, _HOSTF, _HOSTF, _I, _I, _I, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'vc_voxel2pinds': (_I, [_P, _I, _I, _I, _HOST, _P, _P]),
    'vc_cat2_f32': (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    'vc_gather_rows': (_I, [_P, _P, _P, _I, _I, _P]),
    'vc_stvd_ws_bytes': (_Z, [_I]),
    'vc_stvd_partition': (_I, [_P, _I, _I, _I, ctypes.c_double, _P, _P, _Z, _P]),
    'vc_stvd_gather': (_I, [_P, _I, _I, _HOST, _I, _P, _P, _I, _P, _Z, _P]),
    'vc_voxel_query': (_I, [_I, _I, _I, _I, _I, _F, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    'vc_group_points': (_I, [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    'vc_group_points_grad': (_I, [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    'vc_exec_state_bytes': (_Z, []),
    'vc_exec_forward': (_I, [_P, _P, _I, _P, _P, _I, _P, _I, _P, _I, _HOST, _I, _P, _I, _I, _I, _P, _Z, _P, _P, _P, _Z, _P, _P, _I,
                             _P, _P, _P, _P]),
    'vc_exec_backward': (_I, [_P, _P, _I, _P, _P, _I, _P, _P, _I, _P, _Z, _P, _P, _P, _P]),
    'vc_conv_wgrad_tc_config': (_I, [_I, _I]),
    'vc_exec_query': (_I, [_P, _I, _I, _P]),
    'vc_exec_timing': (_I, [_I]),
    'vc_exec_timing_read': (_I, [_P, _P, _I]),
    'vc_allreduce_peer_flag_words': (_I, [_I]),
    'vc_allreduce_peer_f32': (_I, [_P, _P, _I, _I, _P, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_uint, _F, _P, _P]),
}

_lib = None


class VirConvLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VirConvLibraryError(
            f'{LIB_PATH} is missing: build it with `python -m virconv_b200.build` '
            f'(or __graft_entry__.build()).  virconv_b200 has no CPU or PyTorch fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    # programmatic dependent launch of the conv / BN chain: implemented and parity-tested, but measured no faster on
    # B200 (3.18-3.24 ms/step with, 3.24 without; e2e 0.1 ms worse) -> off unless asked for
    lib.vc_set_pdl(1 if os.environ.get('VIRCONV_PDL', '0') == '1' else 0)
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().vc_last_error()
        raise VirConvLibraryError(f'{what} failed (rc={rc}): {msg.decode() if msg else ""}')


def host_i32(vals):
    return (c_int32 * len(vals))(*[int(v) for v in vals])


def host_f32(vals):
    return (c_float * len(vals))(*[float(v) for v in vals])
