"""ctypes binding of libstyle_b200.so (the C ABI in include/style_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception
is raised.  The library is built in-tree by ``style_transfer_b200/csrc/build.sh`` (or
``__graft_entry__.build()``).
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libstyle_b200.so')

ST_PREC_FP32, ST_PREC_BF16, ST_PREC_FP16, ST_PREC_TC32 = 0, 1, 2, 3
ST_COMM_ID_BYTES = 128
ST_CONV3X3, ST_POOL_MAX, ST_POOL_AVE = 0, 1, 2


class StError(RuntimeError):
    """A libstyle_b200 call returned a negative status."""


class LayerDesc(C.Structure):
    _fields_ = [('kind', C.c_int32), ('bottom', C.c_int32), ('cin', C.c_int32),
                ('cout', C.c_int32)]


class LossSpec(C.Structure):
    _fields_ = [('blob', C.c_int32), ('use_content', C.c_int32), ('use_style', C.c_int32),
                ('use_dd', C.c_int32), ('content_weight', C.c_float), ('style_weight', C.c_float),
                ('dd_weight', C.c_float)]


_vp, _i, _f, _sz, _l = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_long
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); every name here must be declared in include/style_b200.h
PROTOTYPES = {
    'st_last_error': (C.c_char_p, []),
    'st_version': (_i, []),
    'st_launch_count': (C.c_uint64, []),
    'st_create': (_i, [_i, _i, _i, C.POINTER(LayerDesc), C.POINTER(_vp)]),
    'st_destroy': (_i, [_vp]),
    'st_set_conv_params': (_i, [_vp, _i, _vp, _vp]),
    'st_reserve': (_i, [_vp, _i, _i]),
    'st_device_info': (_i, [_vp, _ip, C.POINTER(_sz)]),
    'st_clear_targets': (_i, [_vp]),
    'st_set_style_gram': (_i, [_vp, _i, _i, _vp, _vp]),
    'st_set_content_features': (_i, [_vp, _i, _i, _vp, _i, _i, _vp]),
    'st_eval_features_tile': (_i, [_vp, _vp, _i, _i, _i, C.POINTER(C.c_int32), C.POINTER(_vp),
                                   _vp]),
    'st_eval_sc_grad_tile': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, C.POINTER(LossSpec), _vp,
                                  _vp, _l, _l, _vp]),
    'st_eval_sc_grad_tiles': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(LossSpec),
                                   _vp, _vp, _vp]),
    'st_eval_sc_grad_tile_range': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                                        C.POINTER(LossSpec), _vp, _vp, _vp]),
    'st_tile_grid': (_i, [_i, _i, _i, _ip, _ip, _ip, _ip]),
    'st_unpack_grad': (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'st_packed_floats': (_sz, [_i, _i, _i, _i]),
    'st_comm_unique_id': (_i, [_vp]),
    'st_comm_init': (_i, [_vp, _vp, _i, _i]),
    'st_comm_destroy': (_i, [_vp]),
    'st_allgather_grad': (_i, [_vp, _vp, _vp, _sz, _vp]),
    'st_gram': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'st_regularizers': (_i, [_vp, _i, _i, C.POINTER(_f), _f, _f, _f, _f, _vp, _f, _i, _i, _vp,
                             _vp, _vp]),
    'st_unpack_regularize': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, C.POINTER(_f), _f, _f, _f, _f,
                                  _vp, _f, _vp, _vp, _vp]),
    'st_adam_step': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _f, _f, _f, _vp]),
    'st_lbfgs_inv_hv': (_i, [_vp, _sz, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_double),
                             _vp, _vp, _vp]),
    'st_lbfgs_step': (_i, [_vp, _sz, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp]),
    'st_lbfgs_commit': (_i, [_vp, _vp, _sz, _i, _vp, _vp, _vp, _vp]),
    'st_resize_f32': (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'st_resample_coeffs': (_i, [_i, _i, _i, _ip, _vp, _vp]),
    'st_iter_stats': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'st_get_image_u8': (_i, [_vp, _i, _i, C.POINTER(_f), _i, _vp, _vp]),
    'st_output_step': (_i, [_vp, _vp, _i, _i, C.POINTER(_f), _i, _vp, _vp, _vp]),
    'st_dot': (_i, [_vp, _vp, _sz, _vp, _vp]),
    'st_asum': (_i, [_vp, _sz, _vp, _vp]),
    'st_axpby': (_i, [_f, _vp, _f, _vp, _sz, _vp]),
    'st_timing_enable': (_i, [_i]),
    'st_timing_read': (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double),
                            C.POINTER(C.c_uint64)]),
    'st_timing_reset': (_i, []),
}

_lib = None


def load():
    """Loads the shared library (once) and installs the prototypes."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise StError('libstyle_b200.so is not built: run style_transfer_b200/csrc/build.sh '
                          '(there is no CPU fallback)')
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        msg = load().st_last_error()
        raise StError('libstyle_b200 error %d: %s' % (status, msg.decode() if msg else '?'))


def call(name, *args):
    check(getattr(load(), name)(*args))
