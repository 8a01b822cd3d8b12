"""ctypes binding of libvsl.so (the C ABI declared in include/vsl.h).

There is deliberately NO fallback: if the shared library has not been built, or a call is made with
non-CUDA tensors, this raises.  Build with `python -c "import __graft_entry__ as g; g.build()"`.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('VSL_LIB_PATH') or os.path.join(_HERE, 'libvsl.so')  # override: kernel experiments

POSE_FORMATS = {'eular': 0, 'euler': 0, 'angleaxis': 1, 'matrix': 2}
MASK_NONE, MASK_EXP, MASK_CONST = 0, 1, 2
IMG_F32, IMG_U8_255, IMG_U8_255_CENTRED, IMG_U8_RAW = 0, 1, 2, 3
IMG_FORMATS = {'f32': 0, 'u8_255': 1, 'u8_255_centred': 2, 'u8_raw': 3}
MAX_SCALES, MAX_VIEWS = 6, 4

_c_float_p = ctypes.c_void_p
_c_stream = ctypes.c_void_p


class VslLossDesc(ctypes.Structure):
    _fields_ = [('B', ctypes.c_int), ('H', ctypes.c_int), ('W', ctypes.c_int),
                ('S', ctypes.c_int), ('V', ctypes.c_int),
                ('pose_format', ctypes.c_int), ('mask_mode', ctypes.c_int),
                ('pixel_scale_norm', ctypes.c_int), ('depth_is_inverse', ctypes.c_int),
                ('smooth_on_inverse', ctypes.c_int),
                ('data_weight', ctypes.c_float), ('smooth_weight', ctypes.c_float),
                ('explain_reg_weight', ctypes.c_float), ('loss_scale', ctypes.c_float),
                ('exact_coords', ctypes.c_int), ('want_src_grad', ctypes.c_int),
                ('x_is_logit', ctypes.c_int), ('disp_scale', ctypes.c_float), ('disp_min', ctypes.c_float),
                ('img_format', ctypes.c_int), ('consist_weight', ctypes.c_float), ('ssim_weight', ctypes.c_float),
                ('ev_main_begin', ctypes.c_void_p), ('ev_main_end', ctypes.c_void_p)]


class VslFlowLossDesc(ctypes.Structure):
    _fields_ = [('B', ctypes.c_int), ('H', ctypes.c_int), ('W', ctypes.c_int), ('S', ctypes.c_int),
                ('smooth_weight', ctypes.c_float), ('depth_weight', ctypes.c_float), ('data_weight', ctypes.c_float),
                ('optflow_weight', ctypes.c_float), ('loss_scale', ctypes.c_float)]


# name -> (restype, argtypes); every symbol include/vsl.h declares
SIGNATURES = {
    'vsl_version': (ctypes.c_int, []),
    'vsl_strerror': (ctypes.c_char_p, [ctypes.c_int]),
    'vsl_pose_vec2mat_fwd': (ctypes.c_int, [_c_float_p, ctypes.c_int, ctypes.c_int, _c_float_p, _c_stream]),
    'vsl_pose_vec2mat_bwd': (ctypes.c_int, [_c_float_p, _c_float_p, ctypes.c_int, ctypes.c_int, _c_float_p, _c_stream]),
    'vsl_warp_ws_bytes': (ctypes.c_size_t, [ctypes.c_int] * 3),
    'vsl_warp_fwd': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_int] * 5 + [_c_float_p] * 5 + [ctypes.c_void_p, _c_stream]),
    'vsl_warp_bwd': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_int] * 5 + [_c_float_p] * 8 + [ctypes.c_void_p, _c_stream]),
    'vsl_bilinear_fwd': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_int] * 6 + [_c_float_p] * 3 + [_c_stream]),
    'vsl_bilinear_bwd': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_int] * 6 + [_c_float_p] * 4 + [_c_stream]),
    'vsl_consist_fwd': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int] * 5 + [_c_float_p, _c_stream]),
    'vsl_consist_bwd': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int] * 5 + [_c_float_p] * 4 + [_c_stream]),
    'vsl_depth_optflow': (ctypes.c_int, [_c_float_p] + [ctypes.c_int] * 3 + [_c_float_p] * 2 + [_c_stream]),
    'vsl_meshgrid': (ctypes.c_int, [ctypes.c_int] * 4 + [_c_float_p, _c_stream]),
    'vsl_pixel2cam_fwd': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int] * 4 + [_c_float_p, _c_stream]),
    'vsl_pixel2cam_bwd': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int] * 4 + [_c_float_p, _c_stream]),
    'vsl_cam2pixel_fwd': (ctypes.c_int, [_c_float_p] * 2 + [ctypes.c_int] * 3 + [_c_float_p] * 2 + [_c_stream]),
    'vsl_cam2pixel_bwd': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_int] * 3 + [_c_float_p] * 2 + [_c_stream]),
    'vsl_axis_angle_fwd': (ctypes.c_int, [_c_float_p] * 2 + [ctypes.c_int, _c_float_p, _c_stream]),
    'vsl_axis_angle_bwd': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int] + [_c_float_p] * 2 + [_c_stream]),
    'vsl_smooth_ws_bytes': (ctypes.c_size_t, [ctypes.c_int] * 4),
    'vsl_smooth_fwd': (ctypes.c_int, [_c_float_p] + [ctypes.c_int] * 5 + [_c_float_p, ctypes.c_void_p, _c_stream]),
    'vsl_smooth_bwd': (ctypes.c_int, [_c_float_p] + [ctypes.c_int] * 5 + [_c_float_p, _c_float_p, _c_stream]),
    'vsl_expreg_ws_bytes': (ctypes.c_size_t, [ctypes.c_longlong]),
    'vsl_expreg_fwd': (ctypes.c_int, [_c_float_p, ctypes.c_longlong, _c_float_p, ctypes.c_void_p, _c_stream]),
    'vsl_expreg_bwd': (ctypes.c_int, [_c_float_p, ctypes.c_longlong, _c_float_p, _c_float_p, _c_stream]),
    'vsl_pyramid': (ctypes.c_int, [_c_float_p] + [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_void_p), _c_stream]),
    'vsl_loss_ws_bytes': (ctypes.c_size_t, [ctypes.POINTER(VslLossDesc)]),
    'vsl_loss_ws_layout': (ctypes.c_int, [ctypes.POINTER(VslLossDesc), ctypes.POINTER(ctypes.c_longlong)]),
    'vsl_loss_fwd_bwd': (ctypes.c_int, [ctypes.POINTER(VslLossDesc), _c_float_p, ctypes.POINTER(ctypes.c_void_p),
                                        ctypes.POINTER(ctypes.c_void_p), _c_float_p, _c_float_p,
                                        ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                        _c_float_p, ctypes.POINTER(ctypes.c_void_p), _c_float_p,
                                        ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                        ctypes.c_void_p, _c_stream]),
    'vsl_loss_consist_fwd_bwd': (ctypes.c_int, [ctypes.POINTER(VslLossDesc), _c_float_p, ctypes.POINTER(ctypes.c_void_p),
                                                ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                                _c_float_p, _c_float_p, ctypes.POINTER(ctypes.c_void_p),
                                                ctypes.POINTER(ctypes.c_void_p), _c_float_p,
                                                ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                                _c_float_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, _c_stream]),
    'vsl_loss_fwd_bwd_u8': (ctypes.c_int, [ctypes.POINTER(VslLossDesc), ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                           ctypes.POINTER(ctypes.c_void_p), _c_float_p, _c_float_p,
                                           ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                           _c_float_p, ctypes.POINTER(ctypes.c_void_p), _c_float_p,
                                           ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, _c_stream]),
    'vsl_flow_loss_ws_bytes': (ctypes.c_size_t, [ctypes.POINTER(VslFlowLossDesc)]),
    'vsl_flow_loss_fwd_bwd': (ctypes.c_int, [ctypes.POINTER(VslFlowLossDesc)] + [_c_float_p] * 3 +
                              [ctypes.POINTER(ctypes.c_void_p)] * 3 + [_c_float_p] * 3 +
                              [ctypes.POINTER(ctypes.c_void_p)] * 3 + [ctypes.c_void_p, _c_stream]),
    'vsl_unpack_strip': (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_int] * 5 + [_c_float_p] * 2 + [_c_stream]),
    'vsl_scale': (ctypes.c_int, [_c_float_p] * 2 + [ctypes.c_longlong] + [_c_float_p] * 2 + [_c_stream]),
    'vsl_adam_step': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_longlong] + [ctypes.c_float] * 4 + [ctypes.c_int, ctypes.c_float, _c_stream]),
    'vsl_peer_alloc': (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    'vsl_peer_free': (ctypes.c_int, [ctypes.c_void_p]),
    'vsl_ipc_get_handle': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    'vsl_ipc_open': (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    'vsl_ipc_close': (ctypes.c_int, [ctypes.c_void_p]),
    'vsl_peer_barrier': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_uint,
                                        ctypes.c_void_p, ctypes.c_longlong, _c_stream]),
    'vsl_dp_adam_step': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p)] * 2 + [ctypes.c_int] * 2 + [_c_float_p] * 2 +
                         [ctypes.c_longlong] * 2 + [ctypes.c_float] * 4 + [ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                                                                          _c_stream]),
    'vsl_dp_step': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p)] * 3 + [ctypes.c_int] * 2 + [_c_float_p] * 2 +
                    [ctypes.c_longlong] * 2 + [ctypes.c_float] * 5 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                                                     _c_stream]),
    'vsl_dp_step_mc': (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p)] + [ctypes.c_void_p] * 3 + [ctypes.c_int] * 2 +
                       [_c_float_p] * 2 + [ctypes.c_longlong] * 2 + [ctypes.c_float] * 5 +
                       [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, _c_stream]),
    'vsl_ssim_ws_bytes': (ctypes.c_size_t, [ctypes.c_int] * 4),
    'vsl_ssim_fwd': (ctypes.c_int, [_c_float_p] * 2 + [ctypes.c_int] * 4 + [_c_float_p] * 2 + [ctypes.c_void_p, _c_stream]),
    'vsl_ssim_bwd': (ctypes.c_int, [_c_float_p] * 2 + [ctypes.c_int] * 4 + [_c_float_p] * 2 + [ctypes.c_int] +
                     [_c_float_p] * 2 + [_c_stream]),
    'vsl_edge_smooth_ws_bytes': (ctypes.c_size_t, [ctypes.c_int] * 3),
    'vsl_edge_smooth_fwd': (ctypes.c_int, [_c_float_p] * 2 + [ctypes.c_int] * 4 + [_c_float_p, ctypes.c_void_p, _c_stream]),
    'vsl_edge_smooth_bwd': (ctypes.c_int, [_c_float_p] * 2 + [ctypes.c_int] * 4 + [_c_float_p] * 3 + [_c_stream]),
}

_lib = None


class VslError(RuntimeError):
    pass


def load():
    """dlopen libvsl.so once and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            'libvsl.so is not built (%s). Run: python -c "import __graft_entry__ as g; g.build()". '
            'There is no CPU or eager fallback for this path.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drifted apart
        fn.restype, fn.argtypes = res, args
    if lib.vsl_version() != 100:
        raise ImportError('libvsl.so version %d does not match the binding (100)' % lib.vsl_version())
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise VslError('%s (code %d)' % (load().vsl_strerror(rc).decode(), rc))


def ptr_array(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*ptrs)
