"""CPU: the C-ABI library loads, exports every symbol include/vsl.h declares, and validates arguments
before touching a device (no compute calls here -- there is no GPU)."""
import ctypes
import os
import re

import pytest
import torch

import __graft_entry__ as graft
from tf_depth_estimation_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    graft.build()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'vsl.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(vsl_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    syms = header_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), 'libvsl.so does not export %s' % s
        assert s in _lib.SIGNATURES, 'binding lacks a prototype for %s' % s
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.vsl_version() == 100


def test_error_codes_without_a_device(lib):
    assert lib.vsl_strerror(0) == b'ok'
    assert lib.vsl_pose_vec2mat_fwd(None, 4, 0, None, None) == -1          # VSL_E_NULL
    assert lib.vsl_pose_vec2mat_fwd(8, 0, 0, 8, None) == -2                # VSL_E_SHAPE
    assert lib.vsl_pose_vec2mat_fwd(8, 4, 2, 8, None) == -3                # VSL_E_FORMAT (matrix is not a vector)
    assert lib.vsl_warp_fwd(16, 16, 16, 16, 2, 8, 8, 5, 0, None, None, None, None, None, 16, None) == -2  # C = 5
    assert lib.vsl_warp_fwd(16, 16, 16, 16, 2, 8, 8, 3, 7, None, None, None, None, None, 16, None) == -3
    assert lib.vsl_warp_fwd(16, 16, 16, 16, 2, 8, 8, 3, 0, None, None, None, None, None, 20, None) == -4  # ws align
    assert lib.vsl_smooth_fwd(16, 1, 2, 8, 1, 0, 16, 16, None) == -2       # H <= 2: no second difference
    assert b'NULL' in lib.vsl_strerror(-1)
    with pytest.raises(_lib.VslError):
        _lib.check(-2)


def test_loss_descriptor_validation_and_workspace(lib):
    d = _lib.VslLossDesc(32, 128, 416, 4, 2, 0, 1, 1, 1, 0, 1.0, 0.5, 0.2, 1.0)
    n = lib.vsl_loss_ws_bytes(ctypes.byref(d))
    # dominated by the V source views re-laid as zero-bordered RGBA at EVERY level (16 B per pixel, 2 pixels of
    # border) plus the RGB pyramid levels 1..3 of the target
    rgba = 2 * sum(32 * ((128 >> s) + 4) * ((416 >> s) + 4) * 16 for s in range(4))
    tgt = 32 * 128 * 416 * 3 * 4 * (1 / 4 + 1 / 16 + 1 / 64)
    assert rgba + tgt < n < (rgba + tgt) * 1.02
    bad = _lib.VslLossDesc(32, 100, 416, 4, 2, 0, 1, 1, 1, 0, 1.0, 0.5, 0.2, 1.0)   # 100 % 8 != 0
    assert lib.vsl_loss_ws_bytes(ctypes.byref(bad)) == 0
    bad = _lib.VslLossDesc(32, 128, 416, 4, 5, 0, 1, 1, 1, 0, 1.0, 0.5, 0.2, 1.0)   # V > VSL_MAX_VIEWS
    assert lib.vsl_loss_ws_bytes(ctypes.byref(bad)) == 0
    assert ctypes.sizeof(_lib.VslLossDesc) == 22 * 4 + 2 * 8   # 22 scalars (ssim_weight is the last), 2 event handles
    ssim = _lib.VslLossDesc(32, 128, 416, 4, 2, 0, 1, 1, 1, 0, 1.0, 0.5, 0.2, 1.0, exact_coords=1, ssim_weight=0.5)
    assert lib.vsl_loss_ws_bytes(ctypes.byref(ssim)) == 0                         # the SSIM term rides on the fast arithmetic
    ssim = _lib.VslLossDesc(32, 128, 416, 4, 2, 0, 1, 1, 1, 0, 1.0, 0.5, 0.2, 1.0, ssim_weight=1.5)
    assert lib.vsl_loss_ws_bytes(ctypes.byref(ssim)) == 0                         # a weight outside [0, 1]


def test_flow_step_and_loader_entries_without_a_device(lib):
    d = _lib.VslFlowLossDesc(B=64, H=192, W=256, S=4, smooth_weight=0.5, depth_weight=1.0, data_weight=1.0,
                             optflow_weight=1.0, loss_scale=1.0)
    n = lib.vsl_flow_loss_ws_bytes(ctypes.byref(d))
    levels = 64 * 192 * 256 * 7 * 4 * (1 / 4 + 1 / 16 + 1 / 64)       # left + right (RGB) + label pyramids, levels 1..3
    assert levels < n < levels * 1.05
    for bad in (dict(H=100), dict(S=7), dict(B=0), dict(H=16, W=16)):     # 100 % 8, too many scales, empty, 2 x 2 coarsest level
        kw = dict(B=64, H=192, W=256, S=4, smooth_weight=0.5, depth_weight=1.0, data_weight=1.0, optflow_weight=1.0, loss_scale=1.0)
        kw.update(bad)
        assert lib.vsl_flow_loss_ws_bytes(ctypes.byref(_lib.VslFlowLossDesc(**kw))) == 0, bad
    P = _lib.ptr_array([16, 16, 16, 16])
    assert lib.vsl_flow_loss_fwd_bwd(ctypes.byref(d), None, 16, 16, P, P, P, 16, 16, 16, P, P, P, 256, None) == -1   # NULL image
    assert lib.vsl_flow_loss_fwd_bwd(ctypes.byref(d), 16, 16, 16, P, P, P, 16, 16, 16, P, P, P, 260, None) == -4    # ws alignment
    assert lib.vsl_unpack_strip(None, 1, 8, 16, 8, 8, 16, 16, None) == -1
    assert lib.vsl_unpack_strip(16, 1, 8, 1, 8, 8, 16, 16, None) == -2                                             # a strip one pixel wide


def test_host_side_rejects_cpu_tensors_and_bad_shapes(lib):
    img = torch.zeros(2, 8, 8, 3)
    with pytest.raises(TypeError, match='no CPU fallback'):
        ops.projective_inverse_warp(img, torch.ones(2, 8, 8), torch.zeros(2, 6), torch.eye(3).repeat(2, 1, 1))
    with pytest.raises(ValueError):
        ops.projective_inverse_warp(img, torch.ones(2, 8, 9), torch.zeros(2, 6), torch.eye(3).repeat(2, 1, 1))
    with pytest.raises(ValueError):
        ops.projective_inverse_warp(img, torch.ones(2, 8, 8), torch.zeros(2, 6), torch.eye(3).repeat(2, 1, 1), 'quat')
    with pytest.raises(ValueError):
        ops.pose_vec2mat(torch.zeros(2, 7))
    with pytest.raises(ValueError):
        ops.compute_exp_reg_loss(torch.zeros(2, 4, 4, 3))


def test_missing_library_fails_loudly(monkeypatch, lib):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libvsl.so')
    with pytest.raises(ImportError, match='no CPU or eager fallback'):
        _lib.load()


def test_error_codes_of_the_extension_and_optimiser_entries(lib):
    """Argument checks return before anything touches a device (fake non-NULL pointers are never dereferenced)."""
    assert lib.vsl_consist_fwd(None, 16, 16, 2, 8, 8, 8, 8, 16, None) == -1
    assert lib.vsl_consist_fwd(16, 16, 20, 2, 8, 8, 8, 8, 16, None) == -4             # coords not 8-byte aligned
    assert lib.vsl_consist_bwd(16, 16, 16, 0, 8, 8, 8, 8, 16, None, 16, None, None) == -2
    assert lib.vsl_ssim_fwd(16, 16, 2, 2, 8, 3, 16, None, None, None) == -2            # H < 3: no 3x3 window
    assert lib.vsl_ssim_fwd(16, 16, 2, 8, 8, 5, 16, None, None, None) == -2            # C > 4
    assert lib.vsl_ssim_fwd(16, 16, 2, 8, 8, 3, None, None, None, None) == -1          # neither map nor loss
    assert lib.vsl_ssim_fwd(16, 16, 2, 8, 8, 3, None, 16, None, None) == -1            # loss without workspace
    assert lib.vsl_ssim_bwd(16, 16, 2, 8, 8, 3, None, None, 0, 16, None, None) == -1   # no upstream gradient
    assert lib.vsl_ssim_ws_bytes(2, 8, 8, 3) == 4 * 1 * 1 * 2 and lib.vsl_ssim_ws_bytes(2, 2, 8, 3) == 0
    assert lib.vsl_edge_smooth_fwd(16, 16, 2, 1, 8, 3, 16, 16, None) == -2             # H < 2
    assert lib.vsl_edge_smooth_bwd(16, 16, 2, 8, 8, 3, None, None, None, None) == -1
    assert lib.vsl_adam_step(None, 16, 16, 16, 10, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, None) == -1
    assert lib.vsl_adam_step(16, 16, 16, 16, 10, 1e-3, 0.9, 0.999, 1e-8, 0, 1.0, None) == -2     # t >= 1
    assert lib.vsl_adam_step(16, 20, 16, 16, 10, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, None) == -4     # ranges out of phase
    ptrs = _lib.ptr_array([16, 32])
    assert lib.vsl_dp_adam_step(ptrs, ptrs, 2, 2, 16, 16, 0, 8, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, None, None) == -2   # rank
    assert lib.vsl_dp_adam_step(ptrs, ptrs, 0, 2, 16, 16, 0, 6, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, None, None) == -2   # hi % 4
    assert lib.vsl_dp_adam_step(ptrs, ptrs, 0, 2, 16, 16, 8, 8, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, None, None) == 0    # empty shard
    assert lib.vsl_dp_adam_step(_lib.ptr_array([16, 36]), ptrs, 0, 2, 16, 16, 0, 8, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0,
                                None, None) == -4
    assert lib.vsl_peer_barrier(ptrs, 0, 17, 1, None, 0, None) == -2                   # world > 16
    assert lib.vsl_peer_barrier(None, 0, 2, 1, None, 0, None) == -1
    assert lib.vsl_peer_barrier(ptrs, 0, 2, 1, None, -5, None) == -2                   # negative timeout
    # the graph-capturable whole step needs its device-resident state block
    assert lib.vsl_dp_step(ptrs, ptrs, ptrs, 0, 2, 16, 16, 0, 8, 1e-3, 0.9, 0.999, 1e-8, 1.0, None, None, 0, None) == -1
    # the multicast (NVLS) form: NULL addresses / state, a single rank, unaligned addresses, a shard off the 16-byte grid
    assert lib.vsl_dp_step_mc(ptrs, 16, 16, 16, 0, 2, 16, 16, 0, 8, 1e-3, 0.9, 0.999, 1e-8, 1.0, None, None, 0, None) == -1
    assert lib.vsl_dp_step_mc(ptrs, None, 16, 16, 0, 2, 16, 16, 0, 8, 1e-3, 0.9, 0.999, 1e-8, 1.0, 16, None, 0, None) == -1
    assert lib.vsl_dp_step_mc(ptrs, 16, 16, 16, 0, 1, 16, 16, 0, 8, 1e-3, 0.9, 0.999, 1e-8, 1.0, 16, None, 0, None) == -2
    assert lib.vsl_dp_step_mc(ptrs, 16, 16, 16, 0, 2, 16, 16, 2, 8, 1e-3, 0.9, 0.999, 1e-8, 1.0, 16, None, 0, None) == -2
    assert lib.vsl_dp_step_mc(ptrs, 20, 16, 16, 0, 2, 16, 16, 0, 8, 1e-3, 0.9, 0.999, 1e-8, 1.0, 16, None, 0, None) == -4
    assert lib.vsl_scale(None, 16, 4, 16, None, None) == -1 and lib.vsl_scale(16, 16, 0, 16, None, None) == -2
    assert lib.vsl_scale(20, 16, 4, 16, None, None) == -4
    assert lib.vsl_ipc_get_handle(None, None) == -1 and lib.vsl_ipc_open(None, None) == -1
    assert lib.vsl_peer_alloc(0, ctypes.byref(ctypes.c_void_p())) == -2


def test_drop_in_modules_export_the_reference_names():
    """what `from utils import *` / `from utils_lr import *` / `from my_losses import *` must provide
    (names and leading argument names as in the reference; import only, nothing is computed without a GPU)."""
    import importlib
    import inspect
    import sys
    compat = os.path.join(ROOT, 'tf_depth_estimation_b200', 'compat')
    sys.path.insert(0, compat)
    saved = {n: sys.modules.pop(n, None) for n in ('utils', 'utils_lr', 'my_losses')}
    try:
        utils, utils_lr, my_losses = (importlib.import_module(n) for n in ('utils', 'utils_lr', 'my_losses'))
        for m in (utils, utils_lr, my_losses):
            assert os.path.dirname(os.path.abspath(m.__file__)) == compat
        want = {
            utils: {'euler2mat': ['z', 'y', 'x'], 'pose_vec2mat': ['vec'], 'pixel2cam': ['depth', 'pixel_coords', 'intrinsics'],
                    'cam2pixel': ['cam_coords', 'proj'], 'meshgrid': ['batch', 'height', 'width'],
                    'projective_inverse_warp': ['img', 'depth', 'pose', 'intrinsics'],
                    'optflow_warp': ['img', 'flowx', 'flowy'], 'bilinear_sampler': ['imgs', 'coords'],
                    'depth_optflow': ['src_pixel_coords']},
            utils_lr: {'euler2mat': ['z', 'y', 'x'], 'axis_angle_to_rotation_matrix': ['axis', 'angle'],
                       'pose_vec2mat': ['vec', 'format'], 'pixel2cam': ['depth', 'pixel_coords', 'intrinsics'],
                       'cam2pixel': ['cam_coords', 'proj'], 'meshgrid': ['batch', 'height', 'width'],
                       'projective_inverse_warp': ['img', 'depth', 'pose', 'intrinsics', 'format'],
                       'bilinear_sampler': ['imgs', 'coords'],
                       'consistent_depth_loss': ['src_depth', 'pred_src_depth', 'coords']},
            my_losses: {'get_reference_explain_mask': ['downscaling', 'FLAGS'], 'compute_smooth_loss': ['pred_disp'],
                        'compute_exp_reg_loss': ['pred', 'ref'],
                        'compute_loss_single_depth': ['pred_depth', 'label', 'global_step', 'FLAGS'],
                        'compute_loss_pairwise_depth': ['image_left', 'image_right', 'pred_depth_left', 'pred_poses_right',
                                                        'pred_exp_logits_left', 'pred_depth_right', 'pred_poses_left',
                                                        'pred_exp_logits_right', 'gt_right_cam', 'intrinsics', 'label',
                                                        'FLAGS', 'global_step'],
                        'view_synthesis_loss': ['tgt', 'srcs']},
        }
        for mod, names in want.items():
            for name, args in names.items():
                fn = getattr(mod, name)
                got = list(inspect.signature(fn).parameters)[:len(args)]
                assert got == args, (mod.__name__, name, got)
                assert name in mod.__all__, (mod.__name__, name)
    finally:
        sys.path.remove(compat)
        for n, m in saved.items():
            sys.modules.pop(n, None)
            if m is not None:
                sys.modules[n] = m
