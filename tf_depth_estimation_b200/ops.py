"""Host side of the view-synthesis-loss path: torch tensors in, libvsl (C ABI) calls out.

torch is used for device memory, streams and autograd bookkeeping only; every number is produced by the
hand-written sm_100a kernels in csrc/.  All tensors must be CUDA float32; anything else raises.
"""
import collections
import ctypes
import threading
import warnings

import torch

from . import _lib
from ._lib import POSE_FORMATS, VslLossDesc, check, ptr_array


def _stream():
    return torch.cuda.current_stream().cuda_stream


class VslLayoutWarning(UserWarning):
    """A non-contiguous input was copied into a packed buffer before the C call."""


_LAYOUT_POLICY = ['copy']


def set_layout_policy(policy):
    """What to do with a non-contiguous input (the C ABI takes packed buffers, as TF hands its kernels):
    'copy'   -- pack it into a fresh buffer and emit a VslLayoutWarning naming the argument (default: the reference's
                call sites pass channel slices such as src_image_stack[:, :, :, 3*i:3*(i+1)], train.py:127, which
                TF materialises too);
    'strict' -- raise ValueError (for callers that want to be sure no hidden copy sits on their hot path).
    Returns the previous policy."""
    if policy not in ('copy', 'strict'):
        raise ValueError("policy must be 'copy' or 'strict'")
    prev, _LAYOUT_POLICY[0] = _LAYOUT_POLICY[0], policy
    return prev


def from_external(t, name='tensor'):
    """Ingress for device memory owned by another framework: anything that speaks DLPack (`__dlpack__`, or a raw
    DLPack capsule such as tf.experimental.dlpack.to_dlpack(x) returns) or `__cuda_array_interface__` (CuPy, Numba)
    becomes a zero-copy torch view.  torch here is the memory / stream plumbing under the C ABI, not a compute path."""
    if isinstance(t, torch.Tensor):
        return t
    if type(t).__name__ == 'PyCapsule':
        return torch.utils.dlpack.from_dlpack(t)
    if hasattr(t, '__dlpack__'):
        return torch.from_dlpack(t)
    if hasattr(t, '__cuda_array_interface__'):
        return torch.as_tensor(t, device='cuda')
    raise TypeError('%s must be a CUDA tensor, a DLPack producer or a __cuda_array_interface__ object, got %s'
                    % (name, type(t).__name__))


def _is_foreign(t):
    return (not isinstance(t, torch.Tensor)) and (type(t).__name__ == 'PyCapsule' or hasattr(t, '__dlpack__') or
                                                    hasattr(t, '__cuda_array_interface__'))


def _ingress(fn):
    """Public entry points accept foreign device tensors for every tensor argument (see from_external)."""
    import functools

    scalars = (torch.Tensor, int, float, bool, str, type(None))

    def conv(a):
        # anything that is neither a tensor nor a plain scalar / flag must be a device-memory producer: lists,
        # numpy arrays etc. raise TypeError in from_external instead of an AttributeError deep inside the op
        return a if isinstance(a, scalars) else from_external(a)

    @functools.wraps(fn)
    def wrapped(*args, **kw):
        return fn(*[conv(a) for a in args], **{k: conv(v) for k, v in kw.items()})
    return wrapped


def to_dlpack(t):
    """Egress: a DLPack capsule of a result (zero copy), e.g. for tf.experimental.dlpack.from_dlpack."""
    return torch.utils.dlpack.to_dlpack(t)


def _f32(t, name):
    t = from_external(t, name)
    if not t.is_cuda:
        raise TypeError('%s must be a CUDA tensor (this path has no CPU fallback)' % name)
    if t.dtype != torch.float32:
        raise TypeError('%s must be float32, got %s' % (name, t.dtype))
    if not t.is_contiguous():
        if _LAYOUT_POLICY[0] == 'strict':
            raise ValueError('%s is not contiguous (shape %s, strides %s) and the layout policy is strict'
                             % (name, tuple(t.shape), tuple(t.stride())))
        warnings.warn('%s is not contiguous (shape %s, strides %s): packed into a copy before the C call'
                      % (name, tuple(t.shape), tuple(t.stride())), VslLayoutWarning, stacklevel=3)
        t = t.contiguous()
    return t


def _img(t, name, dtype):
    """An image argument of the fused step: float32, or the loader's uint8 when the flags say so."""
    if dtype == torch.float32:
        return _f32(t, name)
    t = from_external(t, name)
    if not t.is_cuda:
        raise TypeError('%s must be a CUDA tensor (this path has no CPU fallback)' % name)
    if t.dtype != torch.uint8:
        raise TypeError('%s must be uint8 for this img_format, got %s' % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def _g32(t, name='grad'):
    """An upstream gradient handed over by autograd (never a caller's input): packed without a warning."""
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError('%s must be CUDA float32' % name)
    return t.contiguous()


def _p(t):
    return None if t is None else t.data_ptr()


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _fmt(format):
    try:
        return POSE_FORMATS[format]
    except KeyError:
        raise ValueError("format must be 'eular', 'angleaxis' or 'matrix', got %r" % (format,))


# ----------------------------------------------------------------------------------------------------
class _PoseVec2Mat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vec, fmt):
        lib = _lib.load()
        vec = _f32(vec, 'vec')
        B = vec.shape[0]
        mat = torch.empty(B, 4, 4, device=vec.device, dtype=torch.float32)
        check(lib.vsl_pose_vec2mat_fwd(vec.data_ptr(), B, fmt, mat.data_ptr(), _stream()))
        ctx.save_for_backward(vec)
        ctx.fmt = fmt
        return mat

    @staticmethod
    def backward(ctx, g_mat):
        vec, = ctx.saved_tensors
        g_vec = torch.empty_like(vec)
        check(_lib.load().vsl_pose_vec2mat_bwd(vec.data_ptr(), _g32(g_mat, 'g_mat').data_ptr(), vec.shape[0],
                                               ctx.fmt, g_vec.data_ptr(), _stream()))
        return g_vec, None


@_ingress
def pose_vec2mat(vec, format='eular'):
    if vec.dim() != 2 or vec.shape[1] != 6:
        raise ValueError('vec must be [B, 6], got %s' % (tuple(vec.shape),))
    fmt = _fmt(format)
    if fmt == 2:
        raise ValueError("pose_vec2mat takes 'eular' or 'angleaxis'")
    return _PoseVec2Mat.apply(vec, fmt)


# ----------------------------------------------------------------------------------------------------
class _ProjectiveInverseWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, depth, pose, K, fmt):
        lib = _lib.load()
        ctx.set_materialize_grads(False)   # unused outputs arrive as None, not as zero-filled maps the kernel would read
        img, depth, pose, K = _f32(img, 'img'), _f32(depth, 'depth'), _f32(pose, 'pose'), _f32(K, 'intrinsics')
        B, H, W, C = img.shape
        dev = img.device
        out = torch.empty(B, H, W, C, device=dev)
        coords = torch.empty(B, H, W, 2, device=dev)
        wmask = torch.empty(B, H, W, 1, device=dev)
        z = torch.empty(B, H, W, 1, device=dev)
        pose_mat = torch.empty(B, 4, 4, device=dev)
        ws = _ws(lib.vsl_warp_ws_bytes(B, H, W), dev)
        check(lib.vsl_warp_fwd(img.data_ptr(), depth.data_ptr(), pose.data_ptr(), K.data_ptr(), B, H, W, C, fmt,
                               out.data_ptr(), coords.data_ptr(), wmask.data_ptr(), z.data_ptr(),
                               pose_mat.data_ptr(), ws.data_ptr(), _stream()))
        ctx.save_for_backward(img, depth, pose, K)
        ctx.fmt = fmt
        return out, coords, wmask, z, pose_mat

    @staticmethod
    def backward(ctx, g_out, g_coords, g_wmask, g_z, g_pose_mat):
        lib = _lib.load()
        img, depth, pose, K = ctx.saved_tensors
        B, H, W, C = img.shape
        need_img, need_depth, need_pose = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        gs = [None if g is None else _g32(g, 'grad') for g in (g_out, g_coords, g_wmask, g_z, g_pose_mat)]
        g_img = torch.empty_like(img) if need_img else None
        g_depth = torch.empty_like(depth) if need_depth else None
        g_pose = torch.empty_like(pose) if need_pose else None
        ws = _ws(lib.vsl_warp_ws_bytes(B, H, W), img.device)
        check(lib.vsl_warp_bwd(img.data_ptr(), depth.data_ptr(), pose.data_ptr(), K.data_ptr(), B, H, W, C, ctx.fmt,
                               _p(gs[0]), _p(gs[1]), _p(gs[2]), _p(gs[3]), _p(gs[4]),
                               _p(g_img), _p(g_depth), _p(g_pose), ws.data_ptr(), _stream()))
        return g_img, g_depth, g_pose, None, None


@_ingress
def projective_inverse_warp(img, depth, pose, intrinsics, format='eular'):
    """utils_lr.py:222-256 -> (out_img, src_pixel_coords, wmask, src_depth, pose_mat)."""
    fmt = _fmt(format)
    if img.dim() != 4 or depth.dim() != 3 or tuple(depth.shape) != tuple(img.shape[:3]):
        raise ValueError('img [B,H,W,C] and depth [B,H,W] disagree: %s vs %s' % (tuple(img.shape), tuple(depth.shape)))
    B = img.shape[0]
    want = (B, 4, 4) if fmt == 2 else (B, 6)
    if tuple(pose.shape) != want or tuple(intrinsics.shape) != (B, 3, 3):
        raise ValueError('pose must be %s and intrinsics %s' % (want, (B, 3, 3)))
    return _ProjectiveInverseWarp.apply(img, depth, pose, intrinsics, fmt)


# ----------------------------------------------------------------------------------------------------
class _Bilinear(torch.autograd.Function):
    """coords given (flow is None) or meshgrid + flow."""

    @staticmethod
    def forward(ctx, imgs, coords, flowx, flowy):
        lib = _lib.load()
        ctx.set_materialize_grads(False)
        imgs = _f32(imgs, 'imgs')
        B, Hs, Ws, C = imgs.shape
        if coords is not None:
            coords = _f32(coords, 'coords')
            Ht, Wt = coords.shape[1], coords.shape[2]
        else:
            flowx, flowy = _f32(flowx, 'flowx'), _f32(flowy, 'flowy')
            Ht, Wt = flowx.shape[1], flowx.shape[2]
        out = torch.empty(B, Ht, Wt, C, device=imgs.device)
        wmask = torch.empty(B, Ht, Wt, 1, device=imgs.device)
        check(lib.vsl_bilinear_fwd(imgs.data_ptr(), _p(coords), _p(flowx), _p(flowy), B, Hs, Ws, C, Ht, Wt,
                                   out.data_ptr(), wmask.data_ptr(), None, _stream()))
        ctx.save_for_backward(imgs, coords, flowx, flowy)
        return out, wmask

    @staticmethod
    def backward(ctx, g_out, g_wmask):
        lib = _lib.load()
        imgs, coords, flowx, flowy = ctx.saved_tensors
        B, Hs, Ws, C = imgs.shape
        ref = coords if coords is not None else flowx
        Ht, Wt = ref.shape[1], ref.shape[2]
        g_out = None if g_out is None else _g32(g_out, 'g_out')
        g_wmask = None if g_wmask is None else _g32(g_wmask, 'g_wmask')
        need_c = any(ctx.needs_input_grad[1:])
        g_imgs = torch.empty_like(imgs) if ctx.needs_input_grad[0] else None
        g_coords = torch.empty(B, Ht, Wt, 2, device=imgs.device) if need_c else None
        check(lib.vsl_bilinear_bwd(imgs.data_ptr(), _p(coords), _p(flowx), _p(flowy), B, Hs, Ws, C, Ht, Wt,
                                   _p(g_out), _p(g_wmask), _p(g_imgs), _p(g_coords), _stream()))
        if coords is not None:
            return g_imgs, g_coords, None, None
        gfx = g_coords[..., 0:1].contiguous() if ctx.needs_input_grad[2] else None
        gfy = g_coords[..., 1:2].contiguous() if ctx.needs_input_grad[3] else None
        return g_imgs, None, gfx, gfy


@_ingress
def bilinear_sampler(imgs, coords):
    """utils.py:219-308 -> (output, wmask)."""
    if imgs.dim() != 4 or coords.dim() != 4 or coords.shape[3] != 2 or coords.shape[0] != imgs.shape[0]:
        raise ValueError('imgs [B,Hs,Ws,C] / coords [B,Ht,Wt,2] expected')
    return _Bilinear.apply(imgs, coords, None, None)


@_ingress
def optflow_warp(img, flowx, flowy):
    """utils.py:201-217 -> output_img."""
    B, H, W, _ = img.shape
    if tuple(flowx.shape) != (B, H, W, 1) or tuple(flowy.shape) != (B, H, W, 1):
        raise ValueError('flowx / flowy must be [B,H,W,1]')
    return _Bilinear.apply(img, None, flowx, flowy)[0]


class _DepthOptflow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords):
        coords = _f32(coords, 'coords')
        B, H, W, _ = coords.shape
        fx = torch.empty(B, H, W, 1, device=coords.device)
        fy = torch.empty(B, H, W, 1, device=coords.device)
        check(_lib.load().vsl_depth_optflow(coords.data_ptr(), B, H, W, fx.data_ptr(), fy.data_ptr(), _stream()))
        return fx, fy

    @staticmethod
    def backward(ctx, gfx, gfy):
        return torch.cat([gfx, gfy], dim=3)


@_ingress
def depth_optflow(src_pixel_coords):
    """utils.py:321-338 -> (flowx, flowy)."""
    return _DepthOptflow.apply(src_pixel_coords)


class _ConsistentDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src_depth, pred, coords):
        src_depth, pred, coords = _f32(src_depth, 'src_depth'), _f32(pred, 'pred_src_depth'), _f32(coords, 'coords')
        B, Hs, Ws, _ = src_depth.shape
        Ht, Wt = coords.shape[1], coords.shape[2]
        err = torch.empty(B, Ht, Wt, 1, device=src_depth.device)
        check(_lib.load().vsl_consist_fwd(src_depth.data_ptr(), pred.data_ptr(), coords.data_ptr(), B, Hs, Ws, Ht, Wt,
                                          err.data_ptr(), _stream()))
        ctx.save_for_backward(src_depth, pred, coords)
        return err

    @staticmethod
    def backward(ctx, g_err):
        src_depth, pred, coords = ctx.saved_tensors
        B, Hs, Ws, _ = src_depth.shape
        Ht, Wt = coords.shape[1], coords.shape[2]
        g_s = torch.empty_like(src_depth) if ctx.needs_input_grad[0] else None
        g_p = torch.empty_like(pred) if ctx.needs_input_grad[1] else None
        g_c = torch.empty_like(coords) if ctx.needs_input_grad[2] else None
        check(_lib.load().vsl_consist_bwd(src_depth.data_ptr(), pred.data_ptr(), coords.data_ptr(), B, Hs, Ws, Ht, Wt,
                                          _g32(g_err, 'g_err').data_ptr(), _p(g_s), _p(g_p), _p(g_c), _stream()))
        return g_s, g_p, g_c


@_ingress
def consistent_depth_loss(src_depth, pred_src_depth, coords):
    """utils_lr.py:369-458: |pred_src_depth - bilinear(src_depth, coords)| (no reduction), one kernel each way."""
    if src_depth.dim() != 4 or src_depth.shape[3] != 1 or coords.dim() != 4 or coords.shape[3] != 2:
        raise ValueError('src_depth [B,Hs,Ws,1] / coords [B,Ht,Wt,2] expected')
    if tuple(pred_src_depth.shape) != (coords.shape[0], coords.shape[1], coords.shape[2], 1):
        raise ValueError('pred_src_depth must be [B,Ht,Wt,1]')
    return _ConsistentDepth.apply(src_depth, pred_src_depth, coords)


# ----------------------------------------------------------------------------------------------------
def meshgrid(batch, height, width, is_homogeneous=True, device='cuda'):
    """utils.py:142-166 -> [B, 3|2, H, W] (the reference's float32 linspace grid, not exact integers)."""
    out = torch.empty(batch, 3 if is_homogeneous else 2, height, width, device=device)
    check(_lib.load().vsl_meshgrid(batch, height, width, int(is_homogeneous), out.data_ptr(), _stream()))
    return out


class _Pixel2Cam(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, pc, K, homog):
        depth, pc, K = _f32(depth, 'depth'), _f32(pc, 'pixel_coords'), _f32(K, 'intrinsics')
        B, H, W = depth.shape
        cam = torch.empty(B, 4 if homog else 3, H, W, device=depth.device)
        check(_lib.load().vsl_pixel2cam_fwd(depth.data_ptr(), pc.data_ptr(), K.data_ptr(), B, H, W, int(homog),
                                            cam.data_ptr(), _stream()))
        ctx.save_for_backward(pc, K)
        ctx.homog = homog
        return cam

    @staticmethod
    def backward(ctx, g_cam):
        pc, K = ctx.saved_tensors
        B, _, H, W = pc.shape
        g_depth = torch.empty(B, H, W, device=pc.device)
        check(_lib.load().vsl_pixel2cam_bwd(pc.data_ptr(), K.data_ptr(), _g32(g_cam, 'g_cam').data_ptr(), B, H, W,
                                            int(ctx.homog), g_depth.data_ptr(), _stream()))
        return g_depth, None, None, None


@_ingress
def pixel2cam(depth, pixel_coords, intrinsics, is_homogeneous=True):
    """utils.py:100-119 -> [B, 4|3, H, W].  Differentiable wrt depth."""
    if depth.dim() != 3 or pixel_coords.dim() != 4 or pixel_coords.shape[1] != 3:
        raise ValueError('depth [B,H,W] and pixel_coords [B,3,H,W] expected')
    return _Pixel2Cam.apply(depth, pixel_coords, intrinsics, bool(is_homogeneous))


class _Cam2Pixel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cam, proj):
        ctx.set_materialize_grads(False)
        cam, proj = _f32(cam, 'cam_coords'), _f32(proj, 'proj')
        B, _, H, W = cam.shape
        coords = torch.empty(B, H, W, 2, device=cam.device)
        z = torch.empty(B, H, W, 1, device=cam.device)
        check(_lib.load().vsl_cam2pixel_fwd(cam.data_ptr(), proj.data_ptr(), B, H, W, coords.data_ptr(), z.data_ptr(),
                                            _stream()))
        ctx.save_for_backward(cam, proj)
        return coords, z

    @staticmethod
    def backward(ctx, g_coords, g_z):
        cam, proj = ctx.saved_tensors
        B, _, H, W = cam.shape
        g_cam = torch.empty_like(cam) if ctx.needs_input_grad[0] else None
        g_proj = torch.empty_like(proj) if ctx.needs_input_grad[1] else None
        gc = None if g_coords is None else _g32(g_coords, 'g_coords')
        gz = None if g_z is None else _g32(g_z, 'g_z')
        check(_lib.load().vsl_cam2pixel_bwd(cam.data_ptr(), proj.data_ptr(), _p(gc), _p(gz), B, H, W, _p(g_cam),
                                            _p(g_proj), _stream()))
        return g_cam, g_proj


@_ingress
def cam2pixel(cam_coords, proj):
    """utils_lr.py:172-194 -> (pixel_coords [B,H,W,2], z_u [B,H,W,1])."""
    if cam_coords.dim() != 4 or cam_coords.shape[1] != 4 or tuple(proj.shape) != (cam_coords.shape[0], 4, 4):
        raise ValueError('cam_coords [B,4,H,W] and proj [B,4,4] expected')
    return _Cam2Pixel.apply(cam_coords, proj)


class _AxisAngle(torch.autograd.Function):
    @staticmethod
    def forward(ctx, axis, angle):
        axis, angle = _f32(axis, 'axis'), _f32(angle, 'angle')
        B = axis.shape[0]
        R = torch.empty(B, 3, 3, device=axis.device)
        check(_lib.load().vsl_axis_angle_fwd(axis.data_ptr(), angle.data_ptr(), B, R.data_ptr(), _stream()))
        ctx.save_for_backward(axis, angle)
        return R

    @staticmethod
    def backward(ctx, g_R):
        axis, angle = ctx.saved_tensors
        g_axis, g_angle = torch.empty_like(axis), torch.empty_like(angle)
        check(_lib.load().vsl_axis_angle_bwd(axis.data_ptr(), angle.data_ptr(), _g32(g_R, 'g_R').data_ptr(),
                                             axis.shape[0], g_axis.data_ptr(), g_angle.data_ptr(), _stream()))
        return g_axis, g_angle


@_ingress
def axis_angle_to_rotation_matrix(axis, angle):
    """utils_lr.py:77-103: axis [B,3], angle [B,1,1] -> I + sin(angle) [axis]x + (1-cos(angle)) [axis]x^2."""
    if axis.dim() != 2 or axis.shape[1] != 3 or angle.numel() != axis.shape[0]:
        raise ValueError('axis [B,3] and angle [B,1,1] expected')
    return _AxisAngle.apply(axis, angle.reshape(-1)).reshape(-1, 3, 3)


@_ingress
def euler2mat(z, y, x):
    """utils.py:26-75: z, y, x [B,1] -> R = Rx.Ry.Rz [B,1,3,3] (angles clipped to +-pi)."""
    vec = torch.cat([torch.zeros(z.shape[0], 3, device=z.device), x.reshape(-1, 1), y.reshape(-1, 1), z.reshape(-1, 1)], 1)
    return pose_vec2mat(vec, 'eular')[:, :3, :3].unsqueeze(1)


# ----------------------------------------------------------------------------------------------------
class _SmoothLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inverse):
        lib = _lib.load()
        x = _f32(x, 'pred_disp')
        B, H, W, C = x.shape
        loss = torch.empty((), device=x.device)
        ws = _ws(lib.vsl_smooth_ws_bytes(B, H, W, C), x.device)
        check(lib.vsl_smooth_fwd(x.data_ptr(), B, H, W, C, int(inverse), loss.data_ptr(), ws.data_ptr(), _stream()))
        ctx.save_for_backward(x)
        ctx.inverse = int(inverse)
        return loss

    @staticmethod
    def backward(ctx, g):
        x, = ctx.saved_tensors
        B, H, W, C = x.shape
        g_x = torch.empty_like(x)
        check(_lib.load().vsl_smooth_bwd(x.data_ptr(), B, H, W, C, ctx.inverse, _g32(g, 'g').data_ptr(),
                                         g_x.data_ptr(), _stream()))
        return g_x, None


@_ingress
def compute_smooth_loss(pred_disp, inverse=False):
    """my_losses.py:27-36.  inverse=True evaluates the loss on 1/pred_disp inside the kernel."""
    if pred_disp.dim() != 4:
        raise ValueError('pred_disp must be [B,H,W,C]')
    return _SmoothLoss.apply(pred_disp, inverse)


class _ExpReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits):
        lib = _lib.load()
        logits = _f32(logits, 'pred')
        N = logits.numel() // 2
        loss = torch.empty((), device=logits.device)
        ws = _ws(lib.vsl_expreg_ws_bytes(N), logits.device)
        check(lib.vsl_expreg_fwd(logits.data_ptr(), N, loss.data_ptr(), ws.data_ptr(), _stream()))
        ctx.save_for_backward(logits)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits, = ctx.saved_tensors
        g_l = torch.empty_like(logits)
        check(_lib.load().vsl_expreg_bwd(logits.data_ptr(), logits.numel() // 2, _g32(g, 'g').data_ptr(),
                                         g_l.data_ptr(), _stream()))
        return g_l


@_ingress
def compute_exp_reg_loss(pred, ref=None):
    """my_losses.py:39-43 with ref = the constant [0,1] mask of my_losses.py:14-23 (the only one used)."""
    if pred.shape[-1] != 2:
        raise ValueError('pred must have 2 channels')
    return _ExpReg.apply(pred)


# ----------------------------------------------------------------------------------------------------
# Extensions that the reference does NOT contain (named by BASELINE.json's north_star):
# off unless a caller asks for them; oracle = oracle/vsl_oracle.py, parity unpinned.
class _Ssim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, reduce_mean):
        lib = _lib.load()
        x, y = _f32(x, 'x'), _f32(y, 'y')
        if x.shape != y.shape or x.dim() != 4:
            raise ValueError('x and y must be [B,H,W,C] of the same shape')
        B, H, W, C = x.shape
        ctx.save_for_backward(x, y)
        ctx.reduce_mean = bool(reduce_mean)
        if reduce_mean:
            out = torch.empty((), device=x.device)
            ws = _ws(lib.vsl_ssim_ws_bytes(B, H, W, C), x.device)
            check(lib.vsl_ssim_fwd(x.data_ptr(), y.data_ptr(), B, H, W, C, None, out.data_ptr(), ws.data_ptr(), _stream()))
        else:
            out = torch.empty(B, H - 2, W - 2, C, device=x.device)
            check(lib.vsl_ssim_fwd(x.data_ptr(), y.data_ptr(), B, H, W, C, out.data_ptr(), None, None, _stream()))
        return out

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        B, H, W, C = x.shape
        g = _g32(g, 'g')
        g_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        g_y = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        if g_x is None and g_y is None:
            return None, None, None
        if ctx.reduce_mean:
            check(_lib.load().vsl_ssim_bwd(x.data_ptr(), y.data_ptr(), B, H, W, C, None, g.data_ptr(), 1, _p(g_x), _p(g_y), _stream()))
        else:
            check(_lib.load().vsl_ssim_bwd(x.data_ptr(), y.data_ptr(), B, H, W, C, g.data_ptr(), None, 0, _p(g_x), _p(g_y), _stream()))
        return g_x, g_y, None


@_ingress
def ssim_dissimilarity(x, y):
    """EXTENSION (not in the reference): clip((1 - SSIM(x, y)) / 2, 0, 1), 3x3 VALID pools -> [B,H-2,W-2,C]."""
    return _Ssim.apply(x, y, False)


@_ingress
def ssim_loss(x, y):
    """EXTENSION: mean of ssim_dissimilarity(x, y), reduced inside the kernel (no map is written)."""
    return _Ssim.apply(x, y, True)


class _EdgeSmooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, img):
        lib = _lib.load()
        disp, img = _f32(disp, 'disp'), _f32(img, 'img')
        if disp.dim() != 4 or disp.shape[3] != 1 or img.dim() != 4 or img.shape[:3] != disp.shape[:3]:
            raise ValueError('disp must be [B,H,W,1] and img [B,H,W,C] of the same size')
        B, H, W, C = img.shape
        loss = torch.empty((), device=disp.device)
        ws = _ws(lib.vsl_edge_smooth_ws_bytes(B, H, W), disp.device)
        check(lib.vsl_edge_smooth_fwd(disp.data_ptr(), img.data_ptr(), B, H, W, C, loss.data_ptr(), ws.data_ptr(), _stream()))
        ctx.save_for_backward(disp, img)
        return loss

    @staticmethod
    def backward(ctx, g):
        disp, img = ctx.saved_tensors
        B, H, W, C = img.shape
        g_d = torch.empty_like(disp)
        g_i = torch.empty_like(img) if ctx.needs_input_grad[1] else None
        check(_lib.load().vsl_edge_smooth_bwd(disp.data_ptr(), img.data_ptr(), B, H, W, C, _g32(g, 'g').data_ptr(),
                                              g_d.data_ptr(), _p(g_i), _stream()))
        return g_d, g_i


@_ingress
def edge_aware_smooth_loss(disp, img):
    """EXTENSION (not in the reference): mean(|d_x disp| exp(-mean_c |d_x img|)) + the same along y."""
    return _EdgeSmooth.apply(disp, img)


def adam_step(param, grad, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0, stream=None):
    """tf.train.AdamOptimizer(lr, beta1) applied in place to one flat float32 range (train_depth_then_cam_lr.py:413;
    TensorFlow's ApplyAdam arithmetic).  param / grad / m / v: 1-D CUDA views cut at the same offset of their arenas."""
    ts = (param, grad, m, v)
    for t, name in zip(ts, ('param', 'grad', 'm', 'v')):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise TypeError('%s must be a contiguous CUDA float32 tensor (this path has no CPU fallback)' % name)
        if t.numel() != param.numel():
            raise ValueError('param, grad, m, v must have the same number of elements')
    check(_lib.load().vsl_adam_step(param.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), param.numel(),
                                    float(lr), float(beta1), float(beta2), float(eps), int(step), float(grad_scale),
                                    _stream() if stream is None else stream))


@_ingress
def image_pyramid(img, num_scales):
    """tf.image.resize_area(img, [H/2^s, W/2^s]) for s = 0..S-1 (level 0 is `img` itself). No gradient."""
    lib = _lib.load()
    img = _f32(img.detach(), 'img')
    B, H, W, C = img.shape
    levels = [torch.empty(B, H >> s, W >> s, C, device=img.device) for s in range(1, num_scales)]
    check(lib.vsl_pyramid(img.data_ptr(), B, H, W, C, num_scales, ptr_array([l.data_ptr() for l in levels]), _stream()))
    return [img] + levels


# ----------------------------------------------------------------------------------------------------
class LossFlags(object):
    """Loss configuration with the reference's FLAGS attribute names (train_depth_then_cam_lr.py:44-54)."""

    def __init__(self, **kw):
        self.num_scales = 4
        self.smooth_weight = 0.5
        self.data_weight = 1.0
        self.explain_reg_weight = 0.2
        self.pose_format = 'eular'
        self.pixel_scale_norm = True     # data_weight / 2^s (train.py:135)
        self.depth_is_inverse = True     # warp depth = 1/x (train.py:128)
        self.smooth_on_inverse = False   # smooth(1/x) (train_depth_then_cam_lr.py:217)
        self.exact_coords = False        # True: reference rounding sequence in the fused kernel (slower)
        # x_pyr = pre-activation output of the disparity head; the kernel applies disp_scaling * sigmoid + min_disp
        # (nets_optflow_depth.py:8-9,143-144) and its derivative itself
        self.x_is_logit = False
        self.disp_scaling = 4.0
        self.min_disp = 0.0
        # what the image tensors hold: 'f32' (float32, as the reference's graph sees them) or the loader's uint8
        # before its conversion -- 'u8_255': x / 255.0 (imageselect_Dataloader.py:93), 'u8_255_centred': x / 255.0 - 0.5
        # (imageselect_Dataloader_optflow_dim11.py:128), 'u8_raw': x (imageselect_Dataloader_optflow.py:129).  The
        # kernel converts on load; results are bit-identical to feeding the converted float32 images.
        self.img_format = 'f32'
        # > 0: the left-right depth-consistency term of train_depth_then_cam_lr.py:336-340 (its FLAGS.depth_weight)
        # inside the fused step; view_synthesis_loss then takes the source views' own network outputs (src_x_pyr)
        self.consist_weight = 0.0
        # EXTENSION (not in the reference): a in (0, 1] mixes a 3x3 SSIM dissimilarity into the photometric term of the
        # fused step, data_weight_s * [(1 - a) * mean(|e| m) + a * mean(D m_centre)] (VslLossDesc.ssim_weight)
        self.ssim_weight = 0.0
        self.__dict__.update(kw)


def _arena(shapes, device=None, pinned=False, dtypes=None):
    """One buffer carved into 256-byte aligned views of the given shapes (float32 unless `dtypes` says otherwise)
    -> (buffer, [views]).  Lets a whole set of tensors cross PCIe as ONE copy."""
    dtypes = dtypes or [torch.float32] * len(shapes)
    offs, n = [], 0
    for shp, dt in zip(shapes, dtypes):
        offs.append(n)
        cnt = 1
        for d in shp:
            cnt *= d
        n += (cnt * torch.empty(0, dtype=dt).element_size() + 255) // 256 * 256
    if all(dt == torch.float32 for dt in dtypes):
        buf = torch.empty(max(n, 256) // 4, dtype=torch.float32).pin_memory() if pinned else \
            torch.empty(max(n, 256) // 4, dtype=torch.float32, device=device)
        raw = buf.view(torch.uint8)
    else:
        buf = raw = torch.empty(max(n, 256), dtype=torch.uint8).pin_memory() if pinned else \
            torch.empty(max(n, 256), dtype=torch.uint8, device=device)
    views = []
    for shp, dt, o in zip(shapes, dtypes, offs):
        cnt = 1
        for d in shp:
            cnt *= d
        views.append(raw[o:o + cnt * torch.empty(0, dtype=dt).element_size()].view(dt).view(*shp))
    return buf, views


def _loss_out_shapes(B, H, W, S, V, fmt, mask_mode, consist=False):
    pose_shape = (B, V, 4, 4) if fmt == 2 else (B, V, 6)
    shapes = [(8,)] + [(B, H >> s, W >> s, 1) for s in range(S)] + [pose_shape]
    if mask_mode == _lib.MASK_EXP:
        shapes += [(B, H >> s, W >> s, 2 * V) for s in range(S)]
    if consist:                      # d/d(source views' network outputs), view-major
        shapes += [(B, H >> s, W >> s, 1) for _ in range(V) for s in range(S)]
    return shapes


class LossOutputs(object):
    """Everything one fused step produces -- losses[3], d/dx per scale, d/dposes, d/dlogits per scale -- as views of
    ONE float32 arena (a host pipeline fetches them with a single copy; the upstream gradient is applied to them
    with a single launch)."""

    def __init__(self, B, H, W, S, V, fmt, mask_mode, device, consist=False):
        self.shapes = _loss_out_shapes(B, H, W, S, V, fmt, mask_mode, consist)
        self.arena, views = _arena(self.shapes, device=device)
        self.losses = views[0][:4] if consist else views[0][:3]    # (pixel, smooth, exp[, consist])
        self.consist = views[0][3]           # the depth-consistency term (0 unless the step carries one)
        self.total = views[0][4]             # their sum, written by the same kernel
        self.g_x = views[1:1 + S]
        self.g_poses = views[1 + S]
        self.g_logits = views[2 + S:2 + 2 * S] if mask_mode == _lib.MASK_EXP else None
        self.grad_offset = 64              # floats: the gradients start at the arena's second 256-byte slot
        self._gx_ptrs = ptr_array([t.data_ptr() for t in self.g_x])
        self._gl_ptrs = ptr_array([t.data_ptr() for t in self.g_logits]) if self.g_logits else None
        self.g_src_x = views[len(views) - V * S:] if consist else None     # [v * S + s]
        self._gsx_ptrs = ptr_array([t.data_ptr() for t in self.g_src_x]) if consist else None


def _want(shape, t, name):
    if tuple(t.shape) != tuple(shape):
        raise ValueError('%s must have shape %s, got %s' % (name, tuple(shape), tuple(t.shape)))


def check_loss_shapes(B, H, W, S, V, fmt, mask_mode, tgt, srcs, x_pyr, poses, K_pyr, logits_pyr=None, mask_pyr=None,
                      src_x_pyr=None):
    """ValueError unless every tensor has the shape the fused step indexes it by."""
    if src_x_pyr is not None:
        if len(src_x_pyr) != V or any(len(p) != S for p in src_x_pyr):
            raise ValueError('src_x_pyr must hold %d views x %d levels' % (V, S))
        for v, pyr in enumerate(src_x_pyr):
            for s_, t in enumerate(pyr):
                _want((B, H >> s_, W >> s_, 1), t, 'src_x_pyr[%d][%d]' % (v, s_))
    _want((B, H, W, 3), tgt, 'tgt')
    if len(srcs) != V:
        raise ValueError('srcs must hold %d source views, got %d' % (V, len(srcs)))
    for v, t in enumerate(srcs):
        _want((B, H, W, 3), t, 'srcs[%d]' % v)
    if len(x_pyr) != S:
        raise ValueError('x_pyr must hold num_scales=%d levels (finest first), got %d' % (S, len(x_pyr)))
    for s_, t in enumerate(x_pyr):
        _want((B, H >> s_, W >> s_, 1), t, 'x_pyr[%d]' % s_)
    _want((B, V, 4, 4) if fmt == 2 else (B, V, 6), poses, 'poses')
    _want((B, S, 3, 3), K_pyr, 'K_pyr')
    if mask_mode == _lib.MASK_EXP:
        if logits_pyr is None or len(logits_pyr) != S:
            raise ValueError('logits_pyr must hold %d levels' % S)
        for s_, t in enumerate(logits_pyr):
            _want((B, H >> s_, W >> s_, 2 * V), t, 'logits_pyr[%d]' % s_)
    elif logits_pyr is not None:
        raise ValueError('logits_pyr given but the step runs without the explainability mask')
    if mask_mode == _lib.MASK_CONST:
        if mask_pyr is None or len(mask_pyr) != S:
            raise ValueError('mask_pyr must hold %d levels' % S)
        for s_, t in enumerate(mask_pyr):
            _want((B, H >> s_, W >> s_, 1), t, 'mask_pyr[%d]' % s_)
    elif mask_pyr is not None:
        raise ValueError('mask_pyr given but the step runs without a constant mask')


class ViewSynthesisPlan(object):
    """Pre-allocated state for repeated fused-loss steps at one shape: the descriptor, the workspace and (for the
    plan's own run()/run_bound()) one set of output buffers.  One instance per (shape, flags); the workspace is
    reused by every call, so calls of one plan must be ordered on one stream."""

    def __init__(self, B, H, W, V, flags, mask_mode, device, loss_scale=1.0, want_src_grad=False):
        lib = _lib.load()
        S = flags.num_scales
        self.want_src_grad = bool(want_src_grad)
        self.B, self.H, self.W, self.S, self.V = B, H, W, S, V
        self.mask_mode = mask_mode
        self.device = device
        self.fmt = _fmt(flags.pose_format)
        try:
            self.img_format = _lib.IMG_FORMATS[getattr(flags, 'img_format', 'f32')]
        except KeyError:
            raise ValueError("img_format must be one of %s" % sorted(_lib.IMG_FORMATS))
        self.img_dtype = torch.float32 if self.img_format == _lib.IMG_F32 else torch.uint8
        if self.img_format != _lib.IMG_F32 and want_src_grad:
            raise ValueError('no gradient with respect to uint8 images')
        self.consist_weight = float(getattr(flags, 'consist_weight', 0.0))
        self.consist = self.consist_weight > 0.0
        if self.consist and (want_src_grad or self.img_format != _lib.IMG_F32 or getattr(flags, 'x_is_logit', False)
                             or int(getattr(flags, 'exact_coords', False)) == 1):
            raise ValueError('the consistency term needs float32 images, exact_coords != 1, no x_is_logit and no '
                             'gradient w.r.t. the source images')
        self.desc = VslLossDesc(B, H, W, S, V, self.fmt, mask_mode, int(flags.pixel_scale_norm),
                                int(flags.depth_is_inverse), int(flags.smooth_on_inverse),
                                float(flags.data_weight), float(flags.smooth_weight),
                                float(flags.explain_reg_weight), float(loss_scale),
                                int(getattr(flags, 'exact_coords', False)), int(self.want_src_grad),
                                int(getattr(flags, 'x_is_logit', False)), float(getattr(flags, 'disp_scaling', 4.0)),
                                float(getattr(flags, 'min_disp', 0.0)), self.img_format, self.consist_weight,
                                float(getattr(flags, 'ssim_weight', 0.0)), None, None)
        nbytes = lib.vsl_loss_ws_bytes(self.desc)
        if nbytes == 0:
            raise ValueError('unsupported loss shape B=%d H=%d W=%d S=%d V=%d' % (B, H, W, S, V))
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.out = self.new_outputs()
        self.out_arena, self.out_shapes = self.out.arena, self.out.shapes
        self.losses = self.out.losses
        self.out.arena[:8].zero_()
        self.g_x, self.g_poses, self.g_logits = self.out.g_x, self.out.g_poses, self.out.g_logits
        # d/d(source images): produced only on request (an extra atomic scatter + fold-back pass)
        self.g_srcs = self.new_src_grads() if self.want_src_grad else None
        self.version = 0  # bumped by every run()

    def new_outputs(self):
        return LossOutputs(self.B, self.H, self.W, self.S, self.V, self.fmt, self.mask_mode, self.device, self.consist)

    def new_src_grads(self):
        return [torch.empty(self.B, self.H, self.W, 3, device=self.device) for _ in range(self.V)]

    def check_inputs(self, tgt, srcs, x_pyr, poses, K_pyr, logits_pyr=None, mask_pyr=None, src_x_pyr=None):
        """Every tensor's shape against the plan's (B, H, W, S, V, pose format, mask mode): the C call indexes raw
        pointers by the plan's sizes, so a mismatch must be an error here, never an out-of-bounds read there."""
        if self.consist != (src_x_pyr is not None):
            raise ValueError('src_x_pyr goes with flags.consist_weight > 0 (and only with it)')
        check_loss_shapes(self.B, self.H, self.W, self.S, self.V, self.fmt, self.mask_mode, tgt, srcs, x_pyr, poses,
                          K_pyr, logits_pyr, mask_pyr, src_x_pyr)
        for t in [poses, K_pyr] + list(x_pyr) + list(logits_pyr or []) + list(mask_pyr or []) + \
                [t for p in (src_x_pyr or []) for t in p]:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise TypeError('the fused step takes contiguous CUDA float32 tensors (no CPU fallback)')
        for t in [tgt] + list(srcs):
            if not (t.is_cuda and t.dtype == self.img_dtype and t.is_contiguous()):
                raise TypeError('images must be contiguous CUDA %s tensors for img_format=%r (no CPU fallback)'
                                % (self.img_dtype, self.img_format))

    def bind(self, tgt, srcs, x_pyr, poses, K_pyr, logits_pyr=None, mask_pyr=None, out=None, g_srcs=None,
             src_x_pyr=None):
        """Validate and pre-marshal the C arguments for one set of input (and output) buffers; run_bound(args) then
        costs one ctypes call.  The caller keeps the tensors alive."""
        self.check_inputs(tgt, srcs, x_pyr, poses, K_pyr, logits_pyr, mask_pyr, src_x_pyr)
        out = self.out if out is None else out
        g_srcs = self.g_srcs if g_srcs is None else g_srcs
        if self.consist:                             # vsl_loss_consist_fwd_bwd
            return (self.desc, tgt.data_ptr(), ptr_array([s.data_ptr() for s in srcs]),
                    ptr_array([x.data_ptr() for x in x_pyr]),
                    ptr_array([t.data_ptr() for p in src_x_pyr for t in p]), poses.data_ptr(), K_pyr.data_ptr(),
                    ptr_array([l.data_ptr() for l in logits_pyr]) if logits_pyr is not None else None,
                    ptr_array([m.data_ptr() for m in mask_pyr]) if mask_pyr is not None else None,
                    out.losses.data_ptr(), out._gx_ptrs, out._gsx_ptrs, out.g_poses.data_ptr(), out._gl_ptrs,
                    self.ws.data_ptr())
        head = (self.desc, tgt.data_ptr(), ptr_array([s.data_ptr() for s in srcs]),
                ptr_array([x.data_ptr() for x in x_pyr]), poses.data_ptr(), K_pyr.data_ptr(),
                ptr_array([l.data_ptr() for l in logits_pyr]) if logits_pyr is not None else None,
                ptr_array([m.data_ptr() for m in mask_pyr]) if mask_pyr is not None else None,
                out.losses.data_ptr(), out._gx_ptrs, out.g_poses.data_ptr(), out._gl_ptrs)
        if self.img_format != _lib.IMG_F32:          # vsl_loss_fwd_bwd_u8: no d/d(source images)
            return head + (self.ws.data_ptr(),)
        return head + (ptr_array([t.data_ptr() for t in g_srcs]) if g_srcs else None, self.ws.data_ptr())

    def run_bound(self, args, stream=None):
        self.version += 1
        lib = _lib.load()
        fn = lib.vsl_loss_consist_fwd_bwd if self.consist else (
            lib.vsl_loss_fwd_bwd if self.img_format == _lib.IMG_F32 else lib.vsl_loss_fwd_bwd_u8)
        check(fn(*args, _stream() if stream is None else stream))
        return self.losses

    def prep_levels(self):
        """Views of what the prep launch of the last step left in the workspace: (target RGB levels [B,Hs,Ws,3] -- level
        0 is None unless the images are uint8 --, per source view its zero-bordered RGBA levels [B,Hs+4,Ws+4,4]).
        For tests and tools."""
        import ctypes
        n = self.S + self.V * self.S
        offs = (ctypes.c_longlong * n)()
        check(_lib.load().vsl_loss_ws_layout(self.desc, offs))

        def view(o, shape):
            cnt = 1
            for d in shape:
                cnt *= d
            return self.ws[o:o + 4 * cnt].view(torch.float32).view(*shape)
        B, H, W, S, V = self.B, self.H, self.W, self.S, self.V
        tgt = [view(offs[s], (B, H >> s, W >> s, 3)) if offs[s] >= 0 else None for s in range(S)]
        srcs = [[view(offs[S + v * S + s], (B, (H >> s) + 4, (W >> s) + 4, 4)) for s in range(S)] for v in range(V)]
        return tgt, srcs

    def set_profile_events(self, begin=None, end=None):
        """cudaEvent_t handles (ints) recorded immediately around the fused loss kernel; None switches off."""
        self.desc.ev_main_begin, self.desc.ev_main_end = begin, end

    def run(self, tgt, srcs, x_pyr, poses, K_pyr, logits_pyr=None, mask_pyr=None, src_x_pyr=None):
        """Enqueue one fused forward+backward.  Inputs must already be contiguous CUDA float32 of the plan's
        shapes.  Results land in self.losses / self.g_x / self.g_poses / self.g_logits (/ self.out.g_src_x)."""
        return self.run_bound(self.bind(tgt, srcs, x_pyr, poses, K_pyr, logits_pyr, mask_pyr, src_x_pyr=src_x_pyr))


class _ViewSynthesisLoss(torch.autograd.Function):
    """Forward runs the fused step: the gradients of (pixel + smooth + exp) * loss_scale come out of the same kernel
    pass, into an output arena that belongs to THIS call (so any number of forwards may precede their backwards).
    Backward applies the upstream gradient of `total` to that arena with one launch (vsl_scale; a no-op kernel when
    it is exactly 1, i.e. total.backward()) and hands out views of it."""

    @staticmethod
    def forward(ctx, plan, tgt, K_pyr, poses, n_src, n_x, n_rest, *rest_all):
        srcs = [_img(t, 'src', plan.img_dtype) for t in rest_all[:n_src]]
        x_pyr = [_f32(t, 'x_pyr') for t in rest_all[n_src:n_src + n_x]]
        rest = [_f32(t, 'pyr') for t in rest_all[n_src + n_x:n_src + n_x + n_rest]]
        src_x = [_f32(t, 'src_x_pyr') for t in rest_all[n_src + n_x + n_rest:]]       # view-major, V * S or none
        logits = rest if plan.mask_mode == _lib.MASK_EXP else None
        mask = rest if plan.mask_mode == _lib.MASK_CONST else None
        out = plan.new_outputs()
        g_srcs = plan.new_src_grads() if plan.want_src_grad else None
        src_x_pyr = [src_x[v * n_x:(v + 1) * n_x] for v in range(n_src)] if src_x else None
        plan.run_bound(plan.bind(_img(tgt, 'tgt', plan.img_dtype), srcs, x_pyr, _f32(poses, 'poses'), _f32(K_pyr, 'K_pyr'), logits,
                                 mask, out=out, g_srcs=g_srcs, src_x_pyr=src_x_pyr))
        ctx.out, ctx.g_srcs, ctx.mask_mode = out, g_srcs, plan.mask_mode
        ctx.n_src, ctx.n_x, ctx.n_rest, ctx.n_sx = n_src, n_x, len(rest), len(src_x)
        ctx.applied = None                    # device scalar: the upstream factor the arena currently carries
        # views of this call's own arena: no copy, no reduction launch (the finalize kernel wrote the total too)
        total, losses = out.total.view(()), out.losses.view(-1)
        ctx.mark_non_differentiable(losses)
        return total, losses

    @staticmethod
    def backward(ctx, g_total, _g_losses):
        lib, out = _lib.load(), ctx.out
        g = g_total.detach().to(torch.float32).reshape(1).contiguous()
        n = out.arena.numel() - out.grad_offset
        src_ptr = out.arena.data_ptr() + 4 * out.grad_offset
        if ctx.applied is None:               # the usual single backward: in place
            check(lib.vsl_scale(src_ptr, src_ptr, n, g.data_ptr(), None, _stream()))
            ctx.applied = g
            g_x, g_poses, g_logits = out.g_x, out.g_poses, out.g_logits
            g_src_x = out.g_src_x
            g_srcs = ctx.g_srcs
            if g_srcs is not None:
                for t in g_srcs:
                    check(lib.vsl_scale(t.data_ptr(), t.data_ptr(), t.numel(), g.data_ptr(), None, _stream()))
        else:                                 # backward again (retain_graph): fresh buffers, factor g / applied
            fresh = LossOutputs.__new__(LossOutputs)
            fresh.arena = torch.empty_like(out.arena)
            check(lib.vsl_scale(fresh.arena.data_ptr() + 4 * out.grad_offset, src_ptr, n, g.data_ptr(),
                                ctx.applied.data_ptr(), _stream()))
            views, o = [], 0
            for shp in out.shapes:
                cnt = 1
                for d in shp:
                    cnt *= d
                views.append(fresh.arena[o:o + cnt].view(*shp))
                o += (cnt + 63) // 64 * 64
            S = len(out.g_x)
            g_x, g_poses = views[1:1 + S], views[1 + S]
            g_logits = views[2 + S:2 + 2 * S] if out.g_logits is not None else None
            g_src_x = views[len(views) - ctx.n_sx:] if ctx.n_sx else None
            g_srcs = None
            if ctx.g_srcs is not None:
                g_srcs = [torch.empty_like(t) for t in ctx.g_srcs]
                for d, t in zip(g_srcs, ctx.g_srcs):
                    check(lib.vsl_scale(d.data_ptr(), t.data_ptr(), t.numel(), g.data_ptr(), ctx.applied.data_ptr(),
                                        _stream()))
        gs = tuple(g_srcs) if g_srcs is not None else (None,) * ctx.n_src
        gr = tuple(g_logits) if ctx.mask_mode == _lib.MASK_EXP else (None,) * ctx.n_rest
        return (None, None, None, g_poses, None, None, None) + gs + tuple(g_x) + gr + tuple(g_src_x or ())


_PLANS = collections.OrderedDict()
_PLANS_LOCK = threading.Lock()
_PLANS_MAX = 8


def _plan_for(key, make):
    """Bounded (LRU, _PLANS_MAX entries) and thread-safe cache of plans: a plan owns a workspace as large as the
    re-laid source pyramids, so shapes that are no longer used must not pin device memory for ever."""
    with _PLANS_LOCK:
        plan = _PLANS.get(key)
        if plan is not None:
            _PLANS.move_to_end(key)
            return plan
        plan = _PLANS[key] = make()
        while len(_PLANS) > _PLANS_MAX:
            _PLANS.popitem(last=False)
        return plan


def view_synthesis_loss(tgt, srcs, x_pyr, poses, K_pyr, logits_pyr=None, mask_pyr=None, flags=None, loss_scale=1.0,
                        src_x_pyr=None):
    """The reference's per-scale loss loop (train.py:107-135 + train_depth_then_cam_lr.py:297-328) as ONE fused
    forward+backward call.

    tgt [B,H,W,3]; srcs: list of V [B,H,W,3]; x_pyr: list of S network outputs [B,Hs,Ws,1] (finest first); poses
    [B,V,6] or [B,V,4,4]; K_pyr [B,S,3,3]; logits_pyr: list of S [B,Hs,Ws,2V] (explainability) or mask_pyr: list of S
    constant weights [B,Hs,Ws,1].  Every shape is checked (ValueError) before anything reaches the C call.
    -> (total, losses[3] = pixel, smooth, exp).  `total` is differentiable wrt x_pyr, poses and logits_pyr (and the
    source images if they require grad); the gradients were produced in the same kernel pass as the loss.
    With flags.consist_weight > 0 the step also carries the left-right depth-consistency term
    (train_depth_then_cam_lr.py:336-340): src_x_pyr = per source view the list of ITS S network outputs
    [B,Hs,Ws,1] (depth = x or 1/x like the target's); losses becomes [4] = pixel, smooth, exp, consist and `total`
    is differentiable wrt src_x_pyr too.
    loss_scale: a constant factor of the objective known up front (a data-parallel rank's B_local / B_global share,
    dist.local_loss_scale) -- folded into the kernel's gradients for free; `total` and `losses` stay unscaled means.
    """
    flags = flags or LossFlags()
    if logits_pyr is not None and mask_pyr is not None:
        raise ValueError('give logits_pyr or mask_pyr, not both')
    tgt = from_external(tgt, 'tgt')
    if not tgt.is_cuda:
        raise TypeError('view_synthesis_loss takes CUDA tensors (this path has no CPU fallback)')
    srcs = [from_external(t, 'srcs') for t in srcs]
    x_pyr = [from_external(t, 'x_pyr') for t in x_pyr]
    poses, K_pyr = from_external(poses, 'poses'), from_external(K_pyr, 'K_pyr')
    if tgt.dim() != 4 or tgt.shape[3] != 3:
        raise ValueError('tgt must be [B,H,W,3], got %s' % (tuple(tgt.shape),))
    B, H, W, C = tgt.shape
    V, S = len(srcs), flags.num_scales
    if len(x_pyr) != S:
        raise ValueError('x_pyr must hold num_scales=%d levels' % S)
    mode = _lib.MASK_EXP if logits_pyr is not None else (_lib.MASK_CONST if mask_pyr is not None else _lib.MASK_NONE)
    want_src = any(getattr(t, 'requires_grad', False) for t in srcs)
    rest = [from_external(t, 'pyr') for t in (logits_pyr if logits_pyr is not None else (mask_pyr or []))]
    consist = float(getattr(flags, 'consist_weight', 0.0)) > 0.0
    if consist != (src_x_pyr is not None):
        raise ValueError('src_x_pyr goes with flags.consist_weight > 0 (and only with it)')
    if consist:
        src_x_pyr = [[from_external(t, 'src_x_pyr') for t in p] for p in src_x_pyr]
    # shapes first: a wrong layout is a ValueError here, before any pointer reaches the library
    check_loss_shapes(B, H, W, S, V, _fmt(flags.pose_format), mode, tgt, srcs, x_pyr, poses, K_pyr,
                      rest if mode == _lib.MASK_EXP else None, rest if mode == _lib.MASK_CONST else None, src_x_pyr)
    key = (B, H, W, V, mode, tgt.device, torch.cuda.current_stream(tgt.device).cuda_stream, want_src, float(loss_scale),
           tuple(sorted(flags.__dict__.items())))
    plan = _plan_for(key, lambda: ViewSynthesisPlan(B, H, W, V, flags, mode, tgt.device, loss_scale=loss_scale,
                                                    want_src_grad=want_src))
    return _ViewSynthesisLoss.apply(plan, tgt, K_pyr, poses, V, S, len(rest), *srcs, *x_pyr, *rest,
                                    *[t for p in (src_x_pyr or []) for t in p])


class FlowLossFlags(object):
    """The FLAGS the loss loop of train_optflow_combine.py:138-240 reads (reference attribute names)."""

    def __init__(self, **kw):
        self.num_scales = 4
        self.smooth_weight = 0.5
        self.depth_weight = 1.0
        self.data_weight = 1.0
        self.optflow_weight = 1.0
        self.__dict__.update(kw)


class _FlowDepthLoss(torch.autograd.Function):
    """One vsl_flow_loss_fwd_bwd call: the gradients of (depth + smooth + optflow + pixel) * loss_scale w.r.t. the
    three prediction pyramids come out of the same pass, into an arena that belongs to this call; backward applies
    the upstream gradient of `total` to it with one launch (vsl_scale, a no-op kernel for total.backward())."""

    @staticmethod
    def forward(ctx, desc, left, right, label, proj, K_pyr, S, *pyrs):
        lib = _lib.load()
        left, right, label = _f32(left, 'image_left'), _f32(right, 'image_right'), _f32(label, 'label')
        proj, K_pyr = _f32(proj, 'tgt2src_proj'), _f32(K_pyr, 'K_pyr')
        pyrs = [_f32(t, 'pyr') for t in pyrs]
        B, H, W = desc.B, desc.H, desc.W
        shapes = [(8,)] + [(B, H >> s, W >> s, 1) for _ in range(3) for s in range(S)]
        arena, views = _arena(shapes, device=left.device)
        ws = _ws(lib.vsl_flow_loss_ws_bytes(ctypes.byref(desc)), left.device)
        ptrs = lambda ts: ptr_array([t.data_ptr() for t in ts])
        check(lib.vsl_flow_loss_fwd_bwd(ctypes.byref(desc), left.data_ptr(), right.data_ptr(), label.data_ptr(),
                                        ptrs(pyrs[:S]), ptrs(pyrs[S:2 * S]), ptrs(pyrs[2 * S:]), proj.data_ptr(),
                                        K_pyr.data_ptr(), views[0].data_ptr(), ptrs(views[1:1 + S]),
                                        ptrs(views[1 + S:1 + 2 * S]), ptrs(views[1 + 2 * S:]), ws.data_ptr(), _stream()))
        ctx.arena, ctx.grads, ctx.applied = arena, views[1:], None
        total, losses = views[0][4].view(()), views[0][:4]
        ctx.mark_non_differentiable(losses)
        return total, losses

    @staticmethod
    def backward(ctx, g_total, _g_losses):
        lib = _lib.load()
        if ctx.applied is not None:
            raise RuntimeError('flow_depth_loss: backward through the same call twice is not supported')
        g = g_total.detach().to(torch.float32).reshape(1).contiguous()
        n = ctx.arena.numel() - 64
        p = ctx.arena.data_ptr() + 4 * 64
        check(lib.vsl_scale(p, p, n, g.data_ptr(), None, _stream()))
        ctx.applied = g
        return (None,) * 7 + tuple(ctx.grads)


def flow_depth_loss(image_left, image_right, label, pred_depth, pred_optflow_x, pred_optflow_y, tgt2src_proj, K_pyr,
                    flags=None, loss_scale=1.0):
    """The loss loop of train_optflow_combine.py:138-240 (the DeMoN-pair family) as ONE fused forward+backward call.

    image_left / image_right [B,H,W,3]; label [B,H,W,1] ground-truth inverse depth; pred_depth / pred_optflow_x /
    pred_optflow_y: lists of S network outputs [B,Hs,Ws,1] (finest first; inverse depth, flow in pixels of that
    scale); tgt2src_proj [B,4,4] (the loader's tgt2src_projs[:,0]); K_pyr [B,S,3,3].
    -> (total, losses[4] = depth, smooth, optflow, pixel).  `total` (= total_loss, :240) is differentiable wrt the
    three prediction pyramids; the label, the images and the pose are data."""
    flags = flags or FlowLossFlags()
    image_left = from_external(image_left, 'image_left')
    if not image_left.is_cuda:
        raise TypeError('flow_depth_loss takes CUDA tensors (this path has no CPU fallback)')
    if image_left.dim() != 4 or image_left.shape[3] != 3:
        raise ValueError('image_left must be [B,H,W,3], got %s' % (tuple(image_left.shape),))
    B, H, W, _ = image_left.shape
    S = flags.num_scales
    image_right, label = from_external(image_right, 'image_right'), from_external(label, 'label')
    tgt2src_proj, K_pyr = from_external(tgt2src_proj, 'tgt2src_proj'), from_external(K_pyr, 'K_pyr')
    _want((B, H, W, 3), image_right, 'image_right')
    _want((B, H, W, 1), label, 'label')
    _want((B, 4, 4), tgt2src_proj, 'tgt2src_proj')
    _want((B, S, 3, 3), K_pyr, 'K_pyr')
    pyrs = []
    for name, pyr in (('pred_depth', pred_depth), ('pred_optflow_x', pred_optflow_x), ('pred_optflow_y', pred_optflow_y)):
        if len(pyr) != S:
            raise ValueError('%s must hold num_scales=%d levels' % (name, S))
        for sc, t in enumerate(pyr):
            t = from_external(t, name)
            _want((B, H >> sc, W >> sc, 1), t, '%s[%d]' % (name, sc))
            pyrs.append(t)
    if H % (1 << (S - 1)) or W % (1 << (S - 1)) or (H >> (S - 1)) < 3 or (W >> (S - 1)) < 3:
        raise ValueError('H, W must be divisible by 2^(num_scales-1) with a coarsest level of at least 3 x 3')
    desc = _lib.VslFlowLossDesc(B=B, H=H, W=W, S=S, smooth_weight=flags.smooth_weight, depth_weight=flags.depth_weight,
                                data_weight=flags.data_weight, optflow_weight=flags.optflow_weight,
                                loss_scale=float(loss_scale))
    return _FlowDepthLoss.apply(desc, image_left, image_right, label, tgt2src_proj, K_pyr, S, *pyrs)


class HostPipeline(object):
    """Fused loss steps fed from HOST memory: inputs arrive in pinned host buffers, losses and gradients are
    returned in pinned host buffers.  Three streams (H2D, compute, D2H) and two buffer sets overlap the copies of
    neighbouring steps with the kernels, so steady-state throughput is max(H2D, compute, D2H), not their sum.

    submit(host_inputs) enqueues one step and returns its slot; result(slot) waits for that step's D2H and returns
    (losses[3], g_x list, g_poses, g_logits list) as pinned host tensors (valid until the slot is reused, i.e.
    until two more submits).  host_inputs: dict(tgt, srcs, xs, poses, Kp, lgs) of pinned CPU tensors."""

    DEPTH = 2

    def __init__(self, B, H, W, V, flags, mask_mode, device, loss_scale=1.0):
        self.device = device
        self.plans = [ViewSynthesisPlan(B, H, W, V, flags, mask_mode, device, loss_scale) for _ in range(self.DEPTH)]
        S = flags.num_scales
        pose_shape = (B, V, 4, 4) if self.plans[0].fmt == 2 else (B, V, 6)
        in_shapes = ([(B, H, W, 3)] + [(B, H, W, 3)] * V + [(B, H >> s, W >> s, 1) for s in range(S)] +
                     [pose_shape, (B, S, 3, 3)] +
                     ([(B, H >> s, W >> s, 2 * V) for s in range(S)] if mask_mode == _lib.MASK_EXP else []))
        # images in the dtype the flags name (uint8 = what the reference's loader holds before `/ 255.0`: a quarter of
        # the bytes over PCIe), everything else float32
        in_dtypes = [self.plans[0].img_dtype] * (1 + V) + [torch.float32] * (len(in_shapes) - 1 - V)

        def carve(views):
            return dict(tgt=views[0], srcs=views[1:1 + V], xs=views[1 + V:1 + V + S], poses=views[1 + V + S],
                        Kp=views[2 + V + S], lgs=(views[3 + V + S:3 + V + 2 * S] if mask_mode == _lib.MASK_EXP else None))
        self._in_shapes, self._in_dtypes, self._carve = in_shapes, in_dtypes, carve
        # device inputs of a slot = one arena; host inputs allocated by host_inputs() mirror it, so a step's
        # inputs cross PCIe as ONE copy (and its outputs likewise, from the plan's output arena)
        self.dev_in_arena, self.dev_in = [], []
        for _ in range(self.DEPTH):
            buf, views = _arena(in_shapes, device=device, dtypes=in_dtypes)
            self.dev_in_arena.append(buf)
            self.dev_in.append(carve(views))
        self.bound = [p.bind(d['tgt'], d['srcs'], d['xs'], d['poses'], d['Kp'], d['lgs'])
                      for p, d in zip(self.plans, self.dev_in)]
        self.host_out_arena, self.host_out = [], []
        for p in self.plans:
            buf, views = _arena(p.out_shapes, pinned=True)
            self.host_out_arena.append(buf)
            self.host_out.append(dict(losses=views[0][:3], total=views[0][4], g_x=views[1:1 + S], g_poses=views[1 + S],
                                      g_lgs=(views[2 + S:2 + 2 * S] if mask_mode == _lib.MASK_EXP else [])))
        self.s_in, self.s_comp, self.s_out = (torch.cuda.Stream(device=device) for _ in range(3))
        # the plans' buffers were initialised on the current stream: nothing on the side streams may overtake that
        cur = torch.cuda.current_stream(device)
        for st in (self.s_in, self.s_comp, self.s_out):
            st.wait_stream(cur)
        ev = lambda: [torch.cuda.Event() for _ in range(self.DEPTH)]
        self.ev_in, self.ev_comp, self.ev_out = ev(), ev(), ev()
        self.step = 0

    @staticmethod
    def _flat(d):
        out = []
        for k in ('tgt', 'srcs', 'xs', 'poses', 'Kp', 'lgs'):
            v = d.get(k)
            if v is None:
                continue
            out.extend(v if isinstance(v, (list, tuple)) else [v])
        return out

    def host_inputs(self):
        """A dict of pinned host tensors (tgt, srcs, xs, poses, Kp, lgs) that are views into ONE pinned arena laid out
        like the device-side inputs: submit() then moves a step's inputs with a single copy.  Fill them in place."""
        buf, views = _arena(self._in_shapes, pinned=True, dtypes=self._in_dtypes)
        d = self._carve(views)
        d['_arena'] = buf
        return d

    def bytes_per_step(self):
        h2d = sum(t.numel() * t.element_size() for t in self._flat(self.dev_in[0]))
        o = self.host_out[0]
        d2h = sum(t.numel() * 4 for t in [o['losses'], o['g_poses']] + o['g_x'] + o['g_lgs'])
        return h2d, d2h

    def submit(self, host_inputs):
        k = self.step % self.DEPTH
        plan, dev, out = self.plans[k], self.dev_in[k], self.host_out[k]
        first_use = self.step < self.DEPTH
        with torch.cuda.stream(self.s_in):
            if not first_use:
                self.s_in.wait_event(self.ev_comp[k])          # step-2's kernels have consumed these inputs
            if host_inputs.get('_arena') is not None:
                self.dev_in_arena[k].copy_(host_inputs['_arena'], non_blocking=True)
            else:
                for src, dst in zip(self._flat(host_inputs), self._flat(dev)):
                    dst.copy_(src, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        with torch.cuda.stream(self.s_comp):
            self.s_comp.wait_event(self.ev_in[k])
            if not first_use:
                self.s_comp.wait_event(self.ev_out[k])          # step-2's gradients have left the device
            plan.run_bound(self.bound[k], self.s_comp.cuda_stream)
            self.ev_comp[k].record(self.s_comp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_comp[k])
            self.host_out_arena[k].copy_(plan.out_arena, non_blocking=True)
            self.ev_out[k].record(self.s_out)
        self.step += 1
        return k

    def result(self, slot):
        self.ev_out[slot].synchronize()
        o = self.host_out[slot]
        return o['losses'], o['g_x'], o['g_poses'], o['g_lgs']
