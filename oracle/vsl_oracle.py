"""CPU oracle for the view-synthesis-loss hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (tf_depth_estimation_b200/) never
does; it fails loudly when the CUDA library is missing.

A restatement, in torch-CPU tensor ops, of the arithmetic the reference performs
(file:line citations are relative to /root/reference).  It keeps the reference's
*unfused* structure -- four flat gathers per sampler call, one tensor op per
reference op, autograd for the backward -- so that it is also an honest stand-in
for "the reference's CPU path" when bench.py times it.

PARITY STATUS.  The reference holds no golden vectors, tests or fixtures for this
path and its runtime (TensorFlow 1.x) cannot be installed
here.  The oracle is pinned instead against the reference's OWN SOURCE executed
over a torch-backed TF1 API shim (oracle/tf1_shim, driven by
tests/golden/make_golden.py; fixtures in tests/golden/*.npz).  That pins the op
graph, op order and constants; it does not pin TF's C++ kernel numerics, whose
conventions are stated in oracle/tf1_shim/tensorflow/__init__.py.  SSIM and
edge-aware smoothness (named by BASELINE.json but absent from the reference) are
"parity unpinned -- no reference implementation".

Arithmetic conventions shared with the CUDA kernels (so that coordinates can be
compared bit-for-bit when the pose is given as a matrix):
  * small matmuls are sequential sums over k of separate mul and add (no FMA);
  * the pixel grid is the reference's fp32 linspace grid, not exact integers;
  * K^-1 is a row-pivoted LU inverse (what tf.matrix_inverse/Eigen does);
  * gather indices are exact integers (the reference computes them in fp32 and
    is wrong beyond 2^24 elements -- SURVEY.md D10; below that they agree).
"""
import math

import numpy as np
import torch

EPS_Z = 1e-10  # utils.py:136-137


# ----------------------------------------------------------------------------
# pose  (utils.py:26-98, utils_lr.py:26-149)
# ----------------------------------------------------------------------------
def _mm(a, b):
    """[..., M, K] x [..., K, N], K small: sequential k, separate mul/add."""
    acc = a[..., :, 0:1] * b[..., 0:1, :]
    for k in range(1, a.shape[-1]):
        acc = acc + a[..., :, k:k + 1] * b[..., k:k + 1, :]
    return acc


def euler2mat(z, y, x):
    """utils.py:26-75.  z, y, x: [B, 1] -> R = Rx @ Ry @ Rz, [B, 1, 3, 3]."""
    z = torch.clamp(z, -math.pi, math.pi)[..., None, None]
    y = torch.clamp(y, -math.pi, math.pi)[..., None, None]
    x = torch.clamp(x, -math.pi, math.pi)[..., None, None]
    o, l = torch.zeros_like(z), torch.ones_like(z)
    cz, sz, cy, sy, cx, sx = z.cos(), z.sin(), y.cos(), y.sin(), x.cos(), x.sin()

    def rows(r0, r1, r2):
        return torch.cat([torch.cat(r, dim=3) for r in (r0, r1, r2)], dim=2)

    zm = rows([cz, -sz, o], [sz, cz, o], [o, o, l])
    ym = rows([cy, o, sy], [o, l, o], [-sy, o, cy])
    xm = rows([l, o, o], [o, cx, -sx], [o, sx, cx])
    return _mm(_mm(xm, ym), zm)


def axis_angle_to_rotation_matrix(axis, angle):
    """utils_lr.py:77-103.  axis [B,3] (unit), angle [B,1,1] -> I + sin*A + (1-cos)*A@A."""
    B = axis.shape[0]
    o = torch.zeros(B, dtype=axis.dtype)
    a0, a1, a2 = axis[:, 0], axis[:, 1], axis[:, 2]
    M = torch.stack([torch.stack([o, -a2, a1], 1),
                     torch.stack([o, o, -a0], 1),
                     torch.stack([o, o, o], 1)], 1)
    A = M - M.transpose(1, 2)
    eye = torch.eye(3, dtype=axis.dtype).expand(B, 3, 3)
    one = torch.ones(B, 1, 1, dtype=axis.dtype)
    return eye + torch.sin(angle) * A + (one - torch.cos(angle)) * _mm(A, A)


def pose_vec2mat(vec, format='eular'):
    """utils.py:79-98 / utils_lr.py:106-149.  vec [B,6] = tx,ty,tz,r0,r1,r2 -> [B,4,4].

    'eular': R = euler2mat(rz=r2, ry=r1, rx=r0).  'angleaxis': theta = |r|, axis = r/theta
    (NaN at theta == 0, like the reference).
    """
    B = vec.shape[0]
    t = vec[:, 0:3].unsqueeze(-1)
    if format == 'eular':
        R = euler2mat(vec[:, 5:6], vec[:, 4:5], vec[:, 3:4]).squeeze(1)
    elif format == 'angleaxis':
        r = vec[:, 3:6]
        angle = torch.sqrt(torch.sum(r * r, dim=1)).unsqueeze(-1)
        R = axis_angle_to_rotation_matrix(r / angle, angle.unsqueeze(-1))
    else:
        raise ValueError(format)
    bottom = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=vec.dtype).reshape(1, 1, 4).repeat(B, 1, 1)
    return torch.cat([torch.cat([R, t], dim=2), bottom], dim=1)


# ----------------------------------------------------------------------------
# grid, intrinsics  (utils.py:142-166, Demon_Data_loader.py:14-39)
# ----------------------------------------------------------------------------
def grid_1d(n, dtype=torch.float32):
    """One axis of meshgrid (utils.py:153-159): (linspace(-1,1,n)+1)*0.5*(n-1), TF1 linspace."""
    one = torch.tensor(1.0, dtype=dtype)
    step = (one - (-one)) / torch.tensor(n - 1, dtype=dtype)
    lin = (-one) + step * torch.arange(n).to(dtype)
    return (lin + 1.0) * 0.5 * torch.tensor(n - 1, dtype=dtype)


def meshgrid(batch, height, width, is_homogeneous=True, dtype=torch.float32):
    """utils.py:142-166 -> [B, 3|2, H, W]."""
    xs = grid_1d(width, dtype).reshape(1, width).expand(height, width)
    ys = grid_1d(height, dtype).reshape(height, 1).expand(height, width)
    planes = [xs, ys] + ([torch.ones(height, width, dtype=dtype)] if is_homogeneous else [])
    return torch.stack(planes, 0).unsqueeze(0).repeat(batch, 1, 1, 1)


def lu_inverse_np(m):
    """Row-pivoted LU inverse of one n x n matrix in m.dtype (tf.matrix_inverse, utils.py:114)."""
    n = m.shape[0]
    f = m.dtype.type
    a = m.copy()
    perm = list(range(n))
    for k in range(n):
        p = k + int(np.argmax(np.abs(a[k:, k])))
        if p != k:
            a[[k, p]] = a[[p, k]]
            perm[k], perm[p] = perm[p], perm[k]
        for i in range(k + 1, n):
            a[i, k] = f(a[i, k] / a[k, k])
            for j in range(k + 1, n):
                a[i, j] = f(a[i, j] - f(a[i, k] * a[k, j]))
    inv = np.zeros_like(m)
    for c in range(n):
        y = np.zeros(n, dtype=m.dtype)
        for i in range(n):
            s = f(1.0) if perm[i] == c else f(0.0)
            for j in range(i):
                s = f(s - f(a[i, j] * y[j]))
            y[i] = s
        for i in range(n - 1, -1, -1):
            s = y[i]
            for j in range(i + 1, n):
                s = f(s - f(a[i, j] * inv[j, c]))
            inv[i, c] = f(s / a[i, i])
    return inv


def matrix_inverse(K):
    """Batched small inverse; K carries no gradient anywhere in the reference."""
    a = K.detach().numpy().reshape(-1, K.shape[-2], K.shape[-1])
    out = np.stack([lu_inverse_np(a[i]) for i in range(a.shape[0])])
    return torch.from_numpy(out.reshape(tuple(K.shape)))


def multi_scale_intrinsics(K, num_scales):
    """Demon_Data_loader.py:25-39: fx,fy,cx,cy / 2^s -> [B, S, 3, 3]."""
    out = []
    for s in range(num_scales):
        Ks = torch.zeros_like(K)
        Ks[:, 0, 0] = K[:, 0, 0] / (2 ** s)
        Ks[:, 1, 1] = K[:, 1, 1] / (2 ** s)
        Ks[:, 0, 2] = K[:, 0, 2] / (2 ** s)
        Ks[:, 1, 2] = K[:, 1, 2] / (2 ** s)
        Ks[:, 2, 2] = 1.0
        out.append(Ks)
    return torch.stack(out, dim=1)


# ----------------------------------------------------------------------------
# geometry  (utils.py:100-140, utils_lr.py:152-194)
# ----------------------------------------------------------------------------
def pixel2cam(depth, pixel_coords, intrinsics, is_homogeneous=True):
    """utils.py:100-119: cam = (K^-1 @ p) * depth -> [B, 4|3, H, W]."""
    B, H, W = depth.shape
    p = pixel_coords.reshape(B, 3, -1)
    cam = _mm(matrix_inverse(intrinsics).to(depth.dtype), p) * depth.reshape(B, 1, -1)
    if is_homogeneous:
        cam = torch.cat([cam, torch.ones(B, 1, H * W, dtype=depth.dtype)], dim=1)
    return cam.reshape(B, -1, H, W)


def cam2pixel(cam_coords, proj):
    """utils.py:121-140 / utils_lr.py:172-194 -> coords [B,H,W,2], z_u [B,H,W,1]."""
    B, _, H, W = cam_coords.shape
    u = _mm(proj, cam_coords.reshape(B, 4, -1))
    z = u[:, 2:3]
    xn = u[:, 0:1] / (z + EPS_Z)
    yn = u[:, 1:2] / (z + EPS_Z)
    coords = torch.cat([xn, yn], dim=1).reshape(B, 2, H, W).permute(0, 2, 3, 1)
    return coords, z.reshape(B, H, W, 1)


def bilinear_sampler(imgs, coords):
    """utils.py:219-308: zero-padded bilinear gather.  -> (out [B,Ht,Wt,C], wmask [B,Ht,Wt,1])."""
    B, Hs, Ws, C = imgs.shape
    _, Ht, Wt, _ = coords.shape
    x, y = coords[..., 0:1], coords[..., 1:2]
    x0 = torch.floor(x)
    x1 = x0 + 1
    y0 = torch.floor(y)
    y1 = y0 + 1
    xmax, ymax = float(Ws - 1), float(Hs - 1)
    x0s, x1s = torch.clamp(x0, 0.0, xmax), torch.clamp(x1, 0.0, xmax)
    y0s, y1s = torch.clamp(y0, 0.0, ymax), torch.clamp(y1, 0.0, ymax)
    wx0 = (x1 - x) * (x0 == x0s).to(x.dtype)
    wx1 = (x - x0) * (x1 == x1s).to(x.dtype)
    wy0 = (y1 - y) * (y0 == y0s).to(x.dtype)
    wy1 = (y - y0) * (y1 == y1s).to(x.dtype)
    base = (torch.arange(B, dtype=torch.int64) * (Hs * Ws)).reshape(B, 1, 1, 1)
    xi0, xi1 = x0s.to(torch.int64), x1s.to(torch.int64)
    row0, row1 = base + y0s.to(torch.int64) * Ws, base + y1s.to(torch.int64) * Ws
    flat = imgs.reshape(-1, C)
    shp = (B, Ht, Wt, C)
    im00 = flat[(xi0 + row0).reshape(-1)].reshape(shp)
    im01 = flat[(xi0 + row1).reshape(-1)].reshape(shp)
    im10 = flat[(xi1 + row0).reshape(-1)].reshape(shp)
    im11 = flat[(xi1 + row1).reshape(-1)].reshape(shp)
    w00, w01, w10, w11 = wx0 * wy0, wx0 * wy1, wx1 * wy0, wx1 * wy1
    out = ((w00 * im00 + w01 * im01) + w10 * im10) + w11 * im11
    wmask = ((w00 + w01) + w10) + w11
    return out, wmask


def projective_inverse_warp(img, depth, pose, intrinsics, format='eular'):
    """utils.py:168-199 / utils_lr.py:222-256.

    -> (out_img, src_pixel_coords, wmask, src_depth, pose_mat); API v1 is the first three.
    format 'matrix' takes pose as [B,4,4].
    """
    B, H, W, _ = img.shape
    dt = depth.dtype
    if format in ('eular', 'angleaxis'):
        pose = pose_vec2mat(pose, format)
    cam = pixel2cam(depth, meshgrid(B, H, W, dtype=dt), intrinsics)
    K4 = torch.zeros(B, 4, 4, dtype=dt)
    K4[:, :3, :3] = intrinsics
    K4[:, 3, 3] = 1.0
    proj = _mm(K4, pose)
    coords, z = cam2pixel(cam, proj)
    out, wmask = bilinear_sampler(img, coords)
    return out, coords, wmask, z, pose


def optflow_warp(img, flowx, flowy):
    """utils.py:201-217: coords = grid + flow -> sampler; image only."""
    B, H, W, _ = img.shape
    g = meshgrid(B, H, W, is_homogeneous=False, dtype=img.dtype).permute(0, 2, 3, 1)
    coords = torch.cat([g[..., 0:1] + flowx, g[..., 1:2] + flowy], dim=3)
    return bilinear_sampler(img, coords)[0]


def depth_optflow(src_pixel_coords):
    """utils.py:321-338: flow = coords - grid."""
    B, H, W, _ = src_pixel_coords.shape
    g = meshgrid(B, H, W, is_homogeneous=False, dtype=src_pixel_coords.dtype).permute(0, 2, 3, 1)
    return src_pixel_coords[..., 0:1] - g[..., 0:1], src_pixel_coords[..., 1:2] - g[..., 1:2]


def consistent_depth_loss(src_depth, pred_src_depth, coords):
    """utils_lr.py:369-458: |pred_src_depth - bilinear(src_depth, coords)|, no reduction."""
    return torch.abs(pred_src_depth - bilinear_sampler(src_depth, coords)[0])


# ----------------------------------------------------------------------------
# loss terms  (my_losses.py:14-43, pyramid = tf.image.resize_area)
# ----------------------------------------------------------------------------
def compute_smooth_loss(pred_disp):
    """my_losses.py:27-36: four separate means of |second differences|."""
    def grad(p):
        return p[:, :, 1:, :] - p[:, :, :-1, :], p[:, 1:, :, :] - p[:, :-1, :, :]
    dx, dy = grad(pred_disp)
    dx2, dxdy = grad(dx)
    dydx, dy2 = grad(dy)
    return dx2.abs().mean() + dxdy.abs().mean() + dydx.abs().mean() + dy2.abs().mean()


def compute_exp_reg_loss(pred, ref):
    """my_losses.py:39-43: mean softmax cross-entropy of 2-channel logits against ref."""
    l = -(ref.reshape(-1, 2) * torch.log_softmax(pred.reshape(-1, 2), dim=-1)).sum(-1)
    return l.mean()


def get_reference_explain_mask(downscaling, batch, height, width, dtype=torch.float32):
    """my_losses.py:14-23: constant [0,1] labels at scale `downscaling`."""
    m = torch.zeros(batch, int(height / 2 ** downscaling), int(width / 2 ** downscaling), 2, dtype=dtype)
    m[..., 1] = 1.0
    return m


def images_from_uint8(u8, img_format='u8_255'):
    """The loaders' conversion of decoded uint8 frames to the float32 images the graph sees:
    'u8_255'          tf.to_float(image) / 255.0          imageselect_Dataloader.py:86-93
    'u8_255_centred'  image_seq / 255.0 - 0.5             imageselect_Dataloader_optflow_dim11.py:128
    'u8_raw'          tf.to_float only (normalisation commented out)   imageselect_Dataloader_optflow.py:129
    float32 IEEE division and subtraction, as TF's RealDiv / Sub kernels."""
    x = u8.to(torch.float32)
    if img_format == 'u8_raw':
        return x
    x = x / torch.tensor(255.0, dtype=torch.float32)
    return x - torch.tensor(0.5, dtype=torch.float32) if img_format == 'u8_255_centred' else x


def resize_area(x, oh, ow):
    """tf.image.resize_area for integer shrink factors (e.g. train_depth_then_cam_lr.py:227-232).

    Order of TF's ResizeArea kernel (ComputePatchSum) when every overlap weight is 1: per contributing row the
    fx values are summed left to right, the fy row sums are accumulated top to bottom, then * 1/(fy*fx)."""
    B, H, W, C = x.shape
    assert H % oh == 0 and W % ow == 0
    fy, fx = H // oh, W // ow
    blk = x.reshape(B, oh, fy, ow, fx, C)
    total = None
    for dy in range(fy):
        row = blk[:, :, dy, :, 0, :]
        for dx in range(1, fx):
            row = row + blk[:, :, dy, :, dx, :]
        total = row if total is None else total + row
    return total * (torch.tensor(1.0, dtype=x.dtype) / torch.tensor(float(fy * fx), dtype=x.dtype))


# ----------------------------------------------------------------------------
# the multi-scale composition the fused CUDA entry implements
# ----------------------------------------------------------------------------
class LossFlags(object):
    """Attribute bag in the spirit of the reference's FLAGS (train_depth_then_cam_lr.py:44-54)."""

    def __init__(self, **kw):
        self.num_scales = 4
        self.smooth_weight = 0.5
        self.data_weight = 1.0
        self.explain_reg_weight = 0.2
        self.pose_format = 'eular'          # 'eular' | 'angleaxis' | 'matrix'
        self.pixel_scale_norm = True        # data_weight/2^s (train.py:135) vs not (train_depth_then_cam_lr.py:310)
        self.depth_is_inverse = True        # warp depth = 1/x (train.py:128) vs x
        self.smooth_on_inverse = False      # smooth(1/x) (train_depth_then_cam_lr.py:217) vs smooth(x) (train.py:108)
        self.consist_weight = 0.0           # FLAGS.depth_weight on the consistency term (train_depth_then_cam_lr.py:339-340)
        # EXTENSION, not in the reference ("parity unpinned -- no reference implementation"): a in (0, 1] mixes
        # the 3x3 SSIM dissimilarity into the photometric term: dw * [(1 - a) mean(|e| m) + a mean(D m_centre)]
        self.ssim_weight = 0.0
        self.__dict__.update(kw)


def view_synthesis_loss(tgt, srcs, x_pyr, poses, K_pyr, logits_pyr=None, mask_pyr=None, flags=None, src_x_pyr=None):
    """Per-scale loop of train.py:107-135 with the exp-mask of train_depth_then_cam_lr.py:297-328.

    tgt [B,H,W,3]; srcs: list of V [B,H,W,3]; x_pyr: list of S network outputs [B,Hs,Ws,1];
    poses [B,V,6] or [B,V,4,4]; K_pyr [B,S,3,3]; logits_pyr: list of S [B,Hs,Ws,2V] or None;
    mask_pyr: list of S constant weights [B,Hs,Ws,1] or None (train_optflow_combine.py:176,187).
    src_x_pyr (with flags.consist_weight > 0): per source view its own S network outputs; adds the left-right
    depth-consistency term of train_depth_then_cam_lr.py:336-340 (consistent_depth_loss on the warp's projected depth
    and coordinates, weighted by the same mask as the photometric error).
    -> (pixel_loss, smooth_loss, exp_loss) [+ (consist_loss,) with src_x_pyr]
    """
    f = flags or LossFlags()
    consist = None
    B, H, W, _ = tgt.shape
    zero = torch.zeros((), dtype=tgt.dtype)
    pixel, smooth, exp = zero, zero, zero
    for s in range(f.num_scales):
        hs, ws = int(H / 2 ** s), int(W / 2 ** s)
        x = x_pyr[s]
        smooth = smooth + f.smooth_weight / (2 ** s) * compute_smooth_loss(1.0 / x if f.smooth_on_inverse else x)
        tgt_s = resize_area(tgt, hs, ws)
        depth = (1.0 / x if f.depth_is_inverse else x).squeeze(3)
        dw = f.data_weight / (2 ** s) if f.pixel_scale_norm else f.data_weight
        for v, src in enumerate(srcs):
            src_s = resize_area(src, hs, ws)
            warped, coords, _, z_u, _ = projective_inverse_warp(src_s, depth, poses[:, v], K_pyr[:, s], f.pose_format)
            err = torch.abs(warped - tgt_s)
            cerr = None
            if src_x_pyr is not None:
                sx = src_x_pyr[v][s]
                cerr = consistent_depth_loss(1.0 / sx if f.depth_is_inverse else sx, z_u, coords)
            if logits_pyr is not None:
                lg = logits_pyr[s][..., 2 * v:2 * v + 2]
                if f.explain_reg_weight > 0:
                    ref = get_reference_explain_mask(s, B, H, W, tgt.dtype)
                    exp = exp + f.explain_reg_weight * compute_exp_reg_loss(lg, ref)
                err = err * torch.softmax(lg, dim=-1)[..., 1:2]
                if cerr is not None:
                    cerr = cerr * torch.softmax(lg, dim=-1)[..., 1:2]
            elif mask_pyr is not None:
                err = err * mask_pyr[s]
                if cerr is not None:
                    cerr = cerr * mask_pyr[s]
            a = float(getattr(f, 'ssim_weight', 0.0))
            if a > 0.0:
                D = ssim_dissimilarity(warped, tgt_s)                       # [B,hs-2,ws-2,3], VALID windows
                if logits_pyr is not None:
                    D = D * torch.softmax(lg, dim=-1)[:, 1:-1, 1:-1, 1:2]   # the mask at the window's centre
                elif mask_pyr is not None:
                    D = D * mask_pyr[s][:, 1:-1, 1:-1]
                pixel = pixel + ((1.0 - a) * err.mean() + a * D.mean()) * dw
            else:
                pixel = pixel + err.mean() * dw
            if cerr is not None:
                consist = cerr.mean() * f.consist_weight + (consist if consist is not None else 0.0)
    if src_x_pyr is not None:
        return pixel, smooth, exp, consist
    return pixel, smooth, exp


class FlowLossFlags(object):
    """The FLAGS train_optflow_combine.py reads in its loss loop (:138-210)."""

    def __init__(self, **kw):
        self.num_scales = 4
        self.smooth_weight = 0.5
        self.depth_weight = 1.0
        self.data_weight = 1.0
        self.optflow_weight = 1.0
        self.__dict__.update(kw)


def flow_depth_loss(image_left, image_right, label, pred_depth, pred_optflow_x, pred_optflow_y, proj, K_pyr,
                    flags=None):
    """Per-scale loop of train_optflow_combine.py:138-210 (the DeMoN-pair family, BASELINE configs[3]).

    image_left / image_right [B,H,W,3]; label [B,H,W,1] ground-truth INVERSE depth (the warp uses 1/label, :171);
    pred_depth / pred_optflow_x / pred_optflow_y: lists of S network outputs [B,Hs,Ws,1] (inverse depth, flow in
    pixels of that scale); proj [B,4,4] the loader's target-to-source transform (tgt2src_projs[:,0], :173);
    K_pyr [B,S,3,3].  Per scale: second-order smoothness of the three maps; |label_s - pred_depth|; the right image
    warped by the PREDICTED depth and by the PREDICTED flow, both against the left image and both weighted by the
    validity mask of the GROUND-TRUTH-depth warp (three channels, no gradient: the label is data); and
    |pred_flow - depth_optflow(coords of the ground-truth warp)|.  All weights / 2^s.
    -> (depth_loss, smooth_loss, optflow_loss, pixel_loss)   [total_loss = their sum, :240]
    """
    f = flags or FlowLossFlags()
    B, H, W, _ = image_left.shape
    zero = torch.zeros((), dtype=image_left.dtype)
    depth_loss, smooth, smooth_x, smooth_y, optflow, pixel = zero, zero, zero, zero, zero, zero
    for s in range(f.num_scales):
        k = 1.0 / (2 ** s)
        smooth = smooth + f.smooth_weight / (2 ** s) * compute_smooth_loss(pred_depth[s])
        smooth_x = smooth_x + f.smooth_weight / (2 ** s) * compute_smooth_loss(pred_optflow_x[s])
        smooth_y = smooth_y + f.smooth_weight / (2 ** s) * compute_smooth_loss(pred_optflow_y[s])
        hs, ws = int(H / 2 ** s), int(W / 2 ** s)
        lab = resize_area(label, hs, ws)
        left = resize_area(image_left, hs, ws)
        right = resize_area(image_right, hs, ws)
        depth_loss = depth_loss + torch.abs(lab - pred_depth[s]).mean() * f.depth_weight * k
        _, coords_gt, wmask = projective_inverse_warp(right, (1.0 / lab).squeeze(3), proj, K_pyr[:, s], 'matrix')[:3]
        wmask = torch.cat([wmask, wmask, wmask], dim=3)
        warped = projective_inverse_warp(right, (1.0 / pred_depth[s]).squeeze(3), proj, K_pyr[:, s], 'matrix')[0]
        pixel = pixel + (torch.abs(warped - left) * wmask).mean() * f.data_weight * k
        flowed = optflow_warp(right, pred_optflow_x[s], pred_optflow_y[s])
        pixel = pixel + (torch.abs(flowed - left) * wmask).mean() * f.data_weight * k
        gx, gy = depth_optflow(coords_gt)
        optflow = optflow + torch.abs(pred_optflow_x[s] - gx).mean() * f.optflow_weight * k
        optflow = optflow + torch.abs(pred_optflow_y[s] - gy).mean() * f.optflow_weight * k
    return depth_loss, (smooth + smooth_x) + smooth_y, optflow, pixel      # :238: the three sums, then added


def resize_bilinear_tf1(img, out_h, out_w):
    """tf.image.resize_images(img, [out_h, out_w]) as the loaders call it (bilinear; TF 1.x ResizeBilinear, an
    un-vendored and unpinned third party: kernels/resize_bilinear_op.cc with align_corners = False and no half-pixel
    centres -- PARITY UNPINNED, a restatement of the published algorithm).  img: [B,h,w,C] (any real dtype) -> float32.
    in = out_index * (in_size / float(out_size)); lower = floor(in), upper = min(ceil(in), in_size - 1),
    lerp = in - lower; top = tl + (tr - tl) * x_lerp; bottom likewise; out = top + (bottom - top) * y_lerp."""
    B, h, w, C = img.shape
    f = img.to(torch.float32)

    def axis(n_in, n_out):
        scale = torch.tensor(n_in, dtype=torch.float32) / torch.tensor(n_out, dtype=torch.float32)
        pos = torch.arange(n_out, dtype=torch.float32) * scale
        lo = torch.floor(pos)
        hi = torch.clamp(torch.ceil(pos), max=float(n_in - 1))
        return lo.long(), hi.long(), pos - lo
    y0, y1, ly = axis(h, out_h)
    x0, x1, lx = axis(w, out_w)
    lx = lx.reshape(1, 1, out_w, 1)
    ly = ly.reshape(1, out_h, 1, 1)
    rows0, rows1 = f[:, y0], f[:, y1]
    top = rows0[:, :, x0] + (rows0[:, :, x1] - rows0[:, :, x0]) * lx
    bot = rows1[:, :, x0] + (rows1[:, :, x1] - rows1[:, :, x0]) * lx
    return top + (bot - top) * ly


def unpack_strip(strip_u8, H, W):
    """imageselect_Dataloader_optflow.py:120-133, :218-236: resize the decoded two-frame strip to [H, 2 W], to_float,
    target = columns [0, W), source = columns [W, 2 W)."""
    seq = resize_bilinear_tf1(strip_u8, H, 2 * W)
    return seq[:, :, :W].contiguous(), seq[:, :, W:].contiguous()


# ----------------------------------------------------------------------------
# flagged-off extensions named by BASELINE.json but ABSENT from the reference:
# "parity unpinned -- no reference implementation" 
# ----------------------------------------------------------------------------
def ssim_dissimilarity(x, y):
    """3x3 VALID avg-pool SSIM, C1=0.01^2, C2=0.03^2 -> clip((1-SSIM)/2, 0, 1), [B,H-2,W-2,C]."""
    def pool(t):
        return torch.nn.functional.avg_pool2d(t.permute(0, 3, 1, 2), 3, 1).permute(0, 2, 3, 1)
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    mx, my = pool(x), pool(y)
    sx, sy, sxy = pool(x * x) - mx * mx, pool(y * y) - my * my, pool(x * y) - mx * my
    ssim = ((2 * mx * my + C1) * (2 * sxy + C2)) / ((mx * mx + my * my + C1) * (sx + sy + C2))
    return torch.clamp((1 - ssim) / 2, 0, 1)


def edge_aware_smooth_loss(disp, img):
    """First-order disparity gradients weighted by exp(-mean_c |dI|)."""
    ddx = disp[:, :, 1:] - disp[:, :, :-1]
    ddy = disp[:, 1:] - disp[:, :-1]
    wx = torch.exp(-(img[:, :, 1:] - img[:, :, :-1]).abs().mean(3, keepdim=True))
    wy = torch.exp(-(img[:, 1:] - img[:, :-1]).abs().mean(3, keepdim=True))
    return (ddx.abs() * wx).mean() + (ddy.abs() * wy).mean()


# ----------------------------------------------------------------------------
# the step after the path: tf.train.AdamOptimizer(lr, beta1), train_depth_then_cam_lr.py:413.
# TensorFlow is not vendored by the reference and unpinned; this restates the algorithm TF 1.x documents for
# AdamOptimizer / ApplyAdam (training/adam.py docstring).  Parity unpinned (no reference fixture exists).
# ----------------------------------------------------------------------------
def adam_step_tf(param, grad, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """-> (param, m, v) after step t >= 1:  lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);
    m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  param -= lr_t m / (sqrt(v) + eps)."""
    # TF casts the hyper-parameters to the variable dtype (float32) before use: 1 - float32(0.999) is
    # 9.99987e-4, not 1e-3 -- a 1.3e-5 relative difference in v that belongs to the reference's behaviour
    import numpy as _np
    lr, beta1, beta2, eps = (float(_np.float32(h)) for h in (lr, beta1, beta2, eps))
    lr_t = lr * (1.0 - beta2 ** t) ** 0.5 / (1.0 - beta1 ** t)
    m = beta1 * m + (1.0 - beta1) * grad
    v = beta2 * v + (1.0 - beta2) * grad * grad
    return param - lr_t * m / (v.sqrt() + eps), m, v
