"""Generates tests/golden/flow_golden.npz by EXECUTING THE REFERENCE'S OWN SOURCE.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_flow.py

The flow-and-depth loss of train_optflow_combine.py:138-240 (BASELINE configs[3], the DeMoN-pair family) --
second-order smoothness of the predicted inverse depth and of both flow channels, |label - pred_depth|, the right
image warped by the predicted depth (projective_inverse_warp) and by the predicted flow (optflow_warp), both against
the left image under the validity mask of the GROUND-TRUTH-depth warp, and |pred_flow - depth_optflow(coords of the
ground-truth warp)| -- re-assembled from the reference's functions (utils.py imported unmodified over the torch-backed
TF1 shim; compute_smooth_loss compiled out of my_losses.py) in the order the script applies them.  The script hands
the loader's 4x4 matrix tgt2src_projs[:,0] (:173) to the warp; utils.py's current projective_inverse_warp starts with
pose_vec2mat (6-vector), so -- as the script's own generation of utils.py did -- the warp below is that function's
body from its second statement on (utils.py:185-199), taking the matrix as is.

Stored: inputs, the four loss terms and the gradients of their sum w.r.t. the three prediction pyramids, in fp32
and fp64 (shim float switch).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tf_depth_estimation_b200 import synth  # noqa: E402

assert ref_loader.available(), 'needs /root/reference'
ref = ref_loader.load_module('utils.py', 'ref_utils_v1')
import tensorflow as tf  # noqa: E402  (the shim)

losses = ref_loader.load_functions('my_losses.py', ['compute_smooth_loss'])

OUT = {}
FLAGS = dict(num_scales=4, smooth_weight=0.5, depth_weight=2.0, data_weight=3.0, optflow_weight=0.7)


def put(**arrays):
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().numpy()
        OUT['flow/%s' % k] = np.ascontiguousarray(v)


def T(x, dtype, grad=False):
    t = x.detach().clone().to(dtype).as_subclass(tf.Tensor)
    t.requires_grad_(grad)
    return t


def plain(t):
    return t.detach().as_subclass(torch.Tensor)


def warp_with_matrix(img, depth, pose, intrinsics):
    """utils.py:185-199: projective_inverse_warp after its pose_vec2mat line, calling the reference's own
    meshgrid / pixel2cam / cam2pixel / bilinear_sampler."""
    batch, height, width, _ = img.get_shape().as_list()
    pixel_coords = ref.meshgrid(batch, height, width)
    cam_coords = ref.pixel2cam(depth, pixel_coords, intrinsics)
    filler = tf.constant([0.0, 0.0, 0.0, 1.0], shape=[1, 1, 4])
    filler = tf.tile(filler, [batch, 1, 1])
    intrinsics = tf.concat([intrinsics, tf.zeros([batch, 3, 1])], axis=2)
    intrinsics = tf.concat([intrinsics, filler], axis=1)
    proj_tgt_cam_to_src_pixel = tf.matmul(intrinsics, pose)
    src_pixel_coords = ref.cam2pixel(cam_coords, proj_tgt_cam_to_src_pixel)
    output_img, wmask = ref.bilinear_sampler(img, src_pixel_coords)
    return output_img, src_pixel_coords, wmask


def flow_loss(image_left, image_right, label, pred_depth, pred_optflow_x, pred_optflow_y, tgt2src, intrinsics):
    """train_optflow_combine.py:138-240, names as in the script."""
    B, H, W, _ = image_left.shape
    depth_loss = optflow_loss = pixel_loss = smooth_loss = smooth_loss_optx = smooth_loss_opty = 0
    for s in range(FLAGS['num_scales']):
        smooth_loss += FLAGS['smooth_weight'] / (2 ** s) * losses['compute_smooth_loss'](pred_depth[s])
        smooth_loss_optx += FLAGS['smooth_weight'] / (2 ** s) * losses['compute_smooth_loss'](pred_optflow_x[s])
        smooth_loss_opty += FLAGS['smooth_weight'] / (2 ** s) * losses['compute_smooth_loss'](pred_optflow_y[s])
        hs, ws = int(H / (2 ** s)), int(W / (2 ** s))
        curr_label = tf.image.resize_area(label, [hs, ws])
        curr_image_left = tf.image.resize_area(image_left, [hs, ws])
        curr_image_right = tf.image.resize_area(image_right, [hs, ws])
        curr_depth_error = tf.abs(curr_label - pred_depth[s])
        depth_loss += tf.reduce_mean(curr_depth_error) * FLAGS['depth_weight'] / (2 ** s)
        _, src_pixel_coords_gt, wmask = warp_with_matrix(
            curr_image_right, tf.squeeze(1.0 / curr_label, axis=3), tgt2src, intrinsics[:, s, :, :])
        wmask = tf.concat([wmask, wmask, wmask], axis=3)
        curr_proj_image_depth, _, _ = warp_with_matrix(
            curr_image_right, tf.squeeze(1.0 / pred_depth[s], axis=3), tgt2src, intrinsics[:, s, :, :])
        curr_proj_error_depth = tf.multiply(tf.abs(curr_proj_image_depth - curr_image_left), wmask)
        pixel_loss += tf.reduce_mean(curr_proj_error_depth) * FLAGS['data_weight'] / (2 ** s)
        curr_proj_image_optflow = ref.optflow_warp(curr_image_right, pred_optflow_x[s], pred_optflow_y[s])
        curr_proj_error_optflow = tf.multiply(tf.abs(curr_proj_image_optflow - curr_image_left), wmask)
        pixel_loss += tf.reduce_mean(curr_proj_error_optflow) * FLAGS['data_weight'] / (2 ** s)
        depth_optflow_x, depth_optflow_y = ref.depth_optflow(src_pixel_coords_gt)
        curr_optflow_error_x = tf.abs(pred_optflow_x[s] - depth_optflow_x)
        optflow_loss += tf.reduce_mean(curr_optflow_error_x) * FLAGS['optflow_weight'] / (2 ** s)
        curr_optflow_error_y = tf.abs(pred_optflow_y[s] - depth_optflow_y)
        optflow_loss += tf.reduce_mean(curr_optflow_error_y) * FLAGS['optflow_weight'] / (2 ** s)
    smooth_loss = smooth_loss + smooth_loss_optx + smooth_loss_opty
    return depth_loss, smooth_loss, optflow_loss, pixel_loss


def main():
    S = FLAGS['num_scales']
    d = synth.make_flow_pairs(2, 32, 48, S=S, seed=77, motion=2.0)
    put(left=d['left'], right=d['right'], label=d['label'], proj=d['proj'], K_pyr=d['K_pyr'],
        **{'pred_depth%d' % s: x for s, x in enumerate(d['depth_pyr'])},
        **{'pred_flowx%d' % s: x for s, x in enumerate(d['flowx_pyr'])},
        **{'pred_flowy%d' % s: x for s, x in enumerate(d['flowy_pyr'])})
    OUT['flow/flags'] = np.array(repr(FLAGS))
    for dt, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
        tf.set_float(dt)
        pd = [T(x, dt, True) for x in d['depth_pyr']]
        fx = [T(x, dt, True) for x in d['flowx_pyr']]
        fy = [T(x, dt, True) for x in d['flowy_pyr']]
        terms = flow_loss(T(d['left'], dt), T(d['right'], dt), T(d['label'], dt), pd, fx, fy, T(d['proj'], dt),
                          T(d['K_pyr'], dt))
        grads = torch.autograd.grad(sum(terms), pd + fx + fy)
        put(**{'%s_%s' % (k, tag): plain(v) for k, v in zip(('depth', 'smooth', 'optflow', 'pixel'), terms)})
        put(**{'g_pred_depth%d_%s' % (s, tag): plain(grads[s]) for s in range(S)})
        put(**{'g_pred_flowx%d_%s' % (s, tag): plain(grads[S + s]) for s in range(S)})
        put(**{'g_pred_flowy%d_%s' % (s, tag): plain(grads[2 * S + s]) for s in range(S)})
    tf.set_float(torch.float32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'flow_golden.npz')
    np.savez_compressed(path, **OUT)
    print('wrote %s: %d arrays, %.1f KiB' % (path, len(OUT), os.path.getsize(path) / 1024.0))


if __name__ == '__main__':
    torch.set_num_threads(1)
    main()
