"""Input pipeline of the DeMoN-pair family: the on-disk formats imageselect_Dataloader_optflow.py reads,
and the loader's per-batch tensor work on the GPU.

What the reference's DataLoader does (imageselect_Dataloader_optflow.py) and where it lives here:
  :64-100  read_labeled_image_list   <dataset_dir>/<split>.txt, lines "subfolder a b" -> per sample
           <sub>/<a_b>.jpg (two frames side by side), <sub>/frame<a_b>.jpg_z.bin (raw float32 inverse depth,
           image_height x image_width), <sub>/<a_b>_cam.txt (9 comma-separated floats), <sub>/<a_b>_tgt2src_proj.txt
           (34 space-separated fields: two 4x4 matrices, m_scale, one ignored)            -> PairDataset (host, numpy)
  :120-133 decode_jpeg -> resize_images([resizedheight, 2 resizedwidth]) -> to_float -> unpack_image_sequence
           (:218-236)                                                                     -> vsl_unpack_strip (CUDA)
  :136-143 decode_raw label, reshape                                                      -> host, as is
  :55-58, :239-262 get_multi_scale_intrinsics with the resize ratios                      -> multi_scale_intrinsics (host)
JPEG decoding itself stays on the host (PIL): it is not on the hot path this repository covers.  Queue runners,
shuffling with num_epochs and the (commented-out) augmentation of the reference are not reproduced; `batches()` is a
plain shuffled epoch iterator.  `write_synthetic_dataset` writes a dataset in these formats from synthetic snippets
(there is no dataset offline): it is what the tests and demos read back.
"""
import os

import numpy as np
import torch

from . import _lib
from ._lib import check


class PairDataset(object):
    """File lists and per-sample readers, field for field what read_images_from_disk (:104-183) consumes."""

    def __init__(self, dataset_dir, image_height, image_width, num_scales=4, split='train', resizedheight=240,
                 resizedwidth=720):
        self.dataset_dir, self.split = dataset_dir, split
        self.image_height, self.image_width = image_height, image_width
        self.num_scales = num_scales
        self.resizedheight, self.resizedwidth = resizedheight, resizedwidth
        with open(os.path.join(dataset_dir, '%s.txt' % split)) as fh:
            frames = [l for l in fh.read().splitlines() if l.strip()]
        self.samples = []
        for line in frames:                                   # :79-93
            sub, a, b = line.split(' ')[:3]
            fid = a + '_' + b
            d = os.path.join(dataset_dir, sub)
            self.samples.append(dict(image=os.path.join(d, fid + '.jpg'), cam=os.path.join(d, fid + '_cam.txt'),
                                     depth=os.path.join(d, 'frame' + fid + '.jpg_z.bin'),
                                     proj=os.path.join(d, fid + '_tgt2src_proj.txt')))

    def __len__(self):
        return len(self.samples)

    def read(self, i):
        """-> dict(strip uint8 [h,w,3], label float32 [image_height,image_width,1], K float32 [3,3],
        projs float32 [2,4,4], m_scale float)."""
        from PIL import Image
        s = self.samples[i]
        strip = np.asarray(Image.open(s['image']).convert('RGB'), dtype=np.uint8)
        label = np.fromfile(s['depth'], dtype=np.float32).reshape(self.image_height, self.image_width, 1)   # :136
        with open(s['cam']) as fh:
            K = np.array([float(v) for v in fh.read().strip().split(',')], dtype=np.float32).reshape(3, 3)   # :156-163
        with open(s['proj']) as fh:
            v = [float(x) for x in fh.read().strip().split(' ')]                                              # :169-181
        if len(v) != 34:
            raise ValueError('%s: 34 fields expected, got %d' % (s['proj'], len(v)))
        v = v[:-1]
        m_scale = v[-1]
        projs = np.array(v[:-1], dtype=np.float32).reshape(2, 4, 4)
        return dict(strip=strip, label=label, K=K, projs=projs, m_scale=m_scale)

    def batches(self, batch_size, seed=0, drop_last=True):
        order = np.random.default_rng(seed).permutation(len(self))
        for lo in range(0, len(order) - (batch_size - 1 if drop_last else 0), batch_size):
            yield [self.read(int(i)) for i in order[lo:lo + batch_size]]


def multi_scale_intrinsics(K, num_scales, x_resize_ratio, y_resize_ratio):
    """get_multi_scale_intrinsics (:239-262): fx, fy, cx, cy / 2^s * resize ratio, float32 -> [B,S,3,3]."""
    K = torch.as_tensor(K, dtype=torch.float32)
    xr, yr = torch.tensor(x_resize_ratio, dtype=torch.float32), torch.tensor(y_resize_ratio, dtype=torch.float32)
    out = torch.zeros(K.shape[0], num_scales, 3, 3)
    for s in range(num_scales):
        out[:, s, 0, 0] = K[:, 0, 0] / (2 ** s) * xr
        out[:, s, 1, 1] = K[:, 1, 1] / (2 ** s) * yr
        out[:, s, 0, 2] = K[:, 0, 2] / (2 ** s) * xr
        out[:, s, 1, 2] = K[:, 1, 2] / (2 ** s) * yr
        out[:, s, 2, 2] = 1.0
    return out


def unpack_strip(strip_u8, H, W, stream=None):
    """uint8 [B,h,w,3] CUDA tensor -> (tgt, src) float32 [B,H,W,3]: resize_images + to_float + unpack_image_sequence on
    the device (vsl_unpack_strip)."""
    if not strip_u8.is_cuda or strip_u8.dtype != torch.uint8 or strip_u8.dim() != 4 or strip_u8.shape[3] != 3:
        raise TypeError('strip must be a CUDA uint8 tensor [B,h,w,3] (this path has no CPU fallback)')
    strip_u8 = strip_u8.contiguous()
    B, h, w, _ = strip_u8.shape
    tgt = torch.empty(B, H, W, 3, device=strip_u8.device)
    src = torch.empty(B, H, W, 3, device=strip_u8.device)
    st = torch.cuda.current_stream(strip_u8.device).cuda_stream if stream is None else stream
    check(_lib.load().vsl_unpack_strip(strip_u8.data_ptr(), B, h, w, H, W, tgt.data_ptr(), src.data_ptr(), st))
    return tgt, src


def load_batch(ds, samples, device):
    """A list of PairDataset.read() dicts -> what load_train_batch (:28-61) returns, on `device`:
    (tgt_image, src_image_stack, label_batch, intrinsics [B,S,3,3], tgt2src_projs [B,2,4,4], m_scale [B]).
    The strips cross PCIe as uint8 (one pinned copy); resize + unpack run on the GPU."""
    strips = torch.from_numpy(np.stack([s['strip'] for s in samples])).pin_memory()
    tgt, src = unpack_strip(strips.to(device, non_blocking=True), ds.resizedheight, ds.resizedwidth)
    label = torch.from_numpy(np.stack([s['label'] for s in samples])).to(device)
    K = torch.from_numpy(np.stack([s['K'] for s in samples]))
    Kp = multi_scale_intrinsics(K, ds.num_scales, ds.resizedwidth / ds.image_width, ds.resizedheight / ds.image_height)
    projs = torch.from_numpy(np.stack([s['projs'] for s in samples])).to(device)
    m_scale = torch.tensor([s['m_scale'] for s in samples], dtype=torch.float32, device=device)
    return tgt, src, label, Kp.to(device), projs, m_scale


def write_synthetic_dataset(root, n, image_height, image_width, seed=0, split='train', quality=95):
    """Writes `n` synthetic pairs in the reference's formats under `root` (one subfolder) -> list of the exact arrays
    written (strip uint8 BEFORE JPEG coding, label, K, projs, m_scale) for round-trip checks."""
    from PIL import Image
    from . import synth
    d = synth.make_flow_pairs(n, image_height, image_width, S=1, seed=seed)
    sub = 'seq0'
    os.makedirs(os.path.join(root, sub), exist_ok=True)
    written, lines = [], []
    g = torch.Generator().manual_seed(seed)
    for i in range(n):
        a, b = '%04d' % i, '%04d' % (i + 1)
        fid = a + '_' + b
        strip = torch.cat([d['left'][i], d['right'][i]], dim=1)
        strip = (strip * 255.0).round().clamp(0, 255).to(torch.uint8).numpy()
        Image.fromarray(strip).save(os.path.join(root, sub, fid + '.jpg'), quality=quality)
        label = d['label'][i].numpy().astype(np.float32)
        label.tofile(os.path.join(root, sub, 'frame' + fid + '.jpg_z.bin'))
        K = d['K'][i].numpy().astype(np.float32)
        with open(os.path.join(root, sub, fid + '_cam.txt'), 'w') as fh:
            fh.write(','.join(repr(float(v)) for v in K.reshape(-1)))
        projs = np.stack([d['proj'][i].numpy(), np.linalg.inv(d['proj'][i].numpy().astype(np.float64)).astype(np.float32)])
        m_scale = float(torch.rand((), generator=g)) + 0.5
        with open(os.path.join(root, sub, fid + '_tgt2src_proj.txt'), 'w') as fh:
            fh.write(' '.join(repr(float(v)) for v in list(projs.reshape(-1)) + [m_scale, 0.0]))
        lines.append('%s %s %s' % (sub, a, b))
        written.append(dict(strip=strip, label=label, K=K, projs=projs.astype(np.float32), m_scale=m_scale))
    with open(os.path.join(root, '%s.txt' % split), 'w') as fh:
        fh.write('\n'.join(lines) + '\n')
    return written
