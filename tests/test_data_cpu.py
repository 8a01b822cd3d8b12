"""The DeMoN-pair loader's on-disk formats (imageselect_Dataloader_optflow.py:64-183) on the host: write -> read round
trip of every file type, the file-name scheme, the multi-scale intrinsics, and the oracle's TF1 bilinear resize on
hand-checkable cases.  No GPU needed."""
import numpy as np
import pytest
import torch

from oracle import vsl_oracle as O
from tf_depth_estimation_b200 import data


def test_dataset_round_trip(tmp_path):
    pytest.importorskip('PIL')
    h, w = 24, 32
    written = data.write_synthetic_dataset(str(tmp_path), 3, h, w, seed=5)
    ds = data.PairDataset(str(tmp_path), h, w, num_scales=3, resizedheight=16, resizedwidth=24)
    assert len(ds) == 3
    assert ds.samples[1]['image'].endswith('seq0/0001_0002.jpg')                       # :83-84
    assert ds.samples[1]['depth'].endswith('seq0/frame0001_0002.jpg_z.bin')            # :87-88
    assert ds.samples[1]['cam'].endswith('seq0/0001_0002_cam.txt') and ds.samples[1]['proj'].endswith('_tgt2src_proj.txt')
    for i, wr in enumerate(written):
        got = ds.read(i)
        assert got['strip'].shape == (h, 2 * w, 3) and got['strip'].dtype == np.uint8
        assert np.abs(got['strip'].astype(np.int32) - wr['strip'].astype(np.int32)).mean() < 4.0   # JPEG is lossy
        assert np.array_equal(got['label'], wr['label'])                                # raw float32: exact
        assert np.array_equal(got['K'], wr['K']) and np.array_equal(got['projs'], wr['projs'])
        assert got['m_scale'] == np.float64(wr['m_scale'])
    batches = list(ds.batches(2, seed=1))
    assert len(batches) == 1 and len(batches[0]) == 2


def test_multi_scale_intrinsics_with_resize_ratio():
    K = torch.tensor([[[100.0, 0, 50.0], [0, 120.0, 40.0], [0, 0, 1.0]]])
    Kp = data.multi_scale_intrinsics(K, 3, 0.5, 0.25)
    assert tuple(Kp.shape) == (1, 3, 3, 3)
    assert Kp[0, 0].tolist() == [[50.0, 0, 25.0], [0, 30.0, 10.0], [0, 0, 1.0]]
    assert Kp[0, 2].tolist() == [[12.5, 0, 6.25], [0, 7.5, 2.5], [0, 0, 1.0]]


def test_tf1_bilinear_resize_known_answers():
    x = torch.arange(8, dtype=torch.float32).reshape(1, 2, 4, 1)        # rows [0 1 2 3], [4 5 6 7]
    assert torch.equal(O.resize_bilinear_tf1(x, 2, 4), x)               # identity
    up = O.resize_bilinear_tf1(x, 2, 8)                                 # scale 0.5: 0, .5, 1, ..., 3, then the edge repeats
    assert up[0, 0, :, 0].tolist() == [0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.0]
    down = O.resize_bilinear_tf1(x, 1, 2)                               # scale 2: samples columns 0 and 2 of row 0
    assert down[0, 0, :, 0].tolist() == [0.0, 2.0]
    u8 = torch.tensor([[[[0, 10, 20]], [[255, 30, 40]]]], dtype=torch.uint8).reshape(1, 2, 1, 3)
    mid = O.resize_bilinear_tf1(u8, 4, 1)                               # rows 0, .5, 1, 1 (clamped)
    assert mid[0, :, 0, 0].tolist() == [0.0, 127.5, 255.0, 255.0]
    tgt, src = O.unpack_strip(torch.arange(2 * 4 * 3, dtype=torch.uint8).reshape(1, 2, 4, 3), 2, 2)
    assert tgt.shape == (1, 2, 2, 3) and src[0, 0, 0].tolist() == [6.0, 7.0, 8.0]   # source = the strip's right half


def test_unpack_strip_has_no_cpu_fallback():
    with pytest.raises(TypeError):
        data.unpack_strip(torch.zeros(1, 4, 8, 3, dtype=torch.uint8), 4, 4)
