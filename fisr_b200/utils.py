"""Host-side helpers with the names and behaviour of the reference's ``utils.py`` (file formats, tiling, colour, PSNR).

These run on the CPU in numpy exactly like the reference's; the arithmetic of the hot path is in libfisr_b200.so.
"""
from __future__ import annotations

import os

import numpy as np


def str2bool(x):                                   # utils.py:8-9
    return x.lower() in ('true')


def check_folder(log_dir):                         # utils.py:12-15
    if not os.path.exists(log_dir):
        os.makedirs(log_dir)
    return log_dir


def _compute_psnr(img_orig, img_out, peak):        # utils.py:23-26
    mse = np.mean(np.square(img_orig - img_out))
    return 10 * np.log10(peak * peak / mse)


FLO_MAGIC = np.float32(202021.25)


def compare_ssim(image_0, image_1, tile_size=7):
    """SSIM as the reference's ``test()`` measures it (FISRnet.py:5,890-891: ``SSIM_PIL.compare_ssim`` on uint8 images).

    SSIM-PIL 1.0.x is not vendored with the reference and not installable here, so this restates its published algorithm
    (parity unpinned): the image is cut into NON-overlapping ``tile_size`` x ``tile_size`` tiles
    (the ragged border is dropped); for every tile and channel
        ssim = (2 m0 m1 + c1) (2 cov + c2) / ((m0^2 + m1^2 + c1) (var0 + var1 + c2)),
    with population statistics (divide by the pixel count), c1 = (0.01 * 255)^2, c2 = (0.03 * 255)^2, and the result is the
    mean over tiles and channels.  Accepts PIL images or uint8 arrays [H, W] / [H, W, C]."""
    a = np.asarray(image_0, dtype=np.float64)
    b = np.asarray(image_1, dtype=np.float64)
    if a.shape != b.shape:
        raise AttributeError('The images do not have the same resolution.')
    if a.ndim == 2:
        a, b = a[..., None], b[..., None]
    t = int(tile_size)
    h, w = a.shape[0] // t * t, a.shape[1] // t * t
    if h == 0 or w == 0:
        raise ValueError('image smaller than one tile')
    tiles = lambda x: x[:h, :w].reshape(h // t, t, w // t, t, -1)
    a, b = tiles(a), tiles(b)
    n = float(t * t)
    s0, s1 = a.sum(axis=(1, 3)), b.sum(axis=(1, 3))
    s00, s11, s01 = (a * a).sum(axis=(1, 3)), (b * b).sum(axis=(1, 3)), (a * b).sum(axis=(1, 3))
    m0, m1 = s0 / n, s1 / n
    var0, var1, cov = s00 / n - m0 * m0, s11 / n - m1 * m1, s01 / n - m0 * m1
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    ssim = (2 * m0 * m1 + c1) * (2 * cov + c2) / ((m0 * m0 + m1 * m1 + c1) * (var0 + var1 + c2))
    return float(ssim.mean())


def read_flo_file_5dim(filename):
    """utils.py:57-74: float32 magic, int32 N, N_seq, h, w, then N*N_seq*h*w*2 float32 -> [N, N_seq, h, w, 2]."""
    with open(filename, 'rb') as f:
        magic = np.fromfile(f, np.float32, count=1)
        if magic.size != 1 or magic[0] != FLO_MAGIC:
            print('Magic number incorrect. Invalid .flo file')
            return None
        N, N_seq, h, w = (int(v) for v in np.fromfile(f, np.int32, count=4))
        print("Reading %d x %d x %d x %d x 2 flow file in .flo format" % (N, N_seq, h, w))
        data = np.fromfile(f, np.float32, count=N * N_seq * h * w * 2)
    return np.resize(data, (N, N_seq, h, w, 2))


def write_flo_file_5dim(flow, filename):
    """Writer of the same format (FISR_tfoptflow/FISR_for_video_pwcnet_predict_from_img_test.py:57-81)."""
    flow = np.ascontiguousarray(flow, dtype=np.float32)
    N, N_seq, h, w, two = flow.shape
    assert two == 2
    with open(filename, 'wb') as f:
        np.array([FLO_MAGIC], np.float32).tofile(f)
        np.array([N, N_seq, h, w], np.int32).tofile(f)
        flow.tofile(f)


def _h5_read(fname, key):
    """``h5py.File(fname, 'r')[key][()]`` (utils.py:31-34,47): h5py when it is installed, else the built-in minimal HDF5 reader
    (fisr_b200/hdf5_min.py -- superblock v0-v3, contiguous / chunked + deflate / shuffle / fletcher32 datasets)."""
    try:
        import h5py
    except ImportError:
        from .hdf5_min import read_dataset
        return read_dataset(fname, key)
    with h5py.File(fname, 'r') as f:
        return f[key][()]


def read_mat_file(data_fname, label_fname, data_name, label_name):
    """utils.py:29-42: training data / label [N, N_seq, C, W, H] uint8 -> float32 /255, [N, N_seq, H, W, C]."""
    def load(fname, key):
        return np.load(fname) if fname.endswith('.npy') else _h5_read(fname, key)
    data = np.array(load(data_fname, data_name), dtype=np.float32) / 255.
    label = np.array(load(label_fname, label_name), dtype=np.float32) / 255.
    return np.swapaxes(data, 2, 4), np.swapaxes(label, 2, 4)


def read_mat_file_warp(data_fname, data_name):
    """utils.py:45-54: warped frames -> float32 /255, [N, N_seq, H, W, C].  A ``.npy`` file holds that layout already
    (values 0..255); a v7.3 ``.mat`` holds the transpose (MATLAB order), like the files ``hdf5storage`` writes."""
    if data_fname.endswith('.npy'):
        return np.array(np.load(data_fname), dtype=np.float32) / 255.
    data = np.array(_h5_read(data_fname, data_name), dtype=np.float32) / 255.
    return np.transpose(data, (4, 3, 2, 1, 0))


def write_mat_file_warp(data_fname, pred, data_name='pred'):
    """``hdf5storage.write({u'pred': pred}, '.', name, matlab_compatible=True)`` (..warp_img_with_flo.py:131-137): a MATLAB v7.3
    file whose dataset is the transpose of ``pred`` [N-1, 2, h, w, 3] float32 (0..255), class 'single'."""
    from .hdf5_min import write_mat73
    write_mat73(data_fname, {data_name: np.ascontiguousarray(np.transpose(np.asarray(pred, dtype=np.float32), (4, 3, 2, 1, 0)))})


def merge_seq_dim(data):                           # utils.py:78-83
    sz = data.shape
    return np.reshape(np.transpose(data, axes=(0, 2, 3, 1, 4)), (sz[0], sz[2], sz[3], sz[1] * sz[4]))


def split_seq_dim(data):                           # utils.py:86-91
    sz = data.shape
    return np.transpose(np.reshape(data, (sz[0], sz[1], sz[2], sz[3] // 3, 3)), axes=(0, 3, 1, 2, 4))


def yuv2rgb_constants():
    """The 3x3 matrix T and the offset of ``YUV2RGB_matlab`` (utils.py:106-110) as 12 float64 values (T row-major, then offset)."""
    Tinv = np.array([[0.00456621, 0., 0.00625893], [0.00456621, -0.00153632, -0.00318811], [0.00456621, 0.00791071, 0.]])
    offset = [[16], [128], [128]]
    T = 255 * Tinv
    offset = 255 * Tinv @ offset
    return np.concatenate([T.ravel(), np.asarray(offset).ravel()])


def YUV2RGB_matlab(yuv):                           # utils.py:106-115
    k = yuv2rgb_constants()
    T, offset = k[:9].reshape(3, 3), k[9:].reshape(3, 1)
    rgb = np.zeros(yuv.shape)
    for p in range(3):
        rgb[:, :, p] = T[p, 0] * yuv[:, :, 0] + T[p, 1] * yuv[:, :, 1] + T[p, 2] * yuv[:, :, 2] - offset[p]
    return np.clip(rgb, 0, 255)
