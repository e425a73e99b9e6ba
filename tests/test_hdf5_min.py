"""CPU: the h5py-free HDF5 reader / writer for MATLAB v7.3 files (reference utils.py:29-54 reads them with h5py,
FISR_for_video_warp_img_with_flo.py:131-137 writes them with hdf5storage)."""
import os

import numpy as np
import pytest

from fisr_b200 import hdf5_min as H
from fisr_b200 import utils


def _matlab_sample():
    import scipy.io
    p = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(p):
        pytest.skip("scipy's MATLAB v7.3 sample file is not installed")
    return p


def test_reads_a_file_written_by_matlab():
    """Golden vector: a v7.3 file MATLAB 7.4 wrote (512-byte user block, superblock v0, old-style root group, version-1 object
    header, data layout v2): ``testdouble = 0:pi/4:2*pi`` stored as a 9 x 1 double, class attribute 'double'."""
    f = H.H5File(_matlab_sample())
    assert f.base == 512 and f.keys() == ["testdouble"]
    a = f["testdouble"]
    assert a.dtype == np.float64 and a.shape == (9, 1)
    assert np.array_equal(a.ravel(), np.arange(9) * (np.pi / 4))
    assert f.attrs("testdouble")["MATLAB_class"] == "double"
    with pytest.raises(KeyError):
        f["nothing"]


def test_contiguous_round_trip_and_matlab_header(tmp_path):
    rng = np.random.default_rng(0)
    arrs = {"pred": rng.standard_normal((3, 5, 4, 2, 2)).astype(np.float32), "LR_data": rng.integers(0, 256, (2, 3, 7), dtype=np.uint8),
            "d": rng.standard_normal((6,)), "i": rng.integers(-1000, 1000, (4, 4)).astype(np.int32)}
    p = str(tmp_path / "x.mat")
    H.write_mat73(p, arrs)
    raw = open(p, "rb").read()
    assert raw[:10] == b"MATLAB 7.3" and raw[124:128] == b"\x00\x02IM" and raw[512:520] == H.SIGNATURE
    f = H.H5File(p)
    assert sorted(f.keys()) == sorted(arrs)
    for k, v in arrs.items():
        got = f[k]
        assert got.dtype == v.dtype and np.array_equal(got, v)
    assert f.attrs("pred")["MATLAB_class"] == "single" and f.attrs("LR_data")["MATLAB_class"] == "uint8"


@pytest.mark.parametrize("compress", [False, True])
def test_chunked_filtered_round_trip(tmp_path, compress):
    """hdf5storage's defaults for big arrays: chunked + shuffle + gzip + fletcher32; ragged edge chunks; a chunk index deeper
    than one B-tree node (btree_k = 2 -> at most 4 entries per node)."""
    rng = np.random.default_rng(1)
    a = (rng.standard_normal((3, 37, 29, 2)) * 40).astype(np.float32)
    a[:, 10:20] = 0                                         # compressible stretch
    p = str(tmp_path / "c.mat")
    H.write_mat73(p, {"pred": a}, chunks=(1, 16, 8, 2), compress=compress, btree_k=2)
    got = H.read_dataset(p, "pred")
    assert got.shape == a.shape and np.array_equal(got, a)
    if compress:
        assert os.path.getsize(p) < a.nbytes * 1.2


def test_fletcher32_known_properties():
    assert H.fletcher32(b"") == 0
    # one 16-bit big-endian word w: sum1 = w, sum2 = w
    assert H.fletcher32(b"\x01\x02") == (0x0102 << 16) | 0x0102
    assert H.fletcher32(b"\x01\x02\x03\x04") == ((0x0102 + 0x0102 + 0x0304) << 16) | (0x0102 + 0x0304)
    big = bytes(range(256)) * 40
    assert H.fletcher32(big) == H.fletcher32(big) and H.fletcher32(big) != H.fletcher32(big[:-2] + b"\0\0")


def test_reference_readers_take_v73_mat_files(tmp_path):
    """utils.read_mat_file / read_mat_file_warp on real .mat paths: the layouts of utils.py:29-54 (MATLAB stores the transpose)."""
    rng = np.random.default_rng(2)
    lr = rng.integers(0, 256, (4, 5, 3, 12, 10), dtype=np.uint8)                 # [N, N_seq, C, W, H] as h5py shows it
    hr = rng.integers(0, 256, (4, 7, 3, 24, 20), dtype=np.uint8)
    H.write_mat73(str(tmp_path / "lr.mat"), {"LR_data": lr})
    H.write_mat73(str(tmp_path / "hr.mat"), {"HR_data": hr})
    data, label = utils.read_mat_file(str(tmp_path / "lr.mat"), str(tmp_path / "hr.mat"), "LR_data", "HR_data")
    assert data.shape == (4, 5, 10, 12, 3) and label.shape == (4, 7, 20, 24, 3) and data.dtype == np.float32
    assert np.array_equal(data, np.swapaxes(lr.astype(np.float32) / 255., 2, 4))
    warp = rng.uniform(0, 255, (3, 2, 16, 20, 3)).astype(np.float32)             # [N-1, 2, h, w, 3] as the warp driver holds it
    path = str(tmp_path / "w_warp.mat")
    utils.write_mat_file_warp(path, warp)
    assert np.array_equal(H.read_dataset(path, "pred"), warp.transpose(4, 3, 2, 1, 0))      # what hdf5storage puts in the file
    got = utils.read_mat_file_warp(path, "pred")
    assert got.shape == warp.shape and np.allclose(got, warp / np.float32(255.))
