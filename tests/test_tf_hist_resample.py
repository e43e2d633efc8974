"""CPU: TransferFunction1D::Resample (row T1) and the histogram the TF editor gets (row T3), product host code through the C
ABI against the reference -- its compiled TransferFunction1D (oracle/_ref/ref_driver tfresample) and the statements of
UVFDataset::GetHistograms run by oracle/ref_extract.py -- plus cases that need no reference."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from imagevis3d_b200 import _lib as L
from oracle import oracle as O
from oracle import ref_extract as R


def _tf_resample(rgba, target):
    rgba = np.ascontiguousarray(rgba, np.float32)
    out = np.zeros((target, 4), np.float32)
    fn = L.load().tvkr_tf1d_resample
    fn.restype = C.c_int
    assert fn(rgba.ctypes.data_as(C.c_void_p), C.c_uint32(len(rgba)), C.c_uint32(target), out.ctypes.data_as(C.c_void_p)) == 0
    return out


def _hist_resample(bins, target):
    bins = np.ascontiguousarray(bins, np.uint64)
    out = np.zeros(target, np.uint32)
    fn = L.load().tvkr_resample_hist1d
    fn.restype = C.c_int
    assert fn(bins.ctypes.data_as(C.c_void_p), C.c_uint64(len(bins)), C.c_uint32(target), out.ctypes.data_as(C.c_void_p)) == 0
    return out


def test_tf_resample_known_answers():
    ramp = np.stack([np.arange(8, dtype=np.float32)] * 4, axis=1)
    assert np.array_equal(_tf_resample(ramp, 8), ramp)
    assert np.array_equal(_tf_resample(ramp, 4)[:, 0], [0.5, 2.5, 4.5, 6.5])            # buckets of two
    # 7 -> 3: the bucket width adapts to what is left: 7/3 = 2, then 5/2 = 2, then 3/1 = 3
    assert np.array_equal(_tf_resample(ramp[:7], 3)[:, 0], [0.5, 2.5, 5.0])
    up = _tf_resample(ramp[:3], 5)[:, 0]                                                 # fPos = i * 2/5
    assert np.allclose(up, [0.0, 0.4, 0.8, 1.2, 1.6], atol=1e-6)


def test_hist_resample_known_answers():
    h = np.full(8192, 3, np.uint64)
    out = _hist_resample(h, 4096)
    assert (out[:-1] == 6).all() and out[-1] == 0          # the last bin is the one the reference's loop never writes
    assert np.array_equal(_hist_resample(np.arange(10, dtype=np.uint64), 10), np.arange(10))


@pytest.mark.skipif(not os.path.isfile(O.ref_driver_path() or ""), reason="oracle/_ref/ref_driver is not built (/root/reference absent)")
@pytest.mark.parametrize("n,target", [(256, 64), (256, 100), (4096, 256), (100, 256), (7, 3), (64, 4096), (300, 299)])
def test_tf_resample_equals_reference(n, target):
    rng = np.random.default_rng(n * 31 + target)
    rgba = rng.random((n, 4)).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        a, b = os.path.join(d, "in.f32"), os.path.join(d, "out.f32")
        rgba.tofile(a)
        subprocess.check_call([O.ref_driver_path(), "tfresample", a, str(n), str(target), b])
        want = np.fromfile(b, np.float32).reshape(-1, 4)
    got = _tf_resample(rgba, target)
    assert want.shape == got.shape and np.array_equal(got, want)


@pytest.mark.skipif(not R.available(), reason="/root/reference is not mounted")
@pytest.mark.parametrize("n,target", [(65536, 4096), (40000, 4096), (4097, 4096), (5000, 4096), (12, 5), (65535, 4096)])
def test_hist_resample_equals_reference_statements(n, target):
    rng = np.random.default_rng(n)
    bins = (rng.random(n) ** 4 * 1e6).astype(np.uint64)
    assert np.array_equal(_hist_resample(bins, target), R.hist_resample(bins, target))
