"""2-D slice views and the MIP preview (GLRenderer::Render2DView, GLRenderer.cpp:1057-1176).  The host set-up
(tvkr_get_slice_params) against a numpy restatement of Render2DView's formulas on the CPU; the kernel against the numpy
restatement of RenderSlice + the slice / MIP shaders on the GPU."""
import numpy as np
import pytest

from oracle import host as H
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _renderer(raw, scale=(1, 1, 1), size=(200, 120)):
    from imagevis3d_b200.renderer import Renderer
    r = Renderer(0)
    r.loadRawVolume(raw, scale=scale, brick_size=32, overlap=2, hist2d=True)
    r.resize(size)
    return r


def _single_brick(raw, brick=32):
    oc = O.build_octree(raw, brick, 2)
    p = oc["plan"]
    idx = 0
    for l in range(int(p.lod_count)):
        n = int(p.lod_bricks[l][0]) * int(p.lod_bricks[l][1]) * int(p.lod_bricks[l][2])
        if n == 1:
            return oc, l, O.brick(oc, idx)
        idx += n
    raise AssertionError


CASES = [(np.uint8, (70, 45, 38), (1, 1, 1)), (np.uint16, (50, 90, 41), (1, 0.5, 2.0))]


@pytest.mark.parametrize("dtype,vol,scale", CASES)
def test_slice_and_mip_views_match_the_restated_shaders(dtype, vol, scale):
    raw = H.synth_blobs(vol, dtype)
    oc, lod, brick = _single_brick(raw)
    r = _renderer(raw, scale)
    tf = H.tf1d_std(256 if dtype == np.uint8 else 4096, 0.4, 0.5)
    tf[:, 3] *= 0.7
    tfb = H.tf_bytes(tf)
    r.set1DTrans(tfb)
    mm = float(oc["minmax"][:, 1].max())
    w, h = 200, 120
    for wm in (0, 1, 2):
        for mip in (False, True):
            for flip in ((False, False), (True, False), (True, True)):
                si = vol[wm] // 2
                want = H.slice_params(vol, scale, [int(v) for v in oc["plan"].lod_size[lod]], brick.shape[::-1], raw.dtype.itemsize, w, h,
                                      wm, si, mip, flip, 0, mm, 0.0)
                got = r.sliceParams(wm, si, mip, flip)
                assert list(got.voxels) == want["voxels"] and got.use_mip == want["use_mip"] and got.window_mode == wm
                assert np.float32(got.slice) == want["slice"] and np.float32(got.trans_scale) == want["trans_scale"]
                assert [np.float32(v) for v in got.min_coords] == want["min_coords"] and [np.float32(v) for v in got.max_coords] == want["max_coords"]
                assert [np.float32(v) for v in got.quad] == want["quad"]
                r.render2DView(wm, si, mip, flip)
                img = r.frameFloat()
                ref = H.slice_view(want, brick, tfb)
                err = np.abs(img - ref)
                # GL_NEAREST table look-ups: a sample within an ulp of a table boundary may pick the neighbouring entry
                assert (err > 1e-6).mean() < 2e-3 and err.max() <= 0.1, (wm, mip, flip, err.max(), (err > 1e-6).mean())
                assert (img[..., 3] == 1.0).all() and img[..., :3].max() > 0.05
    # after a 2-D view the 3-D path starts over
    r.paintUntilConverged(100)
    assert (r.frameFloat()[..., 3] < 1.0).any()
    r.cleanup()


def test_slice_view_with_the_2d_transfer_function():
    raw = H.synth_blobs((64, 64, 48), np.uint8)
    oc, lod, brick = _single_brick(raw)
    r = _renderer(raw, size=(160, 160))
    t2 = np.zeros((256, 256, 4), np.uint8)
    t2[..., 0] = np.arange(256)[None, :]
    t2[..., 1] = np.arange(256)[:, None]
    t2[..., 2] = 90
    t2[..., 3] = 200
    r.set2DTrans(t2, (0.0, 255.0, 0.0, 255.0))
    r.setRenderMode(1)
    h1, mg = O.hist2d(oc, 256, oc["minmax"][:, 1].max())
    want = H.slice_params((64, 64, 48), (1, 1, 1), [int(v) for v in oc["plan"].lod_size[lod]], brick.shape[::-1], 1, 160, 160, 2, 20, False,
                          (False, False), 1, float(oc["minmax"][:, 1].max()), mg)
    got = r.sliceParams(2, 20)
    assert got.render_mode == 1 and np.float32(got.gradient_scale) == want["gradient_scale"]
    r.render2DView(2, 20)
    img = r.frameFloat()
    ref = H.slice_view(want, brick, t2)
    err = np.abs(img - ref)
    assert (err > 1e-6).mean() < 5e-3 and err.max() <= 0.1
    # MIP in 2-D mode classifies with the 1-D function and leaves the 2-D table in place for the next 3-D frame
    r.render2DView(2, 20, mip=True)
    assert r.sliceParams(2, 20, mip=True).render_mode == 0
    r.paintUntilConverged(100)
    assert (r.frameFloat()[..., 3] > 0).any()
    with pytest.raises(Exception):
        r.render2DView(2, 48)                                  # beyond the domain
    r.cleanup()
