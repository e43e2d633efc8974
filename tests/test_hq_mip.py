"""the full-resolution MIP of the capture mode (GLRenderer::Render2DView :1182-1250, AbstrRenderer::PlanHQMIPFrame
:1214-1245, GLRenderer::RenderHQMIPPreLoop :1256-1283, GLRaycaster.cpp:481-529, GLRaycaster-MIP-Rot-FS.glsl, Transfer-MIP-FS.glsl).
The host's plan (LOD, matrices, brick list) against a numpy restatement of those functions; the kernel against the numpy
restatement of the shaders on the same plan; and two properties no restatement is needed for: the MIP of a view along an axis
with the orthographic projection is the maximum of the volume along that axis."""
import numpy as np
import pytest

from oracle import host as H
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _renderer(raw, size, scale=(1, 1, 1)):
    from imagevis3d_b200.renderer import Renderer, RT_CAPTURE
    r = Renderer(0)
    r.loadRawVolume(raw, scale=scale, brick_size=32, overlap=2)
    r.resize(size)
    r.setRendererTarget(RT_CAPTURE)
    return r


CASES = [(np.uint8, (70, 45, 38), (160, 120)), (np.uint16, (50, 90, 41), (120, 150))]


@pytest.mark.parametrize("dtype,vol,size", CASES)
def test_hq_mip_plan_and_picture(dtype, vol, size):
    raw = H.synth_blobs(vol, dtype)
    oc = O.build_octree(raw, 32, 2)
    r = _renderer(raw, size)
    tf = H.tf1d_std(256 if dtype == np.uint8 else 4096, 0.4, 0.5)
    tfb = H.tf_bytes(tf)
    r.set1DTrans(tfb)
    cam = H.Camera()
    w, h = size
    from tests.test_gpu_classic import _device_store
    store = _device_store(r)
    for wm, angle, flip, ortho in [(2, 0.0, (False, False), True), (0, 30.0, (False, False), True), (1, 200.0, (True, False), True),
                                   (2, 45.0, (False, True), False)]:
        p, bricks, lod = r.planHQMIP(wm, angle, flip, ortho)
        assert lod == 0 and len(bricks) == int(np.prod(oc["plan"].lod_bricks[0][:]))            # every brick of the level
        mv, proj = H.hq_mip_matrices(wm, flip, angle, ortho, w, h, cam)
        assert np.allclose(np.array(p.inv_model_view[:]), np.linalg.inv(mv.astype(np.float64)).reshape(-1), atol=3e-6)
        assert np.allclose(np.array(p.inv_projection[:]), np.linalg.inv(proj.astype(np.float64)).reshape(-1), rtol=1e-5, atol=1e-5)
        assert p.bytes_per_voxel == raw.dtype.itemsize
        # GLRaycaster::RenderHQMIPInLoop: fRayStepsize = min(extension / voxels * 0.5 / sampleRateModifier)
        for b in bricks[:5]:
            want = min(np.float32(b.extension[a]) * (np.float32(1.0) / np.float32(b.voxels[a])) * np.float32(0.5) for a in range(3))
            assert b.ray_step == pytest.approx(float(want), rel=1e-6)
        r.renderHQMIP(wm, angle, flip, ortho)
        img = r.frameFloat()
        ref, vmax = H.hq_mip(p, bricks, store, tfb)
        err = np.abs(img - ref)
        # nearest table look-ups: a maximum within an ulp of a table boundary may pick the neighbouring entry
        assert (err > 1e-6).mean() < 5e-3 and err.max() <= 0.1, (wm, angle, flip, ortho, err.max(), (err > 1e-6).mean())
        assert (img[..., 3] == 1.0).all() and img[..., :3].max() > 0.3 and (img[..., :3].sum(axis=2) == 0).mean() > 0.05
    # after a MIP the 3-D path starts over
    r.setRendererTarget(2)
    r.paintUntilConverged(100)
    assert (r.frameFloat()[..., 3] < 1.0).any()
    r.cleanup()


def test_hq_mip_along_an_axis_is_the_volume_maximum():
    """orthographic, unrotated, coronal: every ray runs along z through voxel centres when the window maps voxels one to one
    -- the picture is tf(max over z) wherever the ray meets the volume; checked on the pixels whose ray passes within a
    tenth of a voxel of a voxel centre, against the plain numpy maximum of the raw volume (no brick, no restatement)"""
    vol = (64, 64, 40)
    raw = H.synth_blobs(vol, np.uint8, seed=11)
    w = h = 181                                     # 64 voxels across 1/1.414213 of the window: 2 pixels per voxel, roughly
    r = _renderer(raw, (w, h))
    tfb = np.zeros((256, 4), np.uint8)
    tfb[:, 0] = np.arange(256)                      # red = the value itself
    tfb[:, 3] = 255
    r.set1DTrans(tfb)
    r.renderHQMIP(2, 0.0, (False, False), True)
    img = r.frameFloat()
    zmax = raw.max(axis=0).astype(np.float32)       # [y][x]
    # pixel -> model x, y: ortho half width 0.5 * 1.414213, the volume spans [-0.5, 0.5] in x and y (64 x 64 is the largest side)
    half = np.float32(0.5 * 1.414213)
    xs = ((np.arange(w) + 0.5) / w * 2 - 1) * half
    vx = (xs + 0.5) * 64 - 0.5                       # voxel coordinate
    near = np.abs(vx - np.round(vx)) < 0.1
    inside = (vx > 0.5) & (vx < 62.5)
    sel = np.where(near & inside)[0]
    assert len(sel) >= 10
    got = img[np.ix_(sel, sel)][..., 0] * 255.0
    want = zmax[np.ix_(np.round(vx[sel]).astype(int), np.round(vx[sel]).astype(int))]
    # rays sample every half voxel along z and the filter is trilinear: within 0.1 voxel of a centre column the value is the
    # column's maximum up to the interpolation towards the neighbours
    neigh = np.abs(np.diff(raw.astype(np.float32), axis=2)).max() + np.abs(np.diff(raw.astype(np.float32), axis=1)).max()
    assert np.abs(got - want).max() <= 1.0 + 0.1 * neigh, np.abs(got - want).max()
    assert np.corrcoef(got.reshape(-1), want.reshape(-1))[0, 1] > 0.999
    r.cleanup()


def test_hq_mip_lod_follows_the_window():
    """PlanHQMIPFrame: the domain is halved while its smallest side is not below the largest window side, then one level back"""
    raw = H.synth_blobs((70, 45, 38), np.uint8)
    r = _renderer(raw, (16, 12))
    r.set1DTrans(H.tf_bytes(H.tf1d_std(256, 0.4, 0.5)))
    _, bricks, lod = r.planHQMIP(2)
    assert lod == 1                                            # 38 >= 16 -> 19 >= 16 -> 9: two halvings, one back
    lods = r.datasetInfo().lod_count
    for size in ((16, 12), (9, 9), (40, 30), (37, 80), (5, 4), (200, 100)):
        r.resize(size)                                         # the numpy rule is pinned to the reference's statements (CPU test)
        assert r.planHQMIP(2)[2] == H.hq_mip_lod((70, 45, 38), size, lods), size
        assert r.planHQMIP(2, mip_lod=False)[2] == 0
    r.resize((16, 12))
    _, bricks0, lod0 = r.planHQMIP(2, mip_lod=False)
    assert lod0 == 0 and len(bricks0) > len(bricks)
    r.renderHQMIP(2)
    assert r.frameFloat()[..., :3].max() > 0.2
    with pytest.raises(Exception):
        r.renderHQMIP(5)
    r.cleanup()
