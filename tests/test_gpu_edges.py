"""GPU: degenerate and ragged inputs at the boundaries of the path (the sizes the reference's own tests poke at,
IO/test/rebricking.h, IO/test/quantize.h): volumes smaller than a brick or a voxel thick, last bricks one voxel thick,
single-element and constant arrays in the quantiser, a one-brick dataset through the renderer."""
import numpy as np
import pytest

from oracle import host as H
from oracle import oracle as O
from tests.util import cuda_store_tight, image_errors

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("vol,dtype,brick,overlap", [
    ((1, 1, 1), np.uint8, 64, 2), ((2, 1, 1), np.uint16, 64, 2), ((5, 4, 3), np.uint8, 64, 2), ((61, 1, 1), np.uint8, 64, 2),
    ((64, 64, 64), np.uint8, 64, 2), ((1, 70, 1), np.uint16, 64, 2), ((121, 61, 60), np.uint8, 64, 2), ((17, 17, 17), np.uint16, 8, 2),
    ((33, 9, 40), np.uint8, 12, 4)])
def test_builder_degenerate_volumes(tvk, vol, dtype, brick, overlap):
    raw = H.synth_noise(vol, dtype, seed=sum(vol) + brick, maxval=255 if dtype == np.uint8 else 4095)
    oc = O.build_octree(raw, brick, overlap, False)
    info = tvk.build_octree(raw, brick, overlap)
    store, sizes = cuda_store_tight(tvk)
    assert info.total_bricks == oc["plan"].total_bricks and (sizes == oc["sizes"]).all()
    assert store.size == oc["store"].size and (store == oc["store"]).all()
    assert (tvk.build_minmax() == oc["minmax"]).all()
    h1 = tvk.build_hist1d()
    ref1 = O.hist1d(oc)
    assert h1.size == ref1.size and (h1 == ref1).all()
    h2, mg = tvk.build_hist2d(len(h1), float(oc["minmax"][:, 1].max()))
    r2, rmg = O.hist2d(oc, len(h1), oc["minmax"][:, 1].max())
    assert (h2 == r2).all() and (np.float32(mg) == np.float32(rmg) or (np.isnan(mg) and np.isnan(rmg)))
    tvk.build_release()


@pytest.mark.parametrize("arr", [np.array([7], np.int16), np.array([-3.5], np.float32), np.full(1000, 123, np.uint16),
                                 np.full(513, -1.25, np.float64), np.array([0, 65535], np.uint16), np.array([-128, 127], np.int8),
                                 np.arange(5000, dtype=np.int32) - 2500])
def test_quantize_tiny_and_constant_arrays(tvk, arr):
    want = O.quantize(arr)
    got = tvk.quantize(arr)
    assert got["out_bits"] == want["out_bits"] and got["generated"] == want["generated"]
    if want["generated"]:
        assert (got["data"] == want["data"]).all()
    if want["hist"] is None:
        assert got["hist"] is None
    else:
        assert got["hist"].size == want["hist"].size and (got["hist"] == want["hist"]).all()


def test_one_brick_dataset_through_the_renderer():
    """a volume that fits one brick: LOD 0 is the whole pyramid, the pool has a single (pinned) slot's worth of data"""
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    vol = H.synth_sphere((40, 36, 28), np.uint8)
    r = Renderer(0)
    r.loadRawVolume(vol)
    info = r.datasetInfo()
    tf = H.tf1d_std(info.tf_size, 0.3, 0.3)
    tf[:, 3] *= 0.2
    r.set1DTransFloat(tf)
    r.setLightingEnabled(False)              # (the reference's default is lighting on)
    r.resize((160, 120))
    r.setRendererTarget(RT_HEADLESS)
    assert r.paintUntilConverged() >= 1 and r.converged()
    img = r.frameFloat()
    assert img[..., 3].max() > 0.2 and np.isfinite(img).all()
    # the same through the oracle's frame loop
    from tests.test_gpu_host import _oracle_converged
    want, _, _, _ = _oracle_converged(vol, 64, tf, H.Camera(), 160, 120, tuple(info.pool_capacity))
    psnr, err = image_errors(img, want)
    assert psnr >= 45.0 and err <= 2.0 / 255.0, (psnr, err)
    r.cleanup()


@pytest.mark.parametrize("size", [(1, 1), (8, 4), (9, 5), (1, 37), (640, 1)])
def test_render_target_sizes_off_the_tile_grid(size):
    """render targets smaller than / not a multiple of the 8x4 pixel tiles a warp works on"""
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    from tests.test_gpu_host import _oracle_converged
    vol = H.synth_sphere((40, 36, 28), np.uint8)
    r = Renderer(0)
    r.loadRawVolume(vol)
    info = r.datasetInfo()
    tf = H.tf1d_std(info.tf_size, 0.3, 0.3)
    tf[:, 3] *= 0.2
    r.set1DTransFloat(tf)
    r.setLightingEnabled(False)
    r.resize(size)
    r.setRendererTarget(RT_HEADLESS)
    assert r.paintUntilConverged() >= 1
    img = r.frameFloat()
    assert img.shape == (size[1], size[0], 4) and np.isfinite(img).all()
    want, _, _, _ = _oracle_converged(vol, 64, tf, H.Camera(), size[0], size[1], tuple(info.pool_capacity))
    assert np.abs(img - want).max() <= 2.0 / 255.0
    rgba = r.frameRGBA8()
    assert rgba.shape == (size[1], size[0], 4)
    r.cleanup()


def test_volume_scale_and_odd_aspect_through_the_renderer():
    """anisotropic voxels (dataset scale) and a flat slab of a volume: brick extents, LOD and ray set-up all depend on them"""
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    from tests.test_gpu_host import _oracle_converged
    vol = H.synth_blobs((97, 33, 150), np.uint8)
    r = Renderer(0)
    r.loadRawVolume(vol, brick_size=32, scale=(1.0, 2.5, 0.7))
    info = r.datasetInfo()
    tf = H.tf1d_std(info.tf_size, 0.3, 0.3)
    tf[:, 3] *= 0.15
    r.set1DTransFloat(tf)
    r.setLightingEnabled(True)
    r.resize((171, 133))
    r.setRendererTarget(RT_HEADLESS)
    rot = H.rotation_y(0.9) @ H.rotation_x(-0.5)
    r.setRotation4x4(rot)
    r.paintUntilConverged()
    cam = H.Camera()
    cam.rotation = rot
    want, _, _, _ = _oracle_converged(vol, 32, tf, cam, 171, 133, tuple(info.pool_capacity), lighting=True, scale=(1.0, 2.5, 0.7))
    psnr, err = image_errors(r.frameFloat(), want)
    assert want[..., 3].max() > 0.2 and psnr >= 45.0 and err <= 2.0 / 255.0, (psnr, err)
    r.cleanup()
