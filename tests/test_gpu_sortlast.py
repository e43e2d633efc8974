"""GPU: sort-last z-slab rendering (clip box + slab parameter mapping + ordered compositor) on ONE
GPU, the ranks emulated one after the other.  (a) CUDA slabs + CUDA compositor == oracle slabs +
oracle compositor within the image tolerance; (b) the composited slabs reproduce the single-volume
render (different sample phase per slab, hence a looser bound)."""
import ctypes as C

import numpy as np
import pytest

from imagevis3d_b200 import sortlast as SL
from oracle import host as H
from oracle import oracle as O
from tests.util import image_errors, lockstep_render

pytestmark = pytest.mark.gpu


def _tf():
    t = H.tf1d_std(256, 0.3, 0.3)
    t[:, 3] *= 0.12
    t[:, 1] *= 0.7
    return t


@pytest.mark.parametrize("angle", [0.3, 1.9, 3.5])
def test_slabs_composite_like_one_volume(tvk, angle):
    import torch
    G = (64, 64, 192)
    vol = H.synth_blobs(G, np.uint8, seed=21)
    parts = SL.slab_partition(G, 3, halo=32)
    tf = _tf()
    lo, hi = H.tf_nonzero_limits(tf)
    cam = H.Camera()
    cam.rotation = H.rotation_y(angle) @ H.rotation_x(0.25)
    W, Hh = 160, 120
    # single volume reference (oracle)
    oc = O.build_octree(vol, 20, 2)
    pool = H.OraclePool(oc, (6, 6, 14))
    pool.upload_first_brick()
    pool.recompute_visibility(0, float(lo), float(hi))
    p_global = H.raycast_params(cam, W, Hh, G)
    whole, _, _, _ = H.render_converged(p_global, pool, H.tf_bytes(tf))
    order = SL.depth_order(np.array(p_global.eye_to_model[:]).reshape(4, 4), parts)
    cuda_parts, orc_parts = {}, {}
    for r, part in enumerate(parts):
        local = np.ascontiguousarray(vol[part["zs"]:part["ze"]])
        ocl = O.build_octree(local, 20, 2)
        pl = H.OraclePool(ocl, (6, 6, 10))
        pl.upload_first_brick()
        pl.recompute_visibility(0, float(lo), float(hi))
        p = H.raycast_params(cam, W, Hh, G, slab=part)
        res = lockstep_render(tvk, p, pl, H.tf_bytes(tf))
        psnr, err = image_errors(res["cuda"], res["oracle"])
        assert psnr >= 45.0 and err <= 2.0 / 255.0, (r, psnr, err)
        cuda_parts[r], orc_parts[r] = res["cuda"], res["oracle"]
    # oracle composite
    pix = W * Hh
    plist = [np.ascontiguousarray(orc_parts[r]) for r in order]
    want = np.zeros((Hh, W, 4), np.float32)
    arr = (C.c_void_p * len(plist))(*[a.ctypes.data for a in plist])
    O.lib().orc_composite_ordered(arr, C.c_uint32(len(plist)), C.c_uint64(pix), want.ctypes.data_as(C.c_void_p))
    # CUDA composite
    dev = [torch.from_numpy(np.ascontiguousarray(cuda_parts[r])).cuda() for r in order]
    out = torch.zeros((Hh, W, 4), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    tvk.composite_ordered([d.data_ptr() for d in dev], pix, out.data_ptr())
    tvk.synchronize()
    got = out.cpu().numpy()
    psnr, err = image_errors(got, want)
    assert want.max() > 0.05 and psnr >= 45.0 and err <= 2.0 / 255.0, (psnr, err)
    # against the un-partitioned render: same integral, different sample phase inside each slab
    psnr, err = image_errors(got, whole)
    assert psnr >= 38.0 and err <= 12.0 / 255.0, (psnr, err)


def test_clip_box_validation(tvk):
    from imagevis3d_b200._lib import TvkError
    from tests.util import to_cuda_params
    p = to_cuda_params(H.raycast_params(H.Camera(), 32, 32, (64, 64, 64)))
    p.box_min[2], p.box_max[2] = 0.6, 0.4
    with pytest.raises(TvkError):
        tvk.raycast(p)
