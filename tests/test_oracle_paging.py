"""Oracle (CPU) check of the claim bench.py's e2e leg rests on: under request-driven paging
(GLGridLeaper.cpp:872-890, GLVolumePool::UploadBricks :1720-1789) a brick that RecomputeVisibility flags empty for
the transfer function is never asked for and never uploaded, and the converged frame is the frame of a pool that
holds every brick."""
import numpy as np

from oracle import host as H
from oracle import oracle as O


def _tf(n=256):
    t = H.tf1d_std(n, 0.6, 0.3)          # transparent for the lower values: the shell of the sphere is empty space
    t[:, 3] *= 0.2
    return t


def _setup(vol, brick, capacity):
    oc = O.build_octree(vol, brick, 2)
    pool = H.OraclePool(oc, capacity)
    pool.upload_first_brick()
    tf = _tf()
    lo, hi = H.tf_nonzero_limits(tf)
    maxv = float(oc["minmax"][:, 1].max())
    resc = maxv / (len(tf) - 1)
    pool.recompute_visibility(0, lo * resc, hi * resc)
    cam = H.Camera()
    p = H.raycast_params(cam, 96, 80, oc["vol"], trans_scale=255.0 / maxv)
    return oc, pool, tf, p, (lo * resc, hi * resc)


def test_empty_bricks_are_never_paged_in():
    vol = H.synth_sphere((96, 96, 96), np.uint8)
    oc, pool, tf, p, (lo, hi) = _setup(vol, 32, (6, 6, 6))
    img, subs, paged, _ = H.render_converged(p, pool, H.tf_bytes(tf))
    assert subs > 1 and paged > 0 and img[..., 3].max() > 0.05
    resident = np.nonzero(pool.meta >= 3)[0]
    empty_by_minmax = [i for i in range(pool.total) if not (hi >= pool.minmax[i, 0] and lo <= pool.minmax[i, 1])]
    assert len(empty_by_minmax) > 0, "the scene must contain bricks the transfer function makes empty"
    # no empty brick was ever uploaded (slot_brick keeps every brick a slot has held since the pool was created)
    held = set(int(b) for b in pool.slot_brick if b >= 0)
    assert not (held & set(empty_by_minmax) - {int(pool.lod_offset[-1])})
    assert all(pool.meta[i] in (1, 2) for i in empty_by_minmax if i != int(pool.lod_offset[-1]))
    assert len(resident) < pool.total

    # the same view from a pool in which EVERY brick is resident
    full = H.OraclePool(oc, (6, 6, 6))
    full.upload_first_brick()
    slot = 0
    for idx in range(full.total - 1):
        full._upload(idx, slot, full.time)
        full.time += 1
        slot += 1
    full.recompute_visibility(0, lo, hi)
    p.new_view = 1
    img_full, subs_full, paged_full, _ = H.render_converged(p, full, H.tf_bytes(tf))
    assert subs_full == 1 and paged_full == 0
    # the paged frame ends at the same LOD everywhere once converged, but through another slot assignment: pool-space
    # coordinates round differently in the last place, so compare with the image tolerance of the north star
    assert np.abs(img - img_full).max() <= 2.0 / 255.0 and H.psnr(img, img_full) >= 45.0
