"""Convenience set-up used by bench.py / the full-size tests: build the bricked octree of a raw
volume that already sits in HBM and make EVERY pool-LOD brick resident (in-core rendering)."""
import numpy as np

from . import _lib as L


def load_resident(tvk, raw_dptr, dims, bpv, sampler=L.SAMPLER_EXACT, brick=64, overlap=2, median=False):
    """Returns dict(info, lod_count, total, capacity, offsets, sizes, minmax, slots)."""
    info = tvk.build_octree(raw_dptr=raw_dptr, volume_size=dims, bpv=bpv, brick=brick, overlap=overlap, median=median)
    off, sizes = tvk.build_toc()
    mm = tvk.build_minmax()
    lod_count = next(l for l in range(info.lod_count)
                     if info.lod_bricks[l][0] * info.lod_bricks[l][1] * info.lod_bricks[l][2] == 1) + 1
    total = int(info.lod_offset[lod_count - 1]) + 1
    nbx, nby = int(info.lod_bricks[0][0]) + 1, int(info.lod_bricks[0][1]) + 1
    cap = (nbx, nby, -(-total // (nbx * nby)))
    maxb = tuple(int(v) for v in sizes.max(axis=0))
    got = tvk.pool_create(dims, maxb, overlap, bpv, cap, lod_count, sampler)
    assert got == total
    store = tvk.build_store_ptr()
    slots = np.arange(total, dtype=np.uint32)
    slots[total - 1] = cap[0] * cap[1] * cap[2] - 1          # coarsest brick pinned in the last slot
    tvk.pool_upload_device(slots, [store + int(o) for o in off[:total]], sizes[:total])
    tvk.set_metadata(slots + 3)
    tvk.set_minmax(mm[:total])
    return dict(info=info, lod_count=lod_count, total=total, capacity=cap, offsets=off, sizes=sizes, minmax=mm,
                slots=slots)
