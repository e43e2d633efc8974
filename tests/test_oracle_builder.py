"""CPU: the builder oracle against the reference's golden vectors (and, in the build
container, against the reference itself compiled into oracle/_ref)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import golden_cases


def _check_case(raw, brick, overlap, median, ref):
    oc = O.build_octree(raw, brick, overlap, bool(median))
    p = oc["plan"]
    assert p.lod_count == len(ref["lod_size"])
    for l in range(p.lod_count):
        assert list(p.lod_size[l]) == list(ref["lod_size"][l])
        assert list(p.lod_bricks[l]) == list(ref["lod_bricks"][l])
        assert list(p.lod_aspect[l]) == list(ref["lod_aspect"][l])       # bit-exact doubles
    assert (oc["sizes"] == ref["sizes"]).all()
    assert oc["store"].size == ref["store"].size and (oc["store"] == ref["store"]).all()
    assert (oc["minmax"] == ref["minmax"]).all()
    h1 = O.hist1d(oc)
    assert h1.size == ref["hist1d"].size and (h1 == ref["hist1d"]).all()
    h2, mg = O.hist2d(oc, len(h1), oc["minmax"][:, 1].max())
    assert (h2 == ref["hist2d"]).all()
    assert np.float32(mg) == np.float32(ref["max_grad"])
    return oc


@pytest.mark.parametrize("case", range(7))
def test_oracle_matches_golden(case):
    g = golden_cases()[case]
    brick, overlap, median = [int(v) for v in g["params"]]
    _check_case(g["raw"], brick, overlap, median, g)


def test_rebricking_known_answers():
    """Tuvok/IO/test/rebricking.h:11-34,207-240: 8x8x1 uint16 ramp, brick 16, overlap 2, median."""
    raw = np.arange(64, dtype=np.uint16).reshape(1, 8, 8)
    oc = O.build_octree(raw, 16, 2, True)
    assert list(oc["sizes"][0]) == [12, 12, 5]                 # tdata_simple: bs == 12,12,5
    b = O.brick(oc, 0)
    assert (b[2, 2:10, 2:10] == raw[0]).all()                  # data equality behind the ghost offset
    assert b[:2].max() == 0 and b[3:].max() == 0 and b[2, :2].max() == 0   # ghost cells are zero
    assert oc["minmax"][0, 0] == 0.0 and oc["minmax"][0, 1] == 63.0         # range [0,63]


@pytest.mark.skipif(O.ref_driver_path() is None or not os.path.isdir("/root/reference"),
                    reason="reference build (oracle/_ref) only exists in the build container")
@pytest.mark.parametrize("vol,ty,brick,ov,med", [
    ((37, 29, 23), "u8", 12, 2, 0), ((49, 9, 5), "u8", 12, 2, 1), ((17, 33, 21), "u16", 10, 1, 0),
    ((50, 26, 14), "u16", 16, 4, 0), ((121, 61, 64), "u8", 64, 2, 0)])
def test_oracle_matches_reference_build(tmp_path, vol, ty, brick, ov, med):
    rng = np.random.default_rng(sum(vol))
    dt = np.uint8 if ty == "u8" else np.uint16
    raw = rng.integers(0, 256 if ty == "u8" else 65536, size=vol[::-1]).astype(dt)
    rawf, binf = str(tmp_path / "in.raw"), str(tmp_path / "out.bin")
    raw.tofile(rawf)
    subprocess.check_call([O.ref_driver_path(), "convert", rawf, *map(str, vol), ty, str(brick), str(ov), str(med),
                           "0", binf, "1" if ty == "u8" else "0"], stderr=subprocess.DEVNULL)
    ref = O.parse_ref_convert(binf, raw.dtype.itemsize)
    oc = O.build_octree(raw, brick, ov, bool(med))
    assert (oc["sizes"] == ref["sizes"]).all()
    assert (oc["store"] == ref["store"]).all()
    assert (oc["minmax"] == ref["minmax"]).all()
    if ty == "u8":
        h1 = O.hist1d(oc)
        assert (h1 == ref["hist1d"]).all()
        h2, mg = O.hist2d(oc, len(h1), oc["minmax"][:, 1].max())
        assert (h2 == ref["hist2d"]).all() and np.float32(mg) == np.float32(ref["max_grad"])


def quirk_mask(plan, sizes, brick, ov):
    """Per brick (ToC order): the ghost voxels FillOverlap never writes -- high ghost on two axes, low ghost on the third,
    the neighbours present (LOD >= 1 only).  In the reference's clamp-to-edge mode they keep stale bytes of its un-cleared
    temp buffer (ExtendedOctreeConverter.inc:261); everywhere else a clamp-mode brick is fully defined."""
    masks = []
    idx = 0
    core = brick - 2 * ov
    for l in range(int(plan.lod_count)):
        nb = [int(v) for v in plan.lod_bricks[l]]
        for bz in range(nb[2]):
            for by in range(nb[1]):
                for bx in range(nb[0]):
                    s = [int(v) for v in sizes[idx]]
                    idx += 1
                    k, j, i = np.meshgrid(np.arange(s[2]), np.arange(s[1]), np.arange(s[0]), indexing="ij")
                    P = np.zeros(k.shape, int)
                    M = np.zeros(k.shape, int)
                    beyond = np.zeros(k.shape, bool)
                    dom = [int(v) for v in plan.lod_size[l]]
                    for a, (c, t, b, n) in enumerate(((i, s[0], bx, nb[0]), (j, s[1], by, nb[1]), (k, s[2], bz, nb[2]))):
                        M += ((c < ov) & (b > 0)).astype(int)
                        hi = (c >= t - ov) & (b < n - 1)
                        P += hi.astype(int)
                        # ... and high ghost voxels beyond the domain, copied from the (not yet clamped) ghost region of a
                        # neighbour whose interior is thinner than the overlap
                        beyond |= hi & (b * core + c - ov >= dom[a])
                    if l > 0:
                        masks.append((((P == 2) & (M == 1)) | beyond).reshape(-1))
                        continue
                    # LOD 0 is read by linear file offset (a line towards a thin neighbour runs on into the next row /
                    # slice -- defined, and reproduced); only what lies past the end of the file is stale there
                    box = []
                    for c, t, b, n in ((i, s[0], bx, nb[0]), (j, s[1], by, nb[1]), (k, s[2], bz, nb[2])):
                        box.append(np.clip(c, ov if b == 0 else 0, t - (ov if b == n - 1 else 0) - 1) + b * core - ov)
                    lin = box[0] + box[1] * dom[0] + box[2] * dom[0] * dom[1]
                    masks.append((lin >= dom[0] * dom[1] * dom[2]).reshape(-1))
    return np.concatenate(masks)


@pytest.mark.skipif(O.ref_driver_path() is None or not os.path.isdir("/root/reference"),
                    reason="reference build (oracle/_ref) only exists in the build container")
@pytest.mark.parametrize("vol,ty,brick,ov,med", [
    ((70, 45, 38), "u8", 16, 2, 0), ((37, 29, 23), "u8", 12, 2, 1), ((50, 26, 40), "u16", 16, 2, 0), ((9, 7, 5), "u8", 12, 2, 0),
    ((8, 8, 6), "u8", 16, 2, 0), ((24, 12, 12), "u8", 16, 2, 0)])
def test_oracle_clamp_to_edge_matches_reference_build(tmp_path, vol, ty, brick, ov, med):
    """bClampToEdge: identical to the compiled reference on every voxel the reference defines (all but the stale ghost
    corners, see quirk_mask), min/max and histograms included when no brick has such a corner; the quirk voxels are the
    voxel at their global coordinate."""
    rng = np.random.default_rng(sum(vol) + 1)
    dt = np.uint8 if ty == "u8" else np.uint16
    raw = rng.integers(1, 256 if ty == "u8" else 65536, size=vol[::-1]).astype(dt)
    rawf, binf = str(tmp_path / "in.raw"), str(tmp_path / "out.bin")
    raw.tofile(rawf)
    subprocess.check_call([O.ref_driver_path(), "convert", rawf, *map(str, vol), ty, str(brick), str(ov), str(med),
                           "1", binf, "1" if ty == "u8" else "0"], stderr=subprocess.DEVNULL)
    ref = O.parse_ref_convert(binf, raw.dtype.itemsize)
    oc = O.build_octree(raw, brick, ov, bool(med), clamp=True)
    assert (oc["sizes"] == ref["sizes"]).all()
    q = np.repeat(quirk_mask(oc["plan"], oc["sizes"], brick, ov), raw.dtype.itemsize)
    assert ((oc["store"] == ref["store"]) | q).all()
    # the quirk voxels: the voxel at their global coordinate = what the zero-border build holds there when the corner
    # copy is not skipped ... checked through the flat volume of the level, rebuilt from the brick interiors
    flat_ok = 0
    idx = 0
    core = brick - 2 * ov
    at = 0
    for l in range(int(oc["plan"].lod_count)):
        nb = [int(v) for v in oc["plan"].lod_bricks[l]]
        sz = [int(v) for v in oc["plan"].lod_size[l]]
        flat = np.zeros((sz[2], sz[1], sz[0]), dt)
        first = idx
        for bz in range(nb[2]):
            for by in range(nb[1]):
                for bx in range(nb[0]):
                    b = O.brick(oc, idx); idx += 1
                    inner = b[ov:b.shape[0] - ov, ov:b.shape[1] - ov, ov:b.shape[2] - ov]
                    flat[bz * core:bz * core + inner.shape[0], by * core:by * core + inner.shape[1], bx * core:bx * core + inner.shape[2]] = inner
        idx = first
        for bz in range(nb[2]):
            for by in range(nb[1]):
                for bx in range(nb[0]):
                    b = O.brick(oc, idx); idx += 1
                    if l == 0:                                  # (linear file addressing: compared with the reference above)
                        flat_ok += 1
                        continue
                    k, j, i = np.meshgrid(np.arange(b.shape[0]), np.arange(b.shape[1]), np.arange(b.shape[2]), indexing="ij")
                    g = flat[np.clip(bz * core - ov + k, 0, sz[2] - 1), np.clip(by * core - ov + j, 0, sz[1] - 1), np.clip(bx * core - ov + i, 0, sz[0] - 1)]
                    assert (b == g).all(), (l, bx, by, bz)      # EVERY voxel of a clamp-mode brick of LOD >= 1: the level at the clamped coordinate
                    flat_ok += 1
    assert flat_ok == len(oc["sizes"])
    if not q.any():
        assert (oc["minmax"] == ref["minmax"]).all()
    if ty == "u8":
        h1 = O.hist1d(oc)
        assert (h1 == ref["hist1d"]).all()
        if not q.any():                                           # (stale ghost bytes would enter the gradient of border voxels)
            h2, mg = O.hist2d(oc, len(h1), oc["minmax"][:, 1].max())
            assert (h2 == ref["hist2d"]).all() and np.float32(mg) == np.float32(ref["max_grad"])
