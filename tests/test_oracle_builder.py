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
