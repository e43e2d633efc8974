"""Re-bricking (IO/DynamicBrickingDS.cpp) against vectors written by the REFERENCE ITSELF
(tests/golden/make_golden_rebrick.py: the reference's UVFDataset + DynamicBrickingDS compiled into oracle/_ref/ref_driver).
CPU part: the layout / mapping table of tvk_rebrick_plan / tvk_rebrick_toc (host code of the C-ABI library, no device).
GPU part: tvk_rebrick cuts the target store on the device."""
import gzip
import os

import numpy as np
import pytest

from imagevis3d_b200.context import TvkContext
from imagevis3d_b200.uvf import UvfFile

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    z = np.load(os.path.join(GOLD, "rebrick_cases.npz"))
    out = []
    for name in z["names"]:
        name = str(name)
        for ti, t in enumerate(z[name + "/targets"]):
            k = "%s/t%d" % (name, ti)
            out.append({"name": name, "target": tuple(int(v) for v in t), "uvf_gz": z[name + "/uvf_gz"], "layout": z[k + "/layout"],
                        "domain": z[k + "/domain"], "sizes": z[k + "/sizes"], "payload": z[k + "/payload"],
                        "minmax": [z[k + "/minmax%d" % m] for m in range(3)]})
    return out


CASES = _cases()


def _source(case, tmp_path):
    """the reference-written source file, read with the product's UVF reader -> (desc, store, offsets, sizes, minmax)"""
    f = tmp_path / "src.uvf"
    f.write_bytes(gzip.decompress(case["uvf_gz"].tobytes()))
    with UvfFile(str(f)) as u:
        bpv = 2 if u.info.vol.component_type == 1 else 1
        _, _, _, sizes = u.toc()
        store = u.read_bricks()
        mm = u.minmax()
        src_brick = [int(v) for v in u.info.vol.brick_size]
        overlap = int(u.info.vol.overlap)
    offsets = np.concatenate([[0], np.cumsum(sizes.astype(np.uint64).prod(axis=1) * bpv)[:-1]]).astype(np.uint64)
    desc = TvkContext.rebrick_desc(case["domain"], src_brick, case["target"], overlap, bpv)
    return desc, store, offsets, sizes, mm, bpv


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_rebrick_layout_and_mapping_match_the_reference(ci, tmp_path):
    case = CASES[ci]
    desc, store, soff, ssz, smm, bpv = _source(case, tmp_path)
    info = TvkContext.rebrick_plan(desc)
    lods = len(case["layout"])
    assert info.total_bricks == len(case["sizes"])
    assert [list(info.lod_bricks[l]) for l in range(lods)] == case["layout"].tolist()
    off, sz, si, so = TvkContext.rebrick_toc(desc)
    assert (sz == case["sizes"]).all()
    # the mapping table applied with numpy slicing to the source bricks gives the reference's target bricks
    dt = np.uint8 if bpv == 1 else np.uint16
    at = 0
    for i in range(len(sz)):
        s = ssz[si[i]]
        b = store[int(soff[si[i]]):int(soff[si[i]]) + int(s.prod()) * bpv].view(dt).reshape(s[2], s[1], s[0])
        t = b[so[i][2]:so[i][2] + sz[i][2], so[i][1]:so[i][1] + sz[i][1], so[i][0]:so[i][0] + sz[i][0]]
        n = int(sz[i].prod()) * bpv
        assert t.shape == (sz[i][2], sz[i][1], sz[i][0]), i
        assert (np.ascontiguousarray(t).view(np.uint8).reshape(-1) == case["payload"][at:at + n]).all(), (case["name"], i)
        at += n
        # MM_SOURCE: the source brick's file min/max; MM_PRECOMPUTE / MM_DYNAMIC: the target brick's own voxels
        assert tuple(case["minmax"][0][i]) == (smm[si[i]][0], smm[si[i]][1])
        assert tuple(case["minmax"][1][i]) == (float(t.min()), float(t.max())) == tuple(case["minmax"][2][i])
    assert at == case["payload"].size


def test_rebrick_refuses_what_the_reference_refuses(tmp_path):
    desc, *_ = _source(CASES[0], tmp_path)
    desc.tgt_brick[:] = [11, 12, 12]                          # core 7 does not divide core 16: DynamicBrickingDS throws
    with pytest.raises(Exception):
        TvkContext.rebrick_plan(desc)
    desc.tgt_brick[:] = [12, 12, 12]
    desc.overlap = 1                                          # the reference hard-codes 4 ghost voxels
    with pytest.raises(Exception):
        TvkContext.rebrick_plan(desc)
    desc.overlap = 2
    desc.tgt_brick[:] = [64, 64, 64]                          # larger than the source: clamped to the source size
    assert list(TvkContext.rebrick_plan(desc).tgt_brick) == [20, 20, 20]


@pytest.mark.gpu
@pytest.mark.parametrize("ci", range(len(CASES)))
def test_rebrick_kernel_matches_the_reference(ci, tmp_path):
    case = CASES[ci]
    desc, store, soff, ssz, smm, bpv = _source(case, tmp_path)
    t = TvkContext(0)
    dst, off, sz, mm = t.rebrick(desc, store, soff, ssz)
    t.close()
    at = 0
    for i in range(len(sz)):
        n = int(sz[i].prod()) * bpv
        assert (dst[int(off[i]):int(off[i]) + n] == case["payload"][at:at + n]).all(), (case["name"], i)
        at += n
    assert (mm == case["minmax"][1]).all() and (mm == case["minmax"][2]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_load_rebricked_through_the_renderer(mode, tmp_path):
    """renderer.loadRebricked (AbstrRenderer::LoadRebricked): what the renderer then holds -- written back out with saveUVF
    and read with the UVF reader -- is the reference's DynamicBrickingDS brick for brick, min/max per mode included; and
    the ray-guided renderer draws the same picture from the small bricks as from the file's own."""
    from imagevis3d_b200.renderer import Renderer
    case = CASES[0]                                           # u8 41x37x29, source bricks 20^3 -> 12^3
    src = tmp_path / "src.uvf"
    src.write_bytes(gzip.decompress(case["uvf_gz"].tobytes()))
    frames = []
    for rebrick in (False, True):
        r = Renderer(0)
        if rebrick:
            r.loadRebricked(str(src), 12, mode)
        else:
            r.loadDataset(str(src))
        r.resize((160, 120))
        r.setRenderMode(0)
        r.paintUntilConverged(200)
        frames.append(r.frameFloat())
        if rebrick:
            out = tmp_path / "rb.uvf"
            r.saveUVF(str(out))
            with UvfFile(str(out)) as u:
                assert [int(v) for v in u.info.vol.brick_size] == [12, 12, 12]
                assert (u.toc()[3] == case["sizes"]).all()
                assert (u.read_bricks() == case["payload"]).all()
                assert (u.minmax()[:, :2] == case["minmax"][mode]).all()
        r.cleanup()
    a, b = frames
    assert (b[..., 3] > 0).sum() > 1000
    # (not pixel-identical: the march restarts its sample sequence at every brick entry, in the reference too, so the
    #  sample positions depend on the brick size; the data parity is carried by the byte comparison above)
    mse = float(((a - b) ** 2).mean())
    assert 10.0 * np.log10(1.0 / max(mse, 1e-12)) >= 35.0, mse
