"""GPU: the UVF file boundary on both sides of the hot path.
 * K2 builder output written by CudaGridLeaper::SaveUVF must equal, BYTE FOR BYTE, the UVF the reference wrote
   for the same raw volume (tests/golden/ref_*.uvf.gz, from oracle/_ref/ref_driver writeuvf);
 * CudaGridLeaper::LoadDataset on a reference-written (zlib / lz4) UVF pages those bricks and renders the same
   image as the renderer that built the octree from the raw volume itself."""
import numpy as np
import pytest

from oracle import host as H
from tests.test_uvf import CASES, case_volume, golden_file

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ref_u8_none", "ref_u8_zlib", "ref_u16_zlib"])
def test_builder_to_uvf_is_byte_identical_to_reference(name, tmp_path):
    from imagevis3d_b200.renderer import Renderer
    vol, dt, brick, overlap, comp, level = CASES[name]
    r = Renderer(0)
    r.loadRawVolume(case_volume(name), brick_size=brick, overlap=overlap)
    out = str(tmp_path / "ours.uvf")
    r.saveUVF(out, compression=comp, level=level)
    ours, ref = open(out, "rb").read(), open(golden_file(name, tmp_path), "rb").read()
    assert len(ours) == len(ref) and ours == ref
    r.cleanup()


@pytest.mark.parametrize("name", ["ref_u8_zlib", "ref_u8_lz4", "ref_u16_zlib"])
def test_load_reference_uvf_renders_like_raw(name, tmp_path):
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    vol, dt, brick, overlap, comp, level = CASES[name]
    raw = case_volume(name)
    imgs = []
    for src in ("uvf", "raw"):
        r = Renderer(0)
        if src == "uvf":
            r.loadDataset(golden_file(name, tmp_path))
        else:
            r.loadRawVolume(raw, brick_size=brick, overlap=overlap, hist2d=True)
        info = r.datasetInfo()
        assert list(info.domain) == list(vol)
        tf = H.tf1d_std(info.tf_size, 0.4, 0.4)
        tf[:, 3] *= 0.2
        r.set1DTransFloat(tf)
        r.setLightingEnabled(True)
        r.resize((160, 120))
        r.setRendererTarget(RT_HEADLESS)
        r.setRotation4x4(H.rotation_y(0.5) @ H.rotation_x(0.2))
        r.paintUntilConverged()
        imgs.append((r.frameFloat().copy(), info.tf_size, tuple(info.range), info.max_grad_magnitude))
        r.cleanup()
    assert imgs[0][1:] == imgs[1][1:]
    assert imgs[0][0].max() > 0.05 and np.array_equal(imgs[0][0], imgs[1][0])


def test_load_legacy_rasterdatablock_file(tmp_path):
    """renderer.loadDataset on a legacy single-brick RDB file (c60.uvf's layout, rebuilt from tests/golden/c60.npz):
    the volume is re-bricked to 64^3 by the K2 builder and renders like the raw volume loaded directly."""
    import os
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    from tests.util import write_rdb_uvf
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c60.npz"))
    vol = np.ascontiguousarray(z["brick"])
    path = str(tmp_path / "c60_like.uvf")
    write_rdb_uvf(path, vol, 256, 4)
    imgs = []
    for src in ("uvf", "raw"):
        r = Renderer(0)
        if src == "uvf":
            r.loadDataset(path)
        else:
            r.loadRawVolume(vol, brick_size=64, hist2d=True)
        info = r.datasetInfo()
        assert list(info.domain) == [128, 128, 128] and info.tf_size == 256
        r.set1DTrans(np.ascontiguousarray(z["tf1d_rgba8"]))
        r.resize((256, 256))
        r.setRendererTarget(RT_HEADLESS)
        r.paintUntilConverged()
        imgs.append(r.frameFloat().copy())
        r.cleanup()
    assert imgs[0].max() > 0.3 and np.array_equal(imgs[0], imgs[1])


@pytest.mark.parametrize("kind", ["gridleaper", "raycaster"])
def test_batch_renderer_script(kind, tmp_path):
    """the BatchRenderer flow (loadDataset -> auto-loaded .1dt -> paint until converged -> captureSingleFrame) on a
    c60-like legacy UVF, both renderer types; the PNG must decode to the renderer's own 8-bit frame."""
    import os
    import struct
    import zlib
    from imagevis3d_b200 import batch
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    from tests.util import write_rdb_uvf
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c60.npz"))
    path = str(tmp_path / "c60.uvf")
    write_rdb_uvf(path, np.ascontiguousarray(z["brick"]), 256, 4)
    tfb = np.ascontiguousarray(z["tf1d_rgba8"]).astype(np.float32) / 255.0
    with open(str(tmp_path / "c60.1dt"), "w") as f:                     # TransferFunction1D text format
        f.write("%d\n" % len(tfb) + "".join("%.9g %.9g %.9g %.9g\n" % tuple(r) for r in tfb))
    out = str(tmp_path / "render.png")
    assert batch.main([path, out, "--size", "320x240", "--renderer", kind, "--rotate-y", "25"]) == 0
    raw = open(out, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    w, h = struct.unpack(">II", raw[16:24])
    assert (w, h) == (320, 240)
    idat = b""
    o = 8
    while o < len(raw):
        n, typ = struct.unpack(">I4s", raw[o:o + 8])
        if typ == b"IDAT":
            idat += raw[o + 8:o + 8 + n]
        o += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 4 * w)
    assert (rows[:, 0] == 0).all()                                       # filter type 0 on every row
    img = rows[:, 1:].reshape(h, w, 4)
    assert img[..., 3].max() == 255 and 0.05 < (img[..., 3] > 0).mean() < 0.9
