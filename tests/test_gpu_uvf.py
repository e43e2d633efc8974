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
