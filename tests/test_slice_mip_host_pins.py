"""CPU: the host set-up of the 2-D views against THE REFERENCE'S OWN STATEMENTS, cut out of the .cpp files where
they lie and run (oracle/ref_extract.py): the texture window / aspect ratios / slice coordinate of GLRenderer::Render2DView
(GLRenderer.cpp:1129-1171) against oracle.host.slice_params, and the level-of-detail rule of AbstrRenderer::PlanHQMIPFrame
(AbstrRenderer.cpp:1220-1235) against oracle.host.hq_mip_lod.  The GPU tests compare the product's host (tvkr_get_slice_params,
tvkr_plan_hq_mip) with these two numpy functions field for field.  Needs /root/reference."""
import numpy as np
import pytest

from oracle import host as H
from oracle import ref_extract as X

pytestmark = pytest.mark.skipif(not X.available(), reason="/root/reference is not mounted")


def test_slice_view_setup_equals_the_reference_statements():
    rng = np.random.default_rng(11)
    f = np.float32
    for _ in range(25):
        dom = [int(v) for v in rng.integers(20, 900, 3)]
        lod = [max(1, d >> 3) for d in dom]
        brick = [l + 4 for l in lod]                         # the single brick of that level: its domain + 2 x 2 ghost voxels
        scale = [float(v) for v in rng.choice([0.5, 1.0, 1.0, 2.0, 1.37], 3)]
        w, h = int(rng.integers(100, 2000)), int(rng.integers(100, 1200))
        for wm in (0, 1, 2):
            si = int(rng.integers(0, dom[wm]))
            ref = X.slice2d(dom, lod, brick, scale, (w, h), wm, si)
            got = H.slice_params(dom, scale, lod, brick, 1, w, h, wm, si, False, (False, False), 0, 255.0, 0.0)
            assert [f(v) for v in ref["min_coords"]] == [f(v) for v in got["min_coords"]]
            assert [f(v) for v in ref["max_coords"]] == [f(v) for v in got["max_coords"]]
            assert f(ref["slice"]) == got["slice"]
            # RenderSlice: the quad's half extent = (aspect of the two axes shown) * window aspect, normalised by its maximum
            ah, av = (1 if wm == 0 else 0), (1 if wm == 2 else 2)
            q = [ref["aspect"][ah] * ref["win_aspect"][0], ref["aspect"][av] * ref["win_aspect"][1]]
            q = [f(q[0] / max(q)), f(q[1] / max(q))]
            assert q == [f(v) for v in got["quad"]], (q, got["quad"])


def test_hq_mip_lod_rule_equals_the_reference_statements():
    rng = np.random.default_rng(12)
    for _ in range(60):
        dom = [int(v) for v in rng.integers(8, 5000, 3)]
        win = (int(rng.integers(8, 4000)), int(rng.integers(8, 2200)))
        lods = int(rng.integers(1, 9))
        for mip_lod in (True, False):
            assert H.hq_mip_lod(dom, win, lods, mip_lod) == X.hq_mip_lod(dom, win, lods, mip_lod), (dom, win, lods, mip_lod)
    assert H.hq_mip_lod((2048, 2048, 2048), (1920, 1080), 7) == 0 and H.hq_mip_lod((70, 45, 38), (16, 12), 3) == 1


def test_hq_mip_matrices_equal_the_reference_statements():
    """the orthographic projection of the MIP view (GLRenderer.cpp:1184-1196) and m_maMIPRotation (the whole body of
    GLRenderer::RenderHQMIPPreLoop, :1256-1283), run from the reference's text, against oracle.host.hq_mip_matrices -- which the
    GPU test compares the product's plan with"""
    rng = np.random.default_rng(13)
    for _ in range(30):
        w, h = int(rng.integers(16, 4000)), int(rng.integers(16, 2200))
        wm = int(rng.integers(0, 3))
        flip = (bool(rng.integers(0, 2)), bool(rng.integers(0, 2)))
        ang = float(np.float32(rng.uniform(-360, 360)))
        po, mr = X.mip_matrices((w, h), wm, flip, ang)
        mv, proj = H.hq_mip_matrices(wm, flip, ang, True, w, h, H.Camera())
        assert np.allclose(po, proj, rtol=1e-6, atol=1e-7), np.abs(po - proj).max()
        assert np.allclose(mr, mv, rtol=1e-6, atol=1e-7), np.abs(mr - mv).max()


def test_slice_texture_coordinates_equal_the_quad_the_reference_draws():
    """GLRenderer::RenderSlice (GLRenderer.cpp:922-1014), run from the reference's text with glTexCoord3d / glVertex3d recording
    the quad: its extent and the texture coordinates interpolated over it at the pixel centres against
    oracle.host.slice_texcoords -- which volume axes run along the screen axes in each window mode, and what flipView swaps"""
    rng = np.random.default_rng(14)
    for _ in range(20):
        dom = [int(v) for v in rng.integers(20, 600, 3)]
        lod = [max(1, d >> 3) for d in dom]
        brick = [l + 4 for l in lod]
        scale = [float(v) for v in rng.choice([0.5, 1.0, 1.0, 2.0], 3)]
        w, h = int(rng.integers(40, 160)), int(rng.integers(40, 120))
        for wm in (0, 1, 2):
            flip = (bool(rng.integers(0, 2)), bool(rng.integers(0, 2)))
            si = int(rng.integers(0, dom[wm]))
            ref = X.slice2d(dom, lod, brick, scale, (w, h), wm, si)
            quad = X.slice_quad(wm, flip, ref["slice"], ref["min_coords"], ref["max_coords"], ref["aspect"], ref["win_aspect"])
            p = H.slice_params(dom, scale, lod, brick, 1, w, h, wm, si, False, flip, 0, 255.0, 0.0)
            tl, tr, br, bl = quad                              # (-x,+y) (+x,+y) (+x,-y) (-x,-y)
            assert tl[0] == -tr[0] == bl[0] == -br[0] and tl[1] == tr[1] == -br[1] == -bl[1]
            assert abs(tr[0] - float(p["quad"][0])) <= 2e-7 and abs(tr[1] - float(p["quad"][1])) <= 2e-7    # (printed with nine digits)
            inside, c = H.slice_texcoords(p)
            px, py = np.meshgrid(np.arange(w), np.arange(h))
            nx, ny = 2.0 * (px + 0.5) / w - 1.0, 2.0 * (py + 0.5) / h - 1.0
            u, v = (nx + tr[0]) / (2.0 * tr[0]), (ny + tr[1]) / (2.0 * tr[1])
            assert inside.sum() > 100
            for k in range(3):
                want = (1 - u) * (1 - v) * bl[2 + k] + u * (1 - v) * br[2 + k] + (1 - u) * v * tl[2 + k] + u * v * tr[2 + k]
                assert np.abs(c[k][inside] - want[inside]).max() <= 2e-6, (wm, flip, k, np.abs(c[k][inside] - want[inside]).max())
