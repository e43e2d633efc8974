"""CPU: the restated oracle of the CLASSIC ray-caster (oracle/tvk_oracle_classic.c, incl. ClearView) against THE REFERENCE'S
OWN SHADER TEXT executed on the CPU (oracle/classicref.py: GLRaycaster-{1D,1D-light,2D,2D-light,ISO,ISO-CV}-FS.glsl with the
objects GLRaycaster::LoadShaders links to them, Compose-FS.glsl, Compose-CV-FS.glsl, compiled by g++ against the GLSL
value-type shim).  This pins the classic path's image oracle to something the builder did not write; the CUDA kernel is
compared with that oracle on the GPU (tests/test_gpu_classic.py).  Needs /root/reference (skipped on the GPU box)."""
import numpy as np
import pytest

from oracle import classicref as R
from oracle import host as H
from oracle import oracle as O

pytestmark = pytest.mark.skipif(not R.available(), reason="/root/reference (the shader sources) is not mounted")

from tests import classic_scenes as S
from tests.classic_scenes import CASES, W, Hh

_scene = S.scene


@pytest.mark.parametrize("name,dtype,mode,lighting,clip", CASES)
def test_classic_oracle_equals_reference_shaders(name, dtype, mode, lighting, clip):
    p, bricks, store, tf = _scene(dtype, mode, lighting, clip)
    assert len(bricks) >= 4
    a, _ = H.classic_render(p, bricks, store, tf)
    b = R.render(p, bricks, store, tf)
    assert b[..., 3].max() > 0.3, "empty picture: the scene would be vacuous"
    d = np.abs(a - b).max(axis=2)
    print("%s: max |diff| %.3g, PSNR %.1f dB, pixels > 1e-4: %d of %d" % (name, d.max(), H.psnr(a, b), int((d > 1e-4).sum()), d.size))
    if mode == 2:
        # a hit / miss flip on a silhouette pixel is a full-scale difference: bound their number
        assert (d > 1e-3).sum() <= 3, int((d > 1e-3).sum())
    else:
        # the oracle fuses multiply-adds where a GPU compiler would and skips the shading of transparent samples; the shader
        # text compiled here contracts nothing: rounding-level differences, except where they move a step count by one
        assert d.max() <= 2.0 / 255.0 and H.psnr(a, b) >= 80.0, (d.max(), H.psnr(a, b))
        assert (d > 1e-4).sum() <= max(2, d.size // 500), int((d > 1e-4).sum())


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_clearview_oracle_equals_reference_shaders(dtype):
    """GLRaycaster-ISO-CV-FS.glsl + Compose-CV-FS.glsl: hit targets of both surfaces and the composed picture"""
    p, bricks, store, _ = _scene(dtype, 2, False)
    cv = H.ClearViewParams()
    cv.cv_isoval = 1.6 * p.isoval
    cv.cv_color[:] = [1.0, 0.2, 0.1]
    cv.size, cv.context_scale, cv.border_scale = 3.0, 1.0, 60.0
    a0, ha, _ = H.classic_render_cv(p, cv, bricks, store)
    hit = ha[0][..., 3] != 0
    ys, xs = np.nonzero(hit & (ha[2][..., 3] != 0))
    assert len(ys) > 30, "no pixel shows both surfaces: vacuous"
    cv.pick_pos[:] = ha[0][ys[len(ys) // 2], xs[len(ys) // 2], :3].tolist()       # the lens on the surface
    a, ha, _ = H.classic_render_cv(p, cv, bricks, store)
    b, hb = R.render(p, bricks, store, None, cv)
    same_hits = (ha[0][..., 3] != 0) == (hb[0][..., 3] != 0)
    same_hits2 = (ha[2][..., 3] != 0) == (hb[2][..., 3] != 0)
    assert (~same_hits).sum() <= 3 and (~same_hits2).sum() <= 3
    both = (ha[0][..., 3] != 0) & (hb[0][..., 3] != 0)
    assert np.abs(ha[0][both] - hb[0][both]).max() <= 1e-4 and np.abs(ha[1][both][:, :3] - hb[1][both][:, :3]).max() <= 2e-3
    assert (ha[1][both][:, 3] == hb[1][both][:, 3]).mean() > 0.99                 # the tile the first surface was found in
    both2 = (ha[2][..., 3] != 0) & (hb[2][..., 3] != 0)
    assert both2.sum() > 30
    # the ClearView targets are half precision: a rounding-level difference upstream may land on the next half
    assert np.abs(ha[2][both2] - hb[2][both2]).max() <= 2e-3 and np.abs(ha[3][both2][:, :3] - hb[3][both2][:, :3]).max() <= 4e-3
    d = np.abs(a - b).max(axis=2)
    print("clearview %s: max |diff| %.3g, pixels > 2/255: %d of %d" % (np.dtype(dtype).name, d.max(), int((d > 2 / 255).sum()), d.size))
    # silhouette flips are full scale and the curvature term reads the neighbours' normals: bound their number
    assert (d > 2.0 / 255.0).sum() <= max(6, d.size // 250), int((d > 2.0 / 255.0).sum())
    inside = both2 & (np.sqrt(((ha[0][..., :3] - np.array(cv.pick_pos[:], np.float32)) ** 2).sum(axis=2)) * cv.size < 0.5)
    assert inside.sum() > 10 and (b[inside][:, 0] > b[inside][:, 2] + 0.05).mean() > 0.8      # the focus surface shows through the lens


# ------------------------------------------------------------------ 2-D slice views, MIP, stereo compositions
def _slice_struct(d):
    from imagevis3d_b200 import _lib as L
    p = L.SliceParams()
    p.width, p.height, p.window_mode, p.use_mip = d["width"], d["height"], d["window_mode"], d["use_mip"]
    p.slice = float(d["slice"])
    p.voxels[:] = d["voxels"]
    p.bytes_per_voxel = d["bpv"]
    p.min_coords[:] = [float(v) for v in d["min_coords"]]
    p.max_coords[:] = [float(v) for v in d["max_coords"]]
    p.quad[:] = [float(v) for v in d["quad"]]
    p.render_mode = d["render_mode"]
    p.trans_scale, p.gradient_scale = float(d["trans_scale"]), float(d["gradient_scale"])
    return p


def _single_brick(raw, brick=32):
    oc = O.build_octree(raw, brick, 2)
    pl = oc["plan"]
    idx = 0
    for l in range(int(pl.lod_count)):
        n = int(pl.lod_bricks[l][0]) * int(pl.lod_bricks[l][1]) * int(pl.lod_bricks[l][2])
        if n == 1:
            return oc, l, O.brick(oc, idx)
        idx += n
    raise AssertionError


@pytest.mark.parametrize("dtype,vol,scale", [(np.uint8, (70, 45, 38), (1, 1, 1)), (np.uint16, (50, 90, 41), (1, 0.5, 2.0))])
def test_slice_restatement_equals_reference_shaders(dtype, vol, scale):
    """oracle.host.slice_view (what the GPU test compares the slice kernel with) against 1D-slice-FS / 2D-slice-FS /
    MIP-slice-FS + Transfer-MIP-FS"""
    raw = H.synth_blobs(vol, dtype)
    oc, lod, brick = _single_brick(raw)
    n = 256 if dtype == np.uint8 else 4096
    tf1 = H.tf1d_std(n, 0.4, 0.5)
    tf1[:, 3] *= 0.7
    tf1 = H.tf_bytes(tf1)
    t2 = np.zeros((256, n, 4), np.uint8)
    t2[..., 0] = (np.arange(n) * 255 // (n - 1))[None, :]
    t2[..., 1] = np.arange(256)[:, None]
    t2[..., 2] = 90
    t2[..., 3] = 200
    mm = float(oc["minmax"][:, 1].max())
    w, h = 120, 90
    for wm in (0, 1, 2):
        for mip, mode in ((False, 0), (True, 0), (False, 1)):
            for flip in ((False, False), (True, True)):
                d = H.slice_params(vol, scale, [int(v) for v in oc["plan"].lod_size[lod]], brick.shape[::-1], raw.dtype.itemsize, w, h,
                                   wm, vol[wm] // 2, mip, flip, mode, mm, 0.01 if mode else 0.0)
                tf = t2 if mode == 1 else tf1
                a = H.slice_view(d, brick, tf)
                b = R.slice_view(_slice_struct(d), brick, tf)
                err = np.abs(a - b)
                # nearest table look-ups: a value within an ulp of a table boundary may pick the neighbouring entry
                assert (err > 1e-6).mean() < 3e-3 and err.max() <= 0.1, (wm, mip, mode, flip, err.max(), (err > 1e-6).mean())
                assert b[..., :3].max() > 0.05 and (b[..., 3] == 1.0).all()


def test_hq_mip_restatement_equals_reference_shaders():
    """oracle.host.hq_mip against GLRaycaster-MIP-Rot-FS + Transfer-MIP-FS, orthographic and perspective"""
    from imagevis3d_b200 import _lib as L
    vol = H.synth_blobs((70, 61, 47), np.uint8)
    oc = O.build_octree(vol, 32, 2)
    cam = H.Camera()
    cam.eye = np.array([0, 0, 3.0], np.float32)                    # far enough that the frustum keeps every brick
    w, h = 100, 76
    p0, bricks, _ = H.classic_plan(oc, cam, w, h, (0, 255, 256), lod=0, render_mode=0)
    assert len(bricks) == int(np.prod(oc["plan"].lod_bricks[0][:]))
    mv0 = (cam.rotation @ cam.translation @ H.look_at(cam.eye, cam.at, cam.up)).astype(np.float64)
    tf = H.tf_bytes(H.tf1d_std(256, 0.4, 0.5))
    for wm, angle, ortho in ((2, 0.0, True), (0, 30.0, True), (1, 200.0, False)):
        mv, proj = H.hq_mip_matrices(wm, (False, False), angle, ortho, w, h, H.Camera())
        p = L.MipParams()
        p.width, p.height, p.bytes_per_voxel, p.trans_scale = w, h, 1, p0.trans_scale
        p.inv_model_view[:] = np.linalg.inv(mv.astype(np.float64)).astype(np.float32).reshape(-1).tolist()
        p.inv_projection[:] = np.linalg.inv(proj.astype(np.float64)).astype(np.float32).reshape(-1).tolist()
        bl = []
        for b in bricks:                                           # the same bricks under the MIP's model-view
            nb = H.ClassicBrick.from_buffer_copy(b)
            m = mv0 @ np.array(b.eye_to_tex[:], np.float64).reshape(4, 4)
            nb.eye_to_tex[:] = (np.linalg.inv(mv.astype(np.float64)) @ m).astype(np.float32).reshape(-1).tolist()
            bl.append(nb)
        a, _ = H.hq_mip(p, bl, oc["store"], tf)
        c = R.hq_mip(p, bl, oc["store"], tf)
        err = np.abs(a - c)
        assert (err > 1e-6).mean() < 5e-3 and err.max() <= 0.1, (wm, angle, ortho, err.max(), (err > 1e-6).mean())
        assert c[..., :3].max() > 0.3 and (c[..., :3].sum(axis=2) == 0).mean() > 0.05


def test_stereo_compose_restatement_equals_reference_shaders():
    rng = np.random.default_rng(4)
    l, r = rng.random((20, 32, 4), np.float32), rng.random((20, 32, 4), np.float32)
    assert np.abs(R.stereo_compose(l, r, 0) - H.stereo_compose(l, r, 0)).max() <= 1e-6
    assert np.array_equal(R.stereo_compose(l, r, 1), H.stereo_compose(l, r, 1))
    for af in (0, 1):
        assert np.array_equal(R.stereo_compose(l, r, 3, af_id=af), H.stereo_compose(l, r, 3, af_id=af))
    # side by side: the shader samples GL_NEAREST textures exactly ON a texel boundary (x * 2); which of the two texels an
    # implementation returns there is not defined -- the kernel takes 2x+1, the shader text in fp32 lands on either
    sb = R.stereo_compose(l, r, 2)
    for x in range(32):
        src, s0 = (l, 2 * x) if x < 16 else (r, 2 * x - 32)
        for y in range(20):
            assert np.array_equal(sb[y, x], src[y, s0]) or np.array_equal(sb[y, x], src[y, s0 + 1])
