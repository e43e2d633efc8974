"""GPU: the classic per-brick ray-caster (GLRaycaster + AbstrRenderer::PlanFrame).
 * kernel: tvk_raycast_classic against the oracle's restatement of the GLRaycaster shaders on the SAME frame plan;
 * host: CudaGridLeaper::PlanClassicFrame against the oracle's independent numpy restatement of PlanFrame /
   BuildSubFrameBrickList / SetBrickDepShaderVars / ComputeEyeToTextureMatrix;
 * BASELINE configs[0] (c60, 1024x1024, 1D-TF, lighting on and off) through the classic path."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import host as H
from oracle import oracle as O
from tests.util import image_errors

pytestmark = pytest.mark.gpu


def _tf(n):
    t = H.tf1d_std(n, 0.3, 0.3)
    t[:, 3] *= 0.15
    t[:, 2] *= 0.5
    return t


def _device_store(r):
    """the builder's brick store as the device holds it (same offsets the classic bricks carry); its size is not
    exposed through the renderer handle, so it is found by bisection on tvk_build_get_bricks' range check"""
    lib = r._lib
    lib.tvkr_context.restype = C.c_void_p
    ctx = C.c_void_p(lib.tvkr_context(r._h))
    probe = np.zeros(1, np.uint8)

    def ok(off):
        return lib.tvk_build_get_bricks(ctx, probe.ctypes.data_as(C.c_void_p), C.c_uint64(off), C.c_uint64(1)) == 0
    lo, hi = 0, 1 << 40
    while hi - lo > 1:
        mid = (lo + hi) // 2
        lo, hi = (mid, hi) if ok(mid) else (lo, mid)
    out = np.zeros(lo + 1, np.uint8)
    assert lib.tvk_build_get_bricks(ctx, out.ctypes.data_as(C.c_void_p), C.c_uint64(0), C.c_uint64(out.size)) == 0
    return out


def _to_oracle(p, bricks):
    q = H.ClassicParams()
    assert C.sizeof(q) == C.sizeof(p)
    C.memmove(C.byref(q), C.byref(p), C.sizeof(p))
    out = []
    for b in bricks:
        ob = H.ClassicBrick()
        assert C.sizeof(ob) == C.sizeof(b)
        C.memmove(C.byref(ob), C.byref(b), C.sizeof(b))
        out.append(ob)
    return q, out


@pytest.mark.parametrize("lighting,dtype", [(False, np.uint8), (True, np.uint8), (True, np.uint16)])
def test_classic_kernel_matches_oracle(lighting, dtype):
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    vol = H.synth_blobs((150, 131, 97), dtype)
    r = Renderer(0)
    r.loadRawVolume(vol, brick_size=32)
    r.setRendererType(True)
    info = r.datasetInfo()
    tf = _tf(info.tf_size)
    r.set1DTransFloat(tf)
    r.setLightingEnabled(lighting)
    r.resize((203, 149))
    r.setRendererTarget(RT_HEADLESS)
    r.setRotation4x4(H.rotation_y(0.6) @ H.rotation_x(0.3))
    p, bricks, lod = r.planClassicFrame()
    assert len(bricks) > 20 and lod == 0
    assert r.paintUntilConverged() == 1 and r.converged()
    got = r.frameFloat()
    q, ob = _to_oracle(p, bricks)
    want, samples = H.classic_render(q, ob, _device_store(r), H.tf_bytes(tf))
    psnr, err = image_errors(got, want)
    assert want[..., 3].max() > 0.3 and psnr >= 45.0 and err <= 2.0 / 255.0, (psnr, err)
    assert abs(r.perf()["samples"] - samples) <= samples * 1e-4
    r.cleanup()


def _tf2d(n):
    v = np.linspace(0, 1, n, dtype=np.float32)
    g = np.linspace(0, 1, 256, dtype=np.float32)
    t = np.zeros((256, n, 4), np.float32)
    t[..., 0] = v[None, :]
    t[..., 1] = 0.2 + 0.8 * g[:, None]
    t[..., 2] = 0.4
    t[..., 3] = np.clip((v[None, :] - 0.3) * 0.5, 0, 1) * (0.3 + 0.7 * g[:, None])
    return (np.clip(t, 0, 1) * np.float32(255)).astype(np.uint8)


@pytest.mark.parametrize("lighting,dtype", [(False, np.uint8), (True, np.uint8), (True, np.uint16)])
def test_classic_2d_transfer_function_matches_oracle(lighting, dtype):
    """GLRaycaster-2D-FS.glsl / GLRaycaster-2D-light-FS.glsl"""
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    vol = H.synth_blobs((150, 131, 97), dtype)
    r = Renderer(0)
    r.loadRawVolume(vol, brick_size=32)
    r.setRendererType(True)
    info = r.datasetInfo()
    tf2 = _tf2d(info.tf_size)
    nzv = np.nonzero(tf2[..., 3].max(axis=0))[0]
    r.set2DTrans(tf2, (float(nzv.min()), float(nzv.max()), 0.0, 1.0))
    r.setRenderMode(1)
    r.setLightingEnabled(lighting)
    r.resize((203, 149))
    r.setRendererTarget(RT_HEADLESS)
    r.setRotation4x4(H.rotation_y(-0.8) @ H.rotation_x(0.5))
    p, bricks, lod = r.planClassicFrame()
    assert p.render_mode == 1 and len(bricks) > 20 and lod == 0
    assert r.paintUntilConverged() == 1 and r.converged()
    got = r.frameFloat()
    q, ob = _to_oracle(p, bricks)
    want, samples = H.classic_render(q, ob, _device_store(r), tf2)
    psnr, err = image_errors(got, want)
    assert want[..., 3].max() > 0.3 and psnr >= 45.0 and err <= 2.0 / 255.0, (psnr, err)
    assert len(np.unique(want[..., 1].round(3))) > 50          # the gradient axis of the table is exercised
    assert abs(r.perf()["samples"] - samples) <= samples * 1e-4
    r.cleanup()


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_classic_isosurface_matches_oracle(dtype):
    """GLRaycaster-ISO-FS.glsl + RefineIsosurface.glsl + depth test + Compose-FS.glsl"""
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    vol = H.synth_blobs((150, 131, 97), dtype)
    r = Renderer(0)
    r.loadRawVolume(vol, brick_size=32)
    r.setRendererType(True)
    info = r.datasetInfo()
    r.set1DTransFloat(_tf(info.tf_size))
    r.setRenderMode(2)
    r.setIsoValue(0.25 * float(vol.max()))
    r.setIsosurfaceColor((0.9, 0.6, 0.3))
    r.resize((203, 149))
    r.setRendererTarget(RT_HEADLESS)
    r.setRotation4x4(H.rotation_y(0.6) @ H.rotation_x(0.3))
    p, bricks, lod = r.planClassicFrame()
    assert p.render_mode == 2 and len(bricks) > 5 and lod == 0
    assert r.paintUntilConverged() == 1 and r.converged()
    got = r.frameFloat()
    q, ob = _to_oracle(p, bricks)
    want, samples = H.classic_render(q, ob, _device_store(r), H.tf_bytes(_tf(info.tf_size)))
    cover = float((want[..., 3] > 0).mean())
    assert 0.03 < cover < 0.95, cover
    # a hit either exists in both images or in neither (same fp32 operations), up to a handful of silhouette pixels
    assert float(((got[..., 3] > 0) != (want[..., 3] > 0)).mean()) <= 1e-3
    psnr, err = image_errors(got, want)
    assert psnr >= 45.0, (psnr, err)
    assert abs(r.perf()["samples"] - samples) <= samples * 1e-4
    r.cleanup()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_classic_plan_matches_oracle_host(mode):
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    vol = H.synth_blobs((150, 131, 97), np.uint8)
    r = Renderer(0)
    r.loadRawVolume(vol, brick_size=32)
    r.setRendererType(True)
    info = r.datasetInfo()
    tf = _tf(info.tf_size)
    r.set1DTransFloat(tf)
    if mode == 1:
        tf2 = _tf2d(info.tf_size)
        nzv = np.nonzero(tf2[..., 3].max(axis=0))[0]
        r.set2DTrans(tf2, (float(nzv.min()), float(nzv.max()), 0.0, 1.0))
    r.setRenderMode(mode)
    iso = 0.7 * float(vol.max())
    r.setIsoValue(iso)
    r.setLightingEnabled(True)
    r.resize((203, 149))
    r.setRendererTarget(RT_HEADLESS)
    rot = H.rotation_y(2.1) @ H.rotation_x(-0.4)
    r.setRotation4x4(rot)
    p, bricks, lod = r.planClassicFrame()
    cam = H.Camera()
    cam.rotation = rot
    oc = O.build_octree(vol, 32, 2)
    lo, hi = H.tf_nonzero_limits(tf)
    if mode == 1:
        lo, hi = int(nzv.min()), int(nzv.max())
    maxv = float(vol.max())
    po, bo, lod_o = H.classic_plan(oc, cam, 203, 149, (lo, hi, len(tf)), lighting=True, trans_scale=255.0 / maxv,
                                   render_mode=mode, gradient_scale=p.gradient_scale, iso_value=iso)
    assert lod == lod_o and len(bricks) == len(bo)
    if mode == 2:                                          # bricks below the isovalue are dropped
        p0, b0, _ = H.classic_plan(oc, cam, 203, 149, (lo, hi, len(tf)), lighting=True, trans_scale=255.0 / maxv)
        assert len(bo) < len(b0)
    for name, _ in po._fields_:
        a, b = getattr(p, name), getattr(po, name)
        if hasattr(a, "__len__"):
            assert np.allclose(np.array(a[:], np.float64), np.array(b[:], np.float64), rtol=2e-5, atol=2e-6), name
        else:
            assert a == pytest.approx(b, rel=2e-5, abs=2e-6), name
    for x, y in zip(bricks, bo):
        assert list(x.voxels) == list(y.voxels)
        for name in ("center", "extension", "eye_to_tex"):
            assert np.allclose(np.array(getattr(x, name)[:]), np.array(getattr(y, name)[:]), rtol=3e-5, atol=3e-6), name
        assert x.ray_step == pytest.approx(y.ray_step, rel=1e-6)
    r.cleanup()


@pytest.mark.parametrize("lighting", [True, False])
def test_config1_c60_classic_raycaster(lighting):
    """BASELINE configs[0] as the reference would run it under llvmpipe: OpenGL_Raycaster, 1024x1024, c60.1dt."""
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c60.npz"))
    vol = np.ascontiguousarray(z["brick"])
    tfb = np.ascontiguousarray(z["tf1d_rgba8"])
    r = Renderer(0)
    r.loadRawVolume(vol, brick_size=64)
    r.setRendererType(True)
    r.set1DTrans(tfb)
    r.setLightingEnabled(lighting)
    r.resize((1024, 1024))
    r.setRendererTarget(RT_HEADLESS)
    p, bricks, lod = r.planClassicFrame()
    r.paintUntilConverged()
    got = r.frameFloat()
    q, ob = _to_oracle(p, bricks)
    want, _ = H.classic_render(q, ob, _device_store(r), tfb)
    psnr, err = image_errors(got, want)
    assert want[..., 3].max() > 0.3 and psnr >= 45.0 and err <= 2.0 / 255.0, (psnr, err)
    r.cleanup()


@pytest.mark.parametrize("mode", [0, 2])
def test_classic_clip_plane(mode):
    """m_bClipPlaneOn on the classic path: AbstrRenderer::Clipped drops bricks, RenderBox clips the rest
    (GLRaycaster.cpp:327-337)."""
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    vol = H.synth_blobs((150, 131, 97), np.uint8)
    r = Renderer(0)
    r.loadRawVolume(vol, brick_size=32)
    r.setRendererType(True)
    info = r.datasetInfo()
    tf = _tf(info.tf_size)
    r.set1DTransFloat(tf)
    r.setRenderMode(mode)
    iso = 0.25 * float(vol.max())
    r.setIsoValue(iso)
    r.resize((203, 149))
    r.setRendererTarget(RT_HEADLESS)
    rot = H.rotation_y(0.6) @ H.rotation_x(0.3)
    r.setRotation4x4(rot)
    p0, bricks0, _ = r.planClassicFrame()
    r.paintUntilConverged()
    full = r.frameFloat()
    plane = (0.6, 0.3, 0.74, 0.05)
    r.setClipPlane(plane)
    r.enableClipPlane(True)
    p, bricks, lod = r.planClassicFrame()
    assert p.clip_enabled == 1 and 0 < len(bricks) < len(bricks0)
    r.paintUntilConverged()
    got = r.frameFloat()
    q, ob = _to_oracle(p, bricks)
    want, _ = H.classic_render(q, ob, _device_store(r), H.tf_bytes(tf))
    psnr, err = image_errors(got, want)
    assert psnr >= 45.0, (psnr, err)
    assert np.abs(got - full).max() > 0.05                  # the plane cuts something away
    # the host's plan against the oracle's independent one
    cam = H.Camera()
    cam.rotation = rot
    oc = O.build_octree(vol, 32, 2)
    lo, hi = H.tf_nonzero_limits(tf)
    po, bo, _ = H.classic_plan(oc, cam, 203, 149, (lo, hi, len(tf)), trans_scale=255.0 / float(vol.max()), render_mode=mode,
                               gradient_scale=p.gradient_scale, iso_value=iso, clip_plane=plane)
    assert len(bo) == len(bricks) and po.clip_enabled == 1
    assert np.allclose(np.array(p.clip_plane[:]), np.array(po.clip_plane[:]), rtol=2e-5, atol=2e-6)
    for x, y in zip(bricks, bo):
        assert list(x.voxels) == list(y.voxels) and np.allclose(np.array(x.center[:]), np.array(y.center[:]), rtol=3e-5, atol=3e-6)
    r.cleanup()


def test_classic_coarse_to_fine_ladder():
    """AbstrRenderer::PlanFrame + ComputeMaxLODForCurrentView on the classic path: an interactive view starts coarse and
    every paint() renders the brick list of the next finer LOD down to the one the view needs; the times of the first two
    sub-frames move the start LOD of the NEXT view (three views of grace before it gets coarser).  The last sub-frame is the
    image a headless renderer produces in one paint."""
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS, RT_INTERACTIVE
    vol = H.synth_blobs((150, 131, 97), np.uint8)
    tf = None
    frames = {}
    for target in (RT_HEADLESS, RT_INTERACTIVE):
        r = Renderer(0)
        r.loadRawVolume(vol, brick_size=20)
        r.setRendererType(True)
        info = r.datasetInfo()
        tf = _tf(info.tf_size)
        r.set1DTransFloat(tf)
        r.resize((203, 149))
        r.setRendererTarget(target)
        r.setRotation4x4(H.rotation_y(0.6) @ H.rotation_x(0.3))
        if target == RT_HEADLESS:
            assert r.paintUntilConverged() == 1
            frames["headless"] = r.frameFloat()
            r.cleanup()
            continue
        r.setMaxMSPerFrame(100.0)
        r.overrideSubframeMs(1.0)                    # "fast": deterministic, whatever the kernel really took
        # ---- view 1: no timings yet -> start at the coarsest LOD, one LOD finer per paint
        lods = []
        while True:
            r.paint()
            st = r.lodState()
            lods.append(st["current_lod"])
            if r.converged():
                break
        top = st["max_lod_index"]
        assert top >= 3 and lods == list(range(top, st["min_lod_for_view"] - 1, -1)), lods
        assert st["min_lod_for_view"] == 0
        assert np.array_equal(r.frameFloat(), frames["headless"])
        # ---- views 2..: both timed sub-frames were fast -> the skip grows by one per view, the start LOD gets finer
        starts = []
        for k in range(3):
            r.setRotation4x4(H.rotation_y(0.7 + 0.1 * k) @ H.rotation_x(0.3))
            n = r.paintUntilConverged()
            st = r.lodState()
            starts.append(int(st["start_lod_offset"]))
            assert n == st["start_lod_offset"] - st["min_lod_for_view"] + 1
        assert starts == [top - 1, top - 2, top - 3], starts
        # ---- slow first sub-frames: three views of grace, the fourth starts one LOD coarser
        r.overrideSubframeMs(500.0)
        starts = []
        for k in range(5):
            r.setRotation4x4(H.rotation_y(1.2 + 0.1 * k) @ H.rotation_x(0.3))
            r.paintUntilConverged()
            starts.append(int(r.lodState()["start_lod_offset"]))
        assert starts[0] == starts[1] == starts[2] == starts[3] and starts[4] == starts[3] + 1, starts
        r.cleanup()


def test_classic_stereo_pair():
    """Stereo on the classic path (AbstrRenderer::PlanFrame :1196-1201, BuildLeftEyeSubFrameBrickList :843-858,
    GLRaycaster::Render3DInLoop :389): the right eye draws the left eye's brick SET in its own front-to-back order with its
    own matrices; each eye's picture equals the oracle's restated shaders on that eye's plan, the frame is the composition."""
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    vol = H.synth_blobs((150, 131, 97), np.uint8)
    r = Renderer(0)
    r.loadRawVolume(vol, brick_size=32)
    r.setRendererType(True)
    info = r.datasetInfo()
    tf = _tf(info.tf_size)
    r.set1DTransFloat(tf)
    r.setLightingEnabled(True)
    w, h = 203, 149
    r.resize((w, h))
    r.setRendererTarget(RT_HEADLESS)
    rot = H.rotation_y(0.6) @ H.rotation_x(0.3)
    r.setRotation4x4(rot)
    p_mono, bricks_mono, _ = r.planClassicFrame()
    r.setStereoEnabled(True)
    r.setStereoEyeDist(0.08)
    r.setStereoMode("rb")
    plans = [r.planClassicFrame(eye=e) for e in (0, 1)]
    key = lambda b: int(b.store_offset)                                                # noqa: E731
    assert sorted(map(key, plans[0][1])) == sorted(map(key, plans[1][1])) == sorted(map(key, bricks_mono))
    # each eye's list is sorted by ITS distance to the bricks (nearest corner, pulled in by 0.4999 of the extension)
    cam = H.Camera()
    cam.rotation = rot
    eyes = H.stereo_matrices(cam, np.float32(w) / np.float32(h), 1.0, 0.08)
    for e in (0, 1):
        mv = (cam.rotation @ cam.translation @ eyes[e][0]).astype(np.float32)
        d = []
        for b in plans[e][1]:
            c, ext = np.array(b.center[:], np.float32), np.array(b.extension[:], np.float32)
            corners = np.array([[c[0] + sx * ext[0] * 0.4999, c[1] + sy * ext[1] * 0.4999, c[2] + sz * ext[2] * 0.4999, 1.0]
                                for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32)
            d.append(np.sqrt(((corners @ mv)[:, :3] ** 2).sum(axis=1)).min())
        d = np.array(d)
        assert (np.diff(d) >= -1e-5).all(), e
        assert np.allclose(np.array(plans[e][0].inv_model_view[:]), np.linalg.inv(mv.astype(np.float64)).reshape(-1), atol=2e-5)
        assert np.allclose(np.array(plans[e][0].inv_projection[:]), np.linalg.inv(eyes[e][1].astype(np.float64)).reshape(-1), rtol=1e-4, atol=1e-4)
    assert [key(b) for b in plans[0][1]] != [key(b) for b in plans[1][1]] or len(bricks_mono) < 3
    assert r.paintUntilConverged() == 1 and r.converged()
    store = _device_store(r)
    pics = []
    for e in (0, 1):
        q, ob = _to_oracle(*plans[e][:2])
        want, _ = H.classic_render(q, ob, store, H.tf_bytes(tf))
        got = r.eyeFrameFloat(e)
        psnr, err = image_errors(got, want)
        assert want[..., 3].max() > 0.3 and psnr >= 45.0 and err <= 2.0 / 255.0, (e, psnr, err)
        pics.append(got)
    assert np.abs(pics[0] - pics[1]).max() > 0.05
    assert np.abs(r.frameFloat() - H.stereo_compose(pics[0], pics[1], 0)).max() <= 1e-6
    r.setStereoMode("af")
    r.paintUntilConverged()
    assert np.array_equal(r.frameFloat(), r.eyeFrameFloat(0))
    r.toggleStereoFrame()
    assert np.array_equal(r.frameFloat(), r.eyeFrameFloat(1))
    r.cleanup()


def test_clearview_lens_on_the_classic_isosurface():
    """ClearView (GLRaycaster::Render3DInLoop's second pass + GLRaycaster-ISO-CV-FS.glsl, Compose-CV-FS.glsl): the picture
    against the oracle's restatement on the same plan; outside the lens it is the plain isosurface picture; the focus set
    from a window position is the first surface's hit there; a position without a hit takes the lens away."""
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS, RM_ISOSURFACE
    vol = H.synth_blobs((150, 131, 97), np.uint8)
    r = Renderer(0)
    r.loadRawVolume(vol, brick_size=32)
    r.setRendererType(True)
    w, h = 203, 149
    r.resize((w, h))
    r.setRendererTarget(RT_HEADLESS)
    r.setRotation4x4(H.rotation_y(0.6) @ H.rotation_x(0.3))
    r.setRenderMode(RM_ISOSURFACE)
    iso = 0.12 * float(vol.max())
    r.setIsoValue(iso)
    r.paintUntilConverged()
    plain = r.frameFloat()
    assert (plain[..., 3] == 1.0).mean() > 0.08
    r.setClearViewEnabled(True)
    r.setCVIsoValue(1.6 * iso)
    r.setCVColor((1.0, 0.2, 0.1))
    r.setCVSize(3.0)
    assert r.getClearViewEnabled() and abs(r.getCVIsoValue() - 1.6 * iso) < 1e-3 and r.checkForRedraw()
    store = _device_store(r)

    def both():
        p, bricks, _ = r.planClassicFrame()
        cv = r.clearViewParams()
        r.paintUntilConverged()
        got = r.frameFloat()
        q, ob = _to_oracle(p, bricks)
        cq = H.ClearViewParams()
        assert C.sizeof(cq) == C.sizeof(cv)
        C.memmove(C.byref(cq), C.byref(cv), C.sizeof(cv))
        want, hits, _ = H.classic_render_cv(q, cq, ob, store)
        d = np.abs(got - want).max(axis=2)
        # (a hit / miss flip on a silhouette pixel is a full-scale difference, and the curvature term reads the neighbours' normals)
        assert (d > 2.0 / 255.0).sum() <= max(6, int(0.004 * d.size)), int((d > 2.0 / 255.0).sum())
        return got, want, hits, cv

    got, want, hits, cv = both()
    assert abs(cv.cv_isoval - 1.6 * iso / 256.0) < 1e-6 and cv.size == 3.0 and cv.border_scale == 60.0
    hit = hits[0][..., 3] != 0
    # focus from a window position: the first surface's hit under it
    ys, xs = np.nonzero(hit & (hits[2][..., 3] != 0))
    k = len(ys) // 2
    x, row = int(xs[k]), int(ys[k])                         # row counted from the bottom
    r.setCVFocusPos((x, h - row))
    cv2 = r.clearViewParams()
    assert np.allclose(np.array(cv2.pick_pos[:]), hits[0][row, x, :3], atol=2e-4), (cv2.pick_pos[:], hits[0][row, x])
    assert r.checkForRedraw()
    got2, _, hits2, _ = both()
    assert np.abs(got2 - got).max() > 0.05                   # the lens moved
    dist_w = np.sqrt(((hits2[0][..., :3] - np.array(cv2.pick_pos[:], np.float32)) ** 2).sum(axis=2)) * cv2.size
    outside = hit & (dist_w > 1.0 + 1.5 / 60.0)
    inside = hit & (dist_w < 0.5) & (hits2[2][..., 3] != 0)
    assert outside.sum() > 200 and inside.sum() > 50, (int(outside.sum()), int(inside.sum()))
    # curvature can pull the focus colour in outside the lens too; where it does not, the picture is the plain isosurface
    same = outside & (np.abs(got2 - plain).max(axis=2) <= 1e-6)
    assert same.sum() > 0.5 * outside.sum(), (int(same.sum()), int(outside.sum()))
    # inside the lens the focus surface's colour (red) dominates
    assert (got2[inside][:, 0] > got2[inside][:, 2] + 0.05).mean() > 0.8
    # a window position without a hit sends the lens away: the plain isosurface picture wherever the curvature term is quiet
    ys0, xs0 = np.nonzero(~hit[1:])                         # (row 0 is window y == height, which no mouse position has)
    r.setCVFocusPos((int(xs0[0]), h - (int(ys0[0]) + 1)))
    got3, _, _, cv3 = both()
    assert max(abs(v) for v in cv3.pick_pos[:]) > 1e6
    assert (np.abs(got3 - plain).max(axis=2) <= 1e-6).mean() > 0.9
    # ClearView off: the plain picture again, bit for bit; GLRenderer::Pick reads the same hit buffer
    r.setClearViewEnabled(False)
    r.paintUntilConverged()
    assert np.array_equal(r.frameFloat(), plain)
    assert np.allclose(np.array(r.pick((x, h - row))), hits[0][row, x, :3], atol=1e-5)
    with pytest.raises(Exception, match="No intersection"):
        r.pick((int(xs0[0]), h - (int(ys0[0]) + 1)))
    r.setRenderMode(0)
    with pytest.raises(Exception, match="isosurface"):
        r.pick((x, h - row))
    r.cleanup()
