"""stereo rendering.  CPU: the eye matrices of the host (tvkr_stereo_matrices) and of the oracle's numpy
restatement against the reference's own FLOATMATRIX4::BuildStereoLookAtAndProjection (golden file written by
tests/golden/make_golden_stereo.py from oracle/_ref/stereo_ref, and the program itself where it is built).  GPU: both
eyes through tvkr_paint against the oracle marching the two views in lock step (one request table), and the four compose
modes of GLRenderer::EndFrame against the restated compose shaders."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from imagevis3d_b200 import _lib as L
from oracle import host as H
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "stereo_matrices.json")))
STEREO_REF = os.path.join(HERE, "..", "oracle", "_ref", "stereo_ref")


def _product(args):
    f3 = lambda v: (C.c_float * 3)(*v)                                # noqa: E731
    out = (C.c_float * 64)()
    fn = L.load().tvkr_stereo_matrices
    fn.restype = C.c_int
    rc = fn(f3(args[0:3]), f3(args[3:6]), f3(args[6:9]), *[C.c_float(v) for v in args[9:15]], out)
    assert rc == 0
    return np.array(out[:], np.float32).reshape(4, 16)


def _oracle(args):
    cam = H.Camera()
    cam.eye, cam.at, cam.up = [np.array(args[i:i + 3], np.float32) for i in (0, 3, 6)]
    cam.fov, cam.znear, cam.zfar = args[9], args[11], args[12]
    (vl, pl), (vr, pr) = H.stereo_matrices(cam, np.float32(args[10]), args[13], args[14])
    return np.stack([m.reshape(-1) for m in (vl, vr, pl, pr)])


@pytest.mark.parametrize("case", range(len(GOLDEN)))
def test_stereo_matrices_equal_the_reference(case):
    g = GOLDEN[case]
    want = np.array([g["view_left"], g["view_right"], g["proj_left"], g["proj_right"]], np.float32)
    got = _product(g["args"])
    assert np.array_equal(got, want), np.abs(got - want).max()
    ora = _oracle(g["args"])
    assert np.allclose(ora, want, rtol=2e-6, atol=1e-7), np.abs(ora - want).max()
    # the two eyes: the same rotation, translations that differ, projections skewed by +-shift (off-axis, not toed in)
    assert np.array_equal(want[0][:12], want[1][:12]) and not np.array_equal(want[0][12:], want[1][12:])
    assert want[2][8] == -want[3][8] and want[2][8] > 0


@pytest.mark.skipif(not os.path.isfile(STEREO_REF), reason="oracle/_ref/stereo_ref is not built (/root/reference absent)")
def test_stereo_matrices_random_cameras_against_the_compiled_reference():
    rng = np.random.default_rng(5)
    for _ in range(40):
        a = np.concatenate([rng.uniform(-2, 2, 3), rng.uniform(-0.3, 0.3, 3), [0, 1, 0] + rng.uniform(-0.2, 0.2, 3), [rng.uniform(20, 90)],
                            [rng.uniform(0.5, 2.5)], [rng.uniform(0.005, 0.2)], [rng.uniform(10, 2000)], [rng.uniform(0.5, 3)],
                            [rng.uniform(0.005, 0.1)]]).astype(np.float32)
        txt = subprocess.check_output([STEREO_REF] + [repr(float(v)) for v in a]).decode()
        want = np.array([[float(t) for t in line.split()] for line in txt.strip().splitlines()], np.float32)
        got = _product([float(v) for v in a])
        # (equal to the last bit on most cameras; the matrix product's summation order may differ by an ulp)
        assert np.allclose(got, want, rtol=5e-7, atol=1e-7), (a, np.abs(got - want).max())


def test_compose_restatement_modes():
    """the numpy restatement of the four compose shaders on a 6x4 picture (what the GPU test compares the kernel with)"""
    rng = np.random.default_rng(2)
    l, r = rng.random((4, 6, 4), np.float32), rng.random((4, 6, 4), np.float32)
    rb = H.stereo_compose(l, r, 0)
    assert rb[2, 3, 0] == pytest.approx(float(l[2, 3, 0] * 0.3 + l[2, 3, 1] * 0.59 + l[2, 3, 2] * 0.11), rel=1e-6)
    assert rb[2, 3, 2] == pytest.approx(2 * float(rb[2, 3, 1]), rel=1e-6) and rb[2, 3, 3] == max(l[2, 3, 3], r[2, 3, 3])
    sc = H.stereo_compose(l, r, 1)
    assert np.array_equal(sc[0], l[0]) and np.array_equal(sc[1], r[1]) and np.array_equal(sc[2], l[2])
    sbs = H.stereo_compose(l, r, 2)
    assert np.array_equal(sbs[:, :3], l[:, [1, 3, 5]]) and np.array_equal(sbs[:, 3:], r[:, [1, 3, 5]])
    assert np.array_equal(H.stereo_compose(l, r, 3, af_id=0), l) and np.array_equal(H.stereo_compose(l, r, 3, af_id=1), r)
    assert np.array_equal(H.stereo_compose(l, r, 3, swap=True), r)


def _tf(n):
    tf = H.tf1d_std(n, 0.3, 0.3)
    tf[:, 3] *= 0.15
    return tf


@pytest.mark.gpu
@pytest.mark.parametrize("lighting", [False, True])
def test_stereo_pair_matches_the_oracle_in_lock_step(lighting):
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    from tests.util import image_errors
    vol = H.synth_blobs((150, 131, 97), np.uint8)
    r = Renderer(0)
    r.loadRawVolume(vol, brick_size=20)
    info = r.datasetInfo()
    tf = _tf(info.tf_size)
    r.set1DTransFloat(tf)
    r.setLightingEnabled(lighting)
    w, h = 203, 149
    r.resize((w, h))
    r.setRendererTarget(RT_HEADLESS)
    rot = H.rotation_y(0.6) @ H.rotation_x(0.3)
    r.setRotation4x4(rot)
    # mono frame first: what stereo must NOT be
    r.paintUntilConverged()
    mono = r.frameFloat()
    r.setStereoEnabled(True)
    r.setStereoEyeDist(0.05)
    r.setStereoFocalLength(1.3)
    r.setStereoMode("af")
    assert r.getStereoEnabled() and r.checkForRedraw()
    r.clearWorkingSet()                                        # the oracle's pool starts from the pinned coarsest brick too
    # oracle: the same two views through the restated shader, one pool, one request table
    cam = H.Camera()
    cam.rotation = rot
    oc = O.build_octree(vol, 20, 2)
    pool = H.OraclePool(oc, tuple(info.pool_capacity))
    pool.upload_first_brick()
    maxv = float(oc["minmax"][:int(np.prod(oc["plan"].lod_bricks[0][:])), 1].max())
    lo, hi = H.tf_nonzero_limits(tf)
    pool.recompute_visibility(0, lo * maxv / (len(tf) - 1), hi * maxv / (len(tf) - 1))
    eyes = H.stereo_matrices(cam, np.float32(w) / np.float32(h), 1.3, 0.05)
    po = [H.raycast_params(cam, w, h, oc["vol"], trans_scale=255.0 / maxv, lighting=lighting, view_proj=eyes[e]) for e in (0, 1)]
    for e in (0, 1):
        pc = r.raycastParams(e)
        for name in ("eye_to_model", "model_to_eye", "model_view_it", "inv_projection", "model_eye_pos", "model_light_dir"):
            a, b = np.array(getattr(pc, name)[:], np.float64), np.array(getattr(po[e], name)[:], np.float64)
            assert np.allclose(a, b, rtol=3e-5, atol=3e-6), (e, name, np.abs(a - b).max())
        assert pc.lod_factor == pytest.approx(po[e].lod_factor, rel=1e-6)
    assert not np.allclose(np.array(r.raycastParams(0).eye_to_model[:]), np.array(r.raycastParams(1).eye_to_model[:]))
    want_l, want_r, subs_o = H.render_converged_stereo(po[0], po[1], pool, H.tf_bytes(tf))
    subs = r.paintUntilConverged()
    assert r.converged() and subs == subs_o, (subs, subs_o)
    got_l = r.frameFloat()                                     # alternating-frame mode, frame id 0: the left eye as it is
    r.toggleStereoFrame()
    got_r = r.frameFloat()
    assert np.array_equal(got_l, r.eyeFrameFloat(0)) and np.array_equal(got_r, r.eyeFrameFloat(1))
    for got, want in ((got_l, want_l), (got_r, want_r)):
        psnr, err = image_errors(got, want)
        assert psnr >= 45.0 and err <= 2.0 / 255.0, (psnr, err)
    assert np.abs(got_l - got_r).max() > 0.05 and np.abs(got_l - mono).max() > 0.01
    bricks, _ = r.poolSlots()
    assert set(bricks[bricks >= 0].tolist()) == set(pool.slot_brick[pool.slot_brick >= 0].tolist())
    # the other compose modes are functions of the two eye pictures: bit for bit where they copy, to rounding where they weigh
    r.initStereoFrame()
    for mode in (0, 1, 2, 3):
        for swap in (False, True):
            r.setStereoMode(mode)
            r.setStereoEyeSwap(swap)
            assert r.checkForRedraw()
            r.paintUntilConverged()
            got = r.frameFloat()
            ref = H.stereo_compose(r.eyeFrameFloat(0), r.eyeFrameFloat(1), mode, swap)   # the eye pictures of THIS frame (warm pool)
            if mode == 0:
                assert np.abs(got - ref).max() <= 1e-6
            else:
                assert np.array_equal(got, ref), (mode, swap)
    # the captured frame is the composed one; switching stereo off gives the mono frame again
    r.setStereoMode("sbs")
    r.setStereoEyeSwap(False)
    r.paintUntilConverged()
    f = r.frameFloat()
    rgba = r.frameRGBA8()
    own8 = (np.clip(f[::-1], np.float32(0), np.float32(1)) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    assert np.array_equal(rgba, own8)
    r.setStereoEnabled(False)
    r.paintUntilConverged()
    psnr, err = image_errors(r.frameFloat(), mono)             # (warm pool now: rays resume at other places than in the cold frame)
    assert psnr >= 45.0 and err <= 2.0 / 255.0, (psnr, err)
    r.cleanup()


@pytest.mark.gpu
def test_select_eye_and_compose_through_the_kernel_abi():
    """tvk_select_eye / tvk_compose_stereo misuse: composing without a second eye, bad mode, sizes that differ"""
    lib = L.load()
    from imagevis3d_b200.context import TvkContext
    tvk = TvkContext(0)
    assert lib.tvk_compose_stereo(tvk.h, C.c_uint32(0), C.c_uint32(0), C.c_float(0.5), C.c_uint32(0)) != 0
    assert lib.tvk_select_eye(tvk.h, C.c_uint32(2)) != 0
    assert lib.tvk_select_eye(tvk.h, C.c_uint32(1)) == 0 and lib.tvk_select_eye(tvk.h, C.c_uint32(0)) == 0
    assert lib.tvk_compose_stereo(tvk.h, C.c_uint32(7), C.c_uint32(0), C.c_float(0.5), C.c_uint32(0)) != 0
    tvk.close()
