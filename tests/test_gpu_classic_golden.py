"""GPU: the CUDA classic ray-caster (tvk_raycast_classic / tvk_raycast_classic_cv, kernel-level ABI) against frames rendered by
THE REFERENCE'S OWN SHADERS (tests/golden/classic_frames.npz, written by tests/golden/make_golden_glsl_classic.py from
GLRaycaster-*-FS.glsl, Compose-FS.glsl, Compose-CV-FS.glsl through oracle/classicref.py).  The scene -- volume, bricks, frame
plan -- is rebuilt here from tests/classic_scenes.py (no reference needed); the brick store is the oracle's tight one,
uploaded as it is."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import host as H
from tests import classic_scenes as S

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "classic_frames.npz")


def _to_product(p, bricks, cv):
    from imagevis3d_b200 import _lib as L
    q = L.ClassicParams()
    assert C.sizeof(q) == C.sizeof(p)
    C.memmove(C.byref(q), C.byref(p), C.sizeof(p))
    arr = (L.ClassicBrick * len(bricks))()
    for i, b in enumerate(bricks):
        assert C.sizeof(arr[i]) == C.sizeof(b)
        C.memmove(C.byref(arr[i]), C.byref(b), C.sizeof(b))
    c = None
    if cv is not None:
        c = L.ClearViewParams()
        C.memmove(C.byref(c), C.byref(cv), C.sizeof(cv))
    return q, arr, c


@pytest.mark.parametrize("name", S.NAMES)
def test_cuda_classic_kernel_against_reference_shader_frames(name):
    import torch
    from imagevis3d_b200.context import TvkContext
    gold = np.load(GOLDEN)
    want = gold[name + "/image"]
    p, bricks, store, tf, cv = S.build(name)
    q, arr, c = _to_product(p, bricks, cv)
    dev_store = torch.from_numpy(np.ascontiguousarray(store).view(np.uint8).copy()).cuda()
    tvk = TvkContext(0)
    if p.render_mode == 0:
        tvk.set_tf1d(tf)
    elif p.render_mode == 1:
        tvk.set_tf2d(tf)
    lib = tvk.lib
    if cv is None:
        rc = lib.tvk_raycast_classic(tvk.h, C.byref(q), arr, C.c_uint32(len(bricks)), C.c_void_p(dev_store.data_ptr()))
    else:
        rc = lib.tvk_raycast_classic_cv(tvk.h, C.byref(q), C.byref(c), arr, C.c_uint32(len(bricks)), C.c_void_p(dev_store.data_ptr()))
    assert rc == 0, lib.tvk_last_error(tvk.h)
    got = np.zeros((p.height, p.width, 4), np.float32)
    assert lib.tvk_read_image(tvk.h, got.ctypes.data_as(C.c_void_p)) == 0
    d = np.abs(got - want).max(axis=2)
    print("%s: max |diff| %.3g, PSNR %.1f dB" % (name, d.max(), H.psnr(got, want)))
    assert want[..., 3].max() > 0.3
    if p.render_mode == 2:
        # a hit / miss flip on a silhouette pixel is a full-scale difference (and ClearView's curvature term reads neighbours)
        assert (d > 2.0 / 255.0).sum() <= max(4, d.size // 500), int((d > 2.0 / 255.0).sum())
    else:
        assert d.max() <= 2.0 / 255.0 and H.psnr(got, want) >= 45.0, (d.max(), H.psnr(got, want))
    if cv is not None:
        hit = np.zeros(4, np.float32)
        ys, xs = np.nonzero(gold[name + "/hit_mask"])
        assert lib.tvk_read_iso_hit(tvk.h, C.c_uint32(int(xs[0])), C.c_uint32(int(ys[0])), hit.ctypes.data_as(C.c_void_p)) == 0
        assert hit[3] != 0
    tvk.close()
