"""Scenes of the classic per-brick ray-caster shared by tests/test_glsl_classic_reference.py (C restatement vs the reference's
shader text, CPU), tests/golden/make_golden_glsl_classic.py (frames rendered BY the reference's shaders) and
tests/test_gpu_classic_golden.py (CUDA kernel vs those frames, GPU box).  Everything here runs without /root/reference."""
import numpy as np

from oracle import host as H
from oracle import oracle as O

W, Hh = 96, 72


def _tf(n):
    t = H.tf1d_std(n, 0.3, 0.3)
    t[:, 3] *= 0.15
    t[:, 2] *= 0.5
    return t


def scene(dtype, mode, lighting, clip=None):
    vol = H.synth_blobs((70, 61, 47), dtype)
    oc = O.build_octree(vol, 32, 2)
    cam = H.Camera()
    cam.rotation = H.rotation_y(0.6) @ H.rotation_x(0.3)
    n = 256 if dtype == np.uint8 else 4096
    maxv = float(oc["minmax"][:int(np.prod(oc["plan"].lod_bricks[0][:])), 1].max())
    bits = 8 * np.dtype(dtype).itemsize
    kw = dict(lighting=lighting, trans_scale=((1 << bits) - 1) / maxv, lod=0, render_mode=mode, clip_plane=clip)
    if mode == 0:
        tf = H.tf_bytes(_tf(n))
        lo, hi = H.tf_nonzero_limits(_tf(n))
        p, bricks, _ = H.classic_plan(oc, cam, W, Hh, (lo, hi, n), **kw)
    elif mode == 1:
        v, g = np.linspace(0, 1, n, dtype=np.float32), np.linspace(0, 1, 256, dtype=np.float32)
        t2 = np.zeros((256, n, 4), np.float32)
        t2[..., 0] = v[None, :]; t2[..., 1] = 0.3 + 0.7 * g[:, None]; t2[..., 2] = 0.4          # noqa: E702
        t2[..., 3] = np.clip((v[None, :] - 0.2) * 0.3, 0, 1) * (0.5 + 0.5 * g[:, None])
        tf = (np.clip(t2, 0, 1) * np.float32(255)).astype(np.uint8)
        p, bricks, _ = H.classic_plan(oc, cam, W, Hh, (int(0.2 * n), n - 1, n), gradient_scale=4.0, **kw)
    else:
        tf = np.zeros((n, 4), np.uint8)
        p, bricks, _ = H.classic_plan(oc, cam, W, Hh, (0, n - 1, n), iso_value=0.15 * maxv, **kw)
    return p, bricks, oc["store"], tf


CASES = [("u8_1d", np.uint8, 0, False, None), ("u8_1d_lit", np.uint8, 0, True, None), ("u16_1d_lit", np.uint16, 0, True, None),
         ("u8_2d", np.uint8, 1, False, None), ("u16_2d_lit", np.uint16, 1, True, None), ("u8_iso", np.uint8, 2, False, None),
         ("u16_iso", np.uint16, 2, False, None), ("u8_1d_clip", np.uint8, 0, True, (0.3, 0.2, 1.0, 0.05))]




def build(name):
    """-> (params, bricks, tight brick store, transfer function bytes, ClearView parameters or None)"""
    for n, dtype, mode, lighting, clip in CASES:
        if n == name:
            p, bricks, store, tf = scene(dtype, mode, lighting, clip)
            return p, bricks, store, tf, None
    if name.startswith("clearview_"):
        dtype = np.uint8 if name.endswith("u8") else np.uint16
        p, bricks, store, tf = scene(dtype, 2, False)
        cv = H.ClearViewParams()
        cv.cv_isoval = 1.6 * p.isoval
        cv.cv_color[:] = [1.0, 0.2, 0.1]
        cv.size, cv.context_scale, cv.border_scale = 3.0, 1.0, 60.0
        _, ha, _ = H.classic_render_cv(p, cv, bricks, store)
        ys, xs = np.nonzero((ha[0][..., 3] != 0) & (ha[2][..., 3] != 0))
        assert len(ys) > 30, "no pixel shows both surfaces: vacuous"
        cv.pick_pos[:] = ha[0][ys[len(ys) // 2], xs[len(ys) // 2], :3].tolist()       # the lens on the surface
        return p, bricks, store, tf, cv
    raise KeyError(name)


NAMES = [c[0] for c in CASES] + ["clearview_u8", "clearview_u16"]
