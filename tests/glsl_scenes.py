"""Scenes shared by the three legs of the image-parity chain (test infrastructure):

  tests/test_glsl_reference.py     restated oracle  ==  the reference's own GLSL executed on the CPU   (needs /root/reference)
  tests/golden/make_golden_glsl.py frames + request sets rendered by the reference's GLSL -> tests/golden/glsl_frames.npz
  tests/test_gpu_glsl_golden.py    CUDA path  vs  those committed frames                                (GPU box, no reference)
"""
import numpy as np

from oracle import host as H
from oracle import oracle as O


def _tf(n, center=0.3, inv=0.3, alpha=0.15):
    t = H.tf1d_std(n, center, inv)
    t[:, 3] *= alpha
    t[:, 0] *= 0.9
    t[:, 2] *= 0.6
    return t


def _tf2d(n):
    v = np.linspace(0, 1, n, dtype=np.float32)
    g = np.linspace(0, 1, 256, dtype=np.float32)
    t = np.zeros((256, n, 4), np.float32)                       # [gradient row][value]
    t[..., 0] = v[None, :]
    t[..., 1] = g[:, None]
    t[..., 2] = 0.5
    t[..., 3] = np.clip((v[None, :] - 0.2) * 0.6, 0, 1) * (0.3 + 0.7 * g[:, None])
    return (np.clip(t, 0, 1) * np.float32(255)).astype(np.uint8)


# name -> description.  vol: (dims, dtype, seed); cam: list of ("x"|"y", angle) rotations applied left to right
SCENES = {
    "u8_1d":        dict(vol=((150, 131, 97), "u8", 3), brick=20, cap=(8, 8, 8), mode=0, lit=False, size=(163, 119), cam=[("y", 0.6), ("x", 0.3)]),
    "u8_1d_lit":    dict(vol=((150, 131, 97), "u8", 3), brick=20, cap=(8, 8, 8), mode=0, lit=True, size=(163, 119), cam=[("y", 0.6), ("x", 0.3)], eye=1.05),
    "u16_1d_rate":  dict(vol=((120, 100, 90), "u16", 9), brick=20, cap=(7, 7, 7), mode=0, lit=False, size=(160, 120), cam=[("y", -0.9)],
                         sample_rate=1.7, tf=(0.25, 0.2, 0.3)),
    "u16_2d":       dict(vol=((120, 100, 90), "u16", 9), brick=20, cap=(7, 7, 7), mode=1, lit=False, size=(144, 112), cam=[("x", 0.5), ("y", 0.4)]),
    "u16_2d_lit":   dict(vol=((120, 100, 90), "u16", 9), brick=20, cap=(7, 7, 7), mode=1, lit=True, size=(144, 112), cam=[("x", 0.5), ("y", 0.4)], eye=1.0),
    "u8_iso":       dict(vol=((150, 131, 97), "u8", 3), brick=20, cap=(8, 8, 8), mode=2, lit=False, size=(180, 140), cam=[("y", 2.1), ("x", -0.4)],
                         iso=100.0, eye=1.1),
    "u8_1d_clip":   dict(vol=((150, 131, 97), "u8", 3), brick=20, cap=(8, 8, 8), mode=0, lit=True, size=(150, 110), cam=[("y", 0.9), ("x", -0.2)],
                         clip=(0.6, 0.3, 0.74, 0.05)),
    "u8_1d_reqall": dict(vol=((150, 131, 97), "u8", 3), brick=20, cap=(8, 8, 8), mode=0, lit=False, size=(140, 100), cam=[("y", 3.9)],
                         strategy=1, hash_size=65521),          # > GL_MAX_TEXTURE_SIZE: the 2-D hash-table texture path
    "u8_1d_opaque": dict(vol=((150, 131, 97), "u8", 3), brick=20, cap=(8, 8, 8), mode=0, lit=False, size=(128, 96), cam=[],
                         tf=(0.1, 0.1, 1.0)),
    "u16_1d_lit_aniso": dict(vol=((100, 120, 70), "u16", 5), brick=24, cap=(6, 6, 6), mode=0, lit=True, size=(150, 120),
                             cam=[("y", 1.3), ("x", 0.6)], scale=(1.0, 1.0, 2.5), tf=(0.25, 0.25, 0.25)),
    "u8_1d_small_pool": dict(vol=((150, 131, 97), "u8", 3), brick=20, cap=(4, 4, 3), mode=0, lit=False, size=(120, 90), cam=[("y", 0.6), ("x", 0.3)]),
}

_OCT = {}


def octree(vol_desc, brick):
    key = (vol_desc, brick)
    if key not in _OCT:
        dims, dt, seed = vol_desc
        _OCT[key] = O.build_octree(H.synth_blobs(dims, np.uint8 if dt == "u8" else np.uint16, seed=seed), brick, 2)
    return _OCT[key]


def build(name):
    """-> dict(oc, pool (cold: only the coarsest brick resident, visibility computed), params, tf (bytes or None))"""
    s = SCENES[name]
    oc = octree(s["vol"], s["brick"])
    bits = 8 if s["vol"][1] == "u8" else 16
    n = 256 if bits == 8 else 4096
    maxv = float(oc["minmax"][:, 1].max())
    tmax = float((1 << bits) - 1)
    kw = dict(render_mode=s["mode"], lighting=s["lit"], sample_rate=s.get("sample_rate", 1.0), strategy=s.get("strategy", 3),
              hash_size=s.get("hash_size", 509), scale=s.get("scale", (1, 1, 1)))
    if "clip" in s:
        kw["clip_plane"] = s["clip"]
    pool = H.OraclePool(oc, s["cap"])
    pool.upload_first_brick()
    if s["mode"] == 0:
        tf = _tf(n, *s.get("tf", (0.3, 0.3, 0.15)))
        lo, hi = H.tf_nonzero_limits(tf)
        resc = maxv / float(n - 1) if bits == 16 else 1.0
        if bits == 16:
            kw["trans_scale"] = tmax / maxv
        pool.recompute_visibility(0, lo * resc, hi * resc, -1e308, 1e308)
        tfb = H.tf_bytes(tf)
    elif s["mode"] == 1:
        tfb = _tf2d(256)
        kw["trans_scale"] = tmax / maxv
        kw["gradient_scale"] = 8.0
        pool.recompute_visibility(1, 0.2 * maxv, maxv, -1e308, 1e308)
    else:
        kw["isoval"] = s["iso"] / float(1 << bits)
        pool.recompute_visibility(2, s["iso"], 0.0, -1e308, 1e308)
        tfb = None
    cam = H.Camera()
    if "eye" in s:
        cam.eye = np.array([0, 0, s["eye"]], np.float32)
    rot = np.eye(4, dtype=np.float32)
    for axis, ang in s["cam"]:
        rot = rot @ (H.rotation_y(ang) if axis == "y" else H.rotation_x(ang))
    cam.rotation = rot
    w, h = s["size"]
    params = H.raycast_params(cam, w, h, oc["vol"], **kw)
    return dict(oc=oc, pool=pool, params=params, tf=tfb)


def render_progressive(raycast_fn, scene, max_subframes=100):
    """paint() until converged with `raycast_fn` (oracle.host.raycast or oracle.glslref.raycast) marching and the oracle's
    pool manager paging on ITS requests.  -> dict(image, pos, aux, requests: list of sorted request arrays, subframes)"""
    p, pool, tf = scene["params"], scene["pool"], scene["tf"]
    frame = H.OracleFrame(p.width, p.height, p.hash_table_size)
    tfa = tf if tf is not None else np.zeros((1, 4), np.uint8)
    p.new_view = 1
    reqs, hashes = [], []
    for sub in range(max_subframes):
        raycast_fn(p, pool, tfa, frame)
        p.new_view = 0
        r = H.read_requests(pool, frame, cap=max(4096, int(p.hash_table_size)))
        reqs.append(r.copy())
        hashes.append(frame.hash.copy())
        if len(r) == 0:
            return dict(image=frame.image.copy(), pos=frame.pos.copy(), aux=frame.aux.copy(), requests=reqs, hashes=hashes, subframes=sub + 1)
        pool.upload_requests(r)
    raise AssertionError("did not converge")
