"""TEST INFRASTRUCTURE ONLY -- runs the reference's OWN classic ray-caster shaders (GLRaycaster-*-FS.glsl, Volume3D.glsl,
VRender1D*.glsl, lighting.glsl, Compositing.glsl, RefineIsosurface.glsl, Compose-FS.glsl, Compose-CV-FS.glsl) on the CPU.

oracle/glsl/build_classicref.py compiles the shader text of /root/reference with g++ against the GLSL value-type shim
(oracle/glsl/glsl_runtime.h); this module binds a frame plan (parameters + front-to-back brick list + brick store) to it, the
same buffers oracle.host.classic_render gives the restated oracle (tvk_oracle_classic.c), so that the restatement can be
PINNED to the reference's shader text (tests/test_glsl_classic_reference.py).  Needs /root/reference."""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def available():
    return os.path.isfile("/root/reference/Tuvok/Shaders/GLRaycaster-1D-FS.glsl")


def _lib():
    global _LIB
    if _LIB is None:
        sys.path.insert(0, os.path.join(_HERE, "glsl"))
        import build_classicref as B
        lib = os.path.join(B.OUT, "libclassicref.so")
        srcs = [os.path.join(_HERE, "glsl", f) for f in os.listdir(os.path.join(_HERE, "glsl"))]
        srcs.append(os.path.join(os.path.dirname(_HERE), "include", "tvk_cuda.h"))
        if not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs if os.path.isfile(s)):
            B.build()
        _LIB = C.CDLL(lib)
    return _LIB


def render(p, bricks, store_tight, tf_rgba8=None, cv=None):
    """One sub-frame of the classic ray-caster through the reference's shaders.  p / bricks / cv: the ctypes structures of
    oracle.host (ClassicParams, ClassicBrick, ClearViewParams).  -> image [H][W][4] (and the four hit targets when cv is given)"""
    from . import host as H
    arr = (H.ClassicBrick * max(len(bricks), 1))(*bricks)
    img = np.zeros((p.height, p.width, 4), np.float32)
    tf = np.zeros((1, 4), np.uint8) if tf_rgba8 is None else np.ascontiguousarray(tf_rgba8, np.uint8)
    tf_w, tf_h = (tf.shape[0], 1) if tf.ndim == 2 else (tf.shape[1], tf.shape[0])
    st = np.ascontiguousarray(store_tight, np.uint8)
    hits = np.zeros((4, p.height, p.width, 4), np.float32) if cv is not None else None
    rc = _lib().classicref_render(C.byref(p), C.byref(cv) if cv is not None else None, arr, C.c_uint32(len(bricks)),
                                  st.ctypes.data_as(C.c_void_p), tf.ctypes.data_as(C.c_void_p), C.c_uint32(tf_w), C.c_uint32(tf_h),
                                  img.ctypes.data_as(C.c_void_p), hits.ctypes.data_as(C.c_void_p) if hits is not None else None)
    if rc:
        raise ValueError("classicref_render failed")
    return (img, hits) if cv is not None else img


def slice_view(p, brick, tf_rgba8):
    """1D-slice-FS / 2D-slice-FS / MIP-slice-FS + Transfer-MIP-FS on the quad GLRenderer::RenderSlice draws.
    p: a ctypes tvk_slice_params (same layout as include/tvk_cuda.h); brick [z][y][x] uint8 / uint16"""
    img = np.zeros((p.height, p.width, 4), np.float32)
    tf = np.ascontiguousarray(tf_rgba8, np.uint8)
    tf_w, tf_h = (tf.shape[0], 1) if tf.ndim == 2 else (tf.shape[1], tf.shape[0])
    br = np.ascontiguousarray(brick)
    rc = _lib().classicref_slice(C.byref(p), br.ctypes.data_as(C.c_void_p), tf.ctypes.data_as(C.c_void_p), C.c_uint32(tf_w), C.c_uint32(tf_h),
                                 img.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError("classicref_slice failed")
    return img


def hq_mip(p, bricks, store_tight, tf_rgba8):
    """GLRaycaster-MIP-Rot-FS per brick under GL_MAX + Transfer-MIP-FS.  p: a ctypes tvk_mip_params"""
    from . import host as H
    arr = (H.ClassicBrick * max(len(bricks), 1))(*bricks)
    img = np.zeros((p.height, p.width, 4), np.float32)
    tf = np.ascontiguousarray(tf_rgba8, np.uint8).reshape(-1, 4)
    st = np.ascontiguousarray(store_tight, np.uint8)
    rc = _lib().classicref_mip(C.byref(p), arr, C.c_uint32(len(bricks)), st.ctypes.data_as(C.c_void_p), tf.ctypes.data_as(C.c_void_p),
                               C.c_uint32(len(tf)), img.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError("classicref_mip failed")
    return img


def stereo_compose(left, right, mode, split=0.5, af_id=0):
    """Compose-Anaglyphs-FS (0) / Compose-Scanline-FS (1) / Compose-SBS-FS (2) / Compose-AF-FS (3) over two eye pictures"""
    l, r = np.ascontiguousarray(left, np.float32), np.ascontiguousarray(right, np.float32)
    h, w = l.shape[:2]
    img = np.zeros((h, w, 4), np.float32)
    rc = _lib().classicref_stereo(C.c_uint32(mode), l.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), C.c_uint32(w), C.c_uint32(h),
                                  C.c_float(split), C.c_int(af_id), img.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError("classicref_stereo failed")
    return img
