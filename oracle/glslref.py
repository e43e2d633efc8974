"""TEST INFRASTRUCTURE ONLY -- runs the reference's OWN GLSL ray-casting shaders on the CPU.

oracle/glsl/build_glslref.py compiles the shader text of /root/reference (Tuvok/Shaders/*.glsl plus the strings
GLVolumePool / GLHashTable::GetShaderFragment generate, produced by the reference's own statements) with g++ against
a GLSL value-type shim.  This module binds a scene to it with the same buffers oracle/host.py gives the restated oracle
(tvk_oracle_render.c), so that the restatement can be PINNED: same sub-frame, same pool, same camera -> images, resume
buffers and the request table must agree (tests/test_glsl_reference.py), and frames rendered this way are committed as
golden fixtures for the GPU tests (tests/golden/make_golden_glsl.py), because /root/reference does not exist on the
GPU box.

What is still restated around the shader text: the rasteriser (which pixels the bounding-box faces cover and the
interpolated positions -- the analytic ray/box intersection of the pixel-centre ray, orc_ray_setup_image), the uniform
values (oracle/host.py raycast_params, a restatement of GLGridLeaper::SetupRaycastShader and the camera code) and
GL_LINEAR filtering as the OpenGL specification writes it, in fp32 (real GPUs quantise the weights to 8-9 bits)."""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

from . import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SHADERS = "/root/reference/Tuvok/Shaders"
MAX_3D_TEXTURE = 2048          # GL_MAX_3D_TEXTURE_SIZE / GL_MAX_TEXTURE_SIZE of the hardware the reference targets
MAX_2D_TEXTURE = 16384
_LIBS = {}


def available():
    return os.path.isdir(REF_SHADERS) and os.path.isfile("/root/reference/Tuvok/Renderer/GL/GLVolumePool.cpp")


def _load(pool, strategy, table_size, rehash):
    args = list(pool.vol) + list(pool.max_brick) + [pool.overlap] + list(pool.capacity) + [pool.lod_count, strategy, table_size, rehash,
                                                                                              MAX_3D_TEXTURE, MAX_2D_TEXTURE]
    key = hashlib.sha1(json.dumps(args).encode()).hexdigest()[:16]
    if key not in _LIBS:
        sys.path.insert(0, os.path.join(_HERE, "glsl"))
        import build_glslref as B
        out = os.path.join(B.OUT, key)
        lib = os.path.join(out, "libglslref.so")
        srcs = [os.path.join(_HERE, "glsl", f) for f in os.listdir(os.path.join(_HERE, "glsl"))]
        srcs.append(os.path.join(os.path.dirname(_HERE), "include", "tvk_cuda.h"))
        if not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs if os.path.isfile(s)):
            B.build({"key": key, "args": args})
        _LIBS[key] = (C.CDLL(lib), json.load(open(os.path.join(out, "meta.json"))), out)
    return _LIBS[key]


def shader_text(pool, params):
    """(pool fragment, hash fragment) GLSL the reference generates for this pool -- for inspection / tests"""
    _, _, out = _load(pool, int(params.missing_strategy), int(params.hash_table_size), int(params.rehash_count))
    return open(os.path.join(out, "pool_fragment.glsl")).read(), open(os.path.join(out, "hash_fragment.glsl")).read()


def raycast(params, pool, tf_rgba8, frame, rect=None, clear_hash=True):
    """One sub-frame through the reference's shaders; same contract as oracle.host.raycast (frame.stats is not filled)."""
    lib, meta, _ = _load(pool, int(params.missing_strategy), int(params.hash_table_size), int(params.rehash_count))
    w, h = int(params.width), int(params.height)
    entry = np.zeros((h, w, 4), np.float32)
    exit_eye = np.zeros((h, w, 3), np.float32)
    covered = np.zeros((h, w), np.uint8)
    O.lib().orc_ray_setup_image(C.byref(params), entry.ctypes.data_as(C.c_void_p), exit_eye.ctypes.data_as(C.c_void_p),
                                covered.ctypes.data_as(C.c_void_p))
    if tf_rgba8 is None:
        tf = np.zeros((1, 4), np.uint8)
    else:
        tf = np.ascontiguousarray(tf_rgba8, np.uint8)
    tf_w, tf_h = (tf.shape[0], 1) if tf.ndim == 2 else (tf.shape[1], tf.shape[0])
    mt = meta["meta_tex"]
    assert meta["total_bricks"] == pool.total
    meta_tex = np.zeros(mt[0] * mt[1] * mt[2], np.uint32)
    meta_tex[:pool.total] = pool.meta
    ht = meta["hash_tex"]
    assert ht[0] * ht[1] >= frame.hash.size
    hash_tex = np.zeros(ht[0] * ht[1], np.uint32)
    hash_tex[:frame.hash.size] = frame.hash
    dims = [pool.capacity[a] * pool.max_brick[a] for a in range(3)]
    x0, y0, x1, y1 = (0, 0, w, h) if rect is None else rect
    u3 = C.c_uint32 * 3
    rc = lib.glslref_raycast_rect(C.byref(params), pool.data.ctypes.data_as(C.c_void_p), C.c_uint32(pool.bpv), u3(*dims),
                                  meta_tex.ctypes.data_as(C.c_void_p), u3(*mt), tf.ctypes.data_as(C.c_void_p), C.c_uint32(tf_w),
                                  C.c_uint32(tf_h), entry.ctypes.data_as(C.c_void_p), exit_eye.ctypes.data_as(C.c_void_p),
                                  covered.ctypes.data_as(C.c_void_p), frame.pos.ctypes.data_as(C.c_void_p),
                                  frame.aux.ctypes.data_as(C.c_void_p), frame.image.ctypes.data_as(C.c_void_p),
                                  hash_tex.ctypes.data_as(C.c_void_p), (C.c_uint32 * 2)(*ht), C.c_uint32(x0), C.c_uint32(y0),
                                  C.c_uint32(x1), C.c_uint32(y1), C.c_int(1 if clear_hash else 0))
    assert rc == 0
    frame.hash[:] = hash_tex[:frame.hash.size]
    return frame.image
