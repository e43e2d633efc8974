"""Host-side camera / uniform computation for the ray-march parameters (product code).

Mirrors what GLRenderer::SetViewPort + ComputeViewAndProjection (Tuvok/Renderer/GL/GLRenderer.cpp:
864-920), GLGridLeaper::ComputeEyeToModelMatrix (GLGridLeaper.cpp:560-573), SetupRaycastShader
(:690-752), CullingLOD::SetScreenParams (Tuvok/Renderer/CullingLOD.cpp:57-67) and
GLVolumePool::Enable (GLVolumePool.cpp:790-803) compute on the host, in fp32 with the row-vector
convention of Tuvok/Basics/Vectors.h (v' = v*M).  The C++ host (csrc/host) carries the same
code for the AbstrRenderer-shaped class; this module serves Python callers (bench.py)."""
import math

import numpy as np

from . import _lib as L

f32 = np.float32


def look_at(eye, at, up):
    """MATRIX4::BuildLookAt, Vectors.h:1250-1265."""
    eye, at, up = [np.asarray(v, f32) for v in (eye, at, up)]
    F = at - eye
    S = np.cross(F, up).astype(f32)
    U = np.cross(S, F).astype(f32)
    F, U, S = [v / f32(np.sqrt(f32(np.dot(v, v)))) for v in (F, U, S)]
    m = np.zeros((4, 4), f32)
    m[0, :3] = [S[0], U[0], -F[0]]
    m[1, :3] = [S[1], U[1], -F[1]]
    m[2, :3] = [S[2], U[2], -F[2]]
    m[3] = [-f32(np.dot(S, eye)), -f32(np.dot(U, eye)), f32(np.dot(F, eye)), 1]
    return m


def perspective(fovy_deg, aspect, n, f):
    """MATRIX4::Perspective, Vectors.h:1267-1277."""
    fovy = f32(f32(fovy_deg) * f32(3.14159265358979323846 / 180.0))
    cotan = f32(1.0 / math.tan(float(fovy) / 2.0))
    aspect, n, f = f32(aspect), f32(n), f32(f)
    m = np.zeros((4, 4), f32)
    m[0, 0], m[1, 1] = cotan / aspect, cotan
    m[2, 2], m[2, 3] = -(f + n) / (f - n), -1
    m[3, 2] = f32(-2) * (f * n) / (f - n)
    return m


def inverse(m):
    return np.linalg.inv(np.asarray(m, np.float64)).astype(f32)


def translation(x, y, z):
    m = np.eye(4, dtype=f32)
    m[3, :3] = [x, y, z]
    return m


def scaling(x, y, z):
    return np.diag(np.asarray([x, y, z, 1], f32))


def rotation_x(a):
    c, s = f32(math.cos(a)), f32(math.sin(a))
    m = np.eye(4, dtype=f32)
    m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, s, -s, c
    return m


def rotation_y(a):
    c, s = f32(math.cos(a)), f32(math.sin(a))
    m = np.eye(4, dtype=f32)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, -s, s, c
    return m


def lod_factor(fovy_deg, pixels_y, sse=1.0):
    arg = f32(f32(fovy_deg) * f32(f32(f32(3.1416) / f32(180.0)) / f32(2.0)))
    return f32(f32(f32(2.0) * f32(math.tan(float(arg)))) * f32(sse) / f32(float(pixels_y)))


class Camera:
    """AbstrRenderer defaults (AbstrRenderer.cpp:64-69,71-172)."""

    def __init__(self):
        self.eye = np.array([0, 0, 1.6], f32)
        self.at = np.array([0, 0, 0], f32)
        self.up = np.array([0, 1, 0], f32)
        self.fov, self.znear, self.zfar = 50.0, 0.01, 1000.0
        self.rotation = np.eye(4, dtype=f32)
        self.translation = np.eye(4, dtype=f32)
        self.light_dir = np.array([0, 0, -1], f32)
        self.ambient = np.array([1, 1, 1, 0.1], f32)
        self.diffuse = np.array([1, 1, 1, 1], f32)
        self.specular = np.array([1, 1, 1, 1], f32)
        self.iso_color = np.array([0.5, 0.5, 0.5], f32)


def raycast_params(cam, width, height, domain_size, scale=(1, 1, 1), render_mode=0, lighting=False,
                   sample_rate=1.0, trans_scale=1.0, gradient_scale=1.0, isoval=0.5, new_view=True,
                   strategy=L.MBS_SKIP_TWO_LEVELS, rehash=10, hash_size=509, count_samples=True, slab=None):
    p = L.RaycastParams()
    p.width, p.height = width, height
    dom = np.asarray(domain_size, f32)
    sc = np.asarray(scale, f32)
    extent = dom * sc
    extent = extent / extent.max()
    view = look_at(cam.eye, cam.at, cam.up)
    proj = perspective(cam.fov, f32(width) / f32(height), cam.znear, cam.zfar)
    mv = (cam.rotation @ cam.translation @ view).astype(f32)
    emm = (inverse(mv) @ scaling(*(f32(1) / extent)) @ translation(0.5, 0.5, 0.5)).astype(f32)
    box_min, box_max = [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]
    wse = float((extent / dom).max())
    emm_global = emm
    pos_scale_m1, pos_offset = [0.0, 0.0, 0.0], [0.0, 0.0, 0.0]
    if slab is not None:
        # sort-last partition: this rank holds voxels [zs, ze) of the global volume (its own slab
        # [z0, z1) plus a halo) as ITS volume; map global normalised coords to the local ones and
        # clip the rays to the slab proper.  domain_size is the GLOBAL size here.
        ld = np.asarray(slab["local_dims"], f32)
        if "lo" in slab:                                    # box partition (slabs are boxes spanning x and y)
            lo3, hi3, s3 = [np.asarray(slab[k], np.float64) for k in ("lo", "hi", "s")]
        else:
            lo3 = np.array([0, 0, slab["z0"]], np.float64)
            hi3 = np.array([dom[0], dom[1], slab["z1"]], np.float64)
            s3 = np.array([0, 0, slab["zs"]], np.float64)
        emm = (emm @ scaling(dom[0] / ld[0], dom[1] / ld[1], dom[2] / ld[2]) @
               translation(-s3[0] / float(ld[0]), -s3[1] / float(ld[1]), -s3[2] / float(ld[2]))).astype(f32)
        for a in range(3):
            box_min[a], box_max[a] = (lo3[a] - s3[a]) / float(ld[a]), (hi3[a] - s3[a]) / float(ld[a])
            # lighting happens in the GLOBAL volume's model space: local position * ld/dom + s/dom
            pos_scale_m1[a], pos_offset[a] = float(ld[a]) / float(dom[a]) - 1.0, float(s3[a]) / float(dom[a])
    p.eye_to_model[:] = emm.reshape(-1).tolist()
    p.model_to_eye[:] = inverse(emm).reshape(-1).tolist()
    p.model_view_it[:] = inverse(mv).T.reshape(-1).tolist()
    p.inv_projection[:] = inverse(proj).reshape(-1).tolist()
    p.lod_factor = float(lod_factor(cam.fov, height))
    p.level_zero_wse = wse
    p.sample_rate_modifier = sample_rate
    p.trans_scale, p.gradient_scale, p.isoval = trans_scale, gradient_scale, isoval
    p.render_mode, p.lighting = render_mode, int(lighting)
    p.light_ambient[:] = (cam.ambient[:3] * cam.ambient[3]).tolist()
    p.light_diffuse[:] = (cam.diffuse[:3] * cam.diffuse[3]).tolist()
    p.light_specular[:] = (cam.specular[:3] * cam.specular[3]).tolist()
    ld = (np.append(cam.light_dir, f32(0)) @ emm_global)[:3]
    p.model_light_dir[:] = (ld / f32(np.sqrt(f32(np.dot(ld, ld))))).tolist()
    p.model_eye_pos[:] = (np.append(cam.eye, f32(1)) @ emm_global)[:3].tolist()
    p.model_pos_scale_m1[:] = pos_scale_m1
    p.model_pos_offset[:] = pos_offset
    p.domain_scale[:] = (f32(1) / (sc / sc.min())).tolist()
    p.iso_color[:] = cam.iso_color.tolist()
    p.compose_light_dir[:] = cam.light_dir.tolist()
    p.new_view = int(new_view)
    p.missing_strategy, p.rehash_count, p.hash_table_size = strategy, rehash, hash_size
    p.clip_enabled = 0
    p.box_min[:] = box_min
    p.box_max[:] = box_max
    p.count_samples = int(count_samples)
    return p
