"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's HOST side of the
GLGridLeaper path in numpy (fp32 where the reference uses float):

  look_at / perspective / inverse     Tuvok/Basics/Vectors.h:1250-1277,1042-1140
  camera -> raycast parameters        Tuvok/Renderer/GL/GLRenderer.cpp:864-920 (SetViewPort,
                                      ComputeViewAndProjection), GLGridLeaper.cpp:560-573
                                      (ComputeEyeToModelMatrix), :690-752 (SetupRaycastShader)
  lod_factor                          Tuvok/Renderer/CullingLOD.cpp:57-67
  pool sizing                         Tuvok/Renderer/GPUMemMan/GPUMemMan.cpp:766-844
  OraclePool (slots, LRU, metadata)   Tuvok/Renderer/GL/GLVolumePool.cpp:673-778,955-959,
                                      992-1018,1720-1789; GLVolumePool.h:25-60
  render loop                         GLGridLeaper::Render3DRegion (GLGridLeaper.cpp:914-1154)

Used by tests/ (and smoke()/bench cpu legs) to drive liboracle.so; never imported by
the product package.
"""
import ctypes as C
import math

import numpy as np

from . import oracle as O

f32 = np.float32

# ---------------------------------------------------------------- ctypes mirrors of include/tvk_cuda.h
TVK_MAX_LOD = 16


class RaycastParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32),
                ("eye_to_model", C.c_float * 16), ("model_to_eye", C.c_float * 16),
                ("model_view_it", C.c_float * 16), ("inv_projection", C.c_float * 16),
                ("lod_factor", C.c_float), ("level_zero_wse", C.c_float), ("sample_rate_modifier", C.c_float),
                ("trans_scale", C.c_float), ("gradient_scale", C.c_float), ("isoval", C.c_float),
                ("render_mode", C.c_uint32), ("lighting", C.c_uint32),
                ("light_ambient", C.c_float * 3), ("light_diffuse", C.c_float * 3), ("light_specular", C.c_float * 3),
                ("model_light_dir", C.c_float * 3), ("model_eye_pos", C.c_float * 3), ("domain_scale", C.c_float * 3),
                ("iso_color", C.c_float * 3), ("compose_light_dir", C.c_float * 3),
                ("new_view", C.c_uint32), ("missing_strategy", C.c_uint32), ("rehash_count", C.c_uint32),
                ("hash_table_size", C.c_uint32), ("clip_enabled", C.c_uint32), ("clip_plane", C.c_float * 4),
                ("box_min", C.c_float * 3), ("box_max", C.c_float * 3), ("count_samples", C.c_uint32),
                ("model_pos_scale_m1", C.c_float * 3), ("model_pos_offset", C.c_float * 3)]


class Visibility(C.Structure):
    _fields_ = [("render_mode", C.c_uint32), ("smin", C.c_double), ("smax", C.c_double),
                ("gmin", C.c_double), ("gmax", C.c_double)]


class OrcPool(C.Structure):
    _fields_ = [("data", C.c_void_p), ("bpv", C.c_uint32), ("volume_size", C.c_uint32 * 3),
                ("max_brick", C.c_uint32 * 3), ("overlap", C.c_uint32), ("capacity", C.c_uint32 * 3),
                ("lod_count", C.c_uint32), ("meta", C.c_void_p)]


# ---------------------------------------------------------------- matrices (row-vector convention v' = v @ M)
def look_at(eye, at, up):
    eye, at, up = [np.asarray(v, f32) for v in (eye, at, up)]
    F = at - eye
    U = up.copy()
    S = np.cross(F, U).astype(f32)
    U = np.cross(S, F).astype(f32)
    F = F / f32(np.sqrt(f32(np.dot(F, F))))
    U = U / f32(np.sqrt(f32(np.dot(U, U))))
    S = S / f32(np.sqrt(f32(np.dot(S, S))))
    m = np.zeros((4, 4), f32)
    m[0, :3] = [S[0], U[0], -F[0]]
    m[1, :3] = [S[1], U[1], -F[1]]
    m[2, :3] = [S[2], U[2], -F[2]]
    m[3] = [-f32(np.dot(S, eye)), -f32(np.dot(U, eye)), f32(np.dot(F, eye)), 1]
    return m


def perspective(fovy_deg, aspect, n, f):
    fovy = f32(f32(fovy_deg) * f32(3.14159265358979323846 / 180.0))
    cotan = f32(1.0 / math.tan(float(fovy) / 2.0))
    aspect, n, f = f32(aspect), f32(n), f32(f)
    m = np.zeros((4, 4), f32)
    m[0, 0] = cotan / aspect
    m[1, 1] = cotan
    m[2, 2] = -(f + n) / (f - n)
    m[2, 3] = -1
    m[3, 2] = f32(-2) * (f * n) / (f - n)
    return m


def inverse(m):
    # the reference uses a cofactor expansion in fp32; a float64 inverse rounded to
    # fp32 is within a few ulp, and images are compared with a tolerance.  Both the
    # oracle and the CUDA path are always fed the SAME parameter struct.
    return np.linalg.inv(m.astype(np.float64)).astype(f32)


def translation(x, y, z):
    m = np.eye(4, dtype=f32)
    m[3, :3] = [x, y, z]
    return m


def scaling(x, y, z):
    return np.diag(np.asarray([x, y, z, 1], f32))


def rotation_y(angle):
    c, s = f32(math.cos(angle)), f32(math.sin(angle))
    m = np.eye(4, dtype=f32)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, -s, s, c
    return m


def rotation_x(angle):
    c, s = f32(math.cos(angle)), f32(math.sin(angle))
    m = np.eye(4, dtype=f32)
    m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, s, -s, c
    return m


def lod_factor(fovy_deg, pixels_y, sse=1.0):
    """CullingLOD::SetScreenParams, CullingLOD.cpp:65-66 (tan evaluated in double as ::tan)."""
    arg = f32(f32(fovy_deg) * f32(f32(f32(3.1416) / f32(180.0)) / f32(2.0)))
    return f32(f32(f32(2.0) * f32(math.tan(float(arg)))) * f32(sse) / f32(float(pixels_y)))


class Camera:
    """Defaults of AbstrRenderer.cpp:64-69,71-172."""

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
                   strategy=3, rehash=10, hash_size=509, count_samples=True, slab=None, clip_plane=None, view_proj=None):
    p = RaycastParams()
    p.width, p.height = width, height
    dom = np.asarray(domain_size, f32)
    sc = np.asarray(scale, f32)
    extent = dom * sc
    extent = extent / extent.max()                      # GetVolumeAABB, AbstrRenderer.cpp:1102-1109
    aspect = f32(width) / f32(height)
    view = look_at(cam.eye, cam.at, cam.up)
    proj = perspective(cam.fov, aspect, cam.znear, cam.zfar)
    if view_proj is not None:                            # one eye of a stereo pair (stereo_matrices)
        view, proj = view_proj
    mv = (cam.rotation @ cam.translation @ view).astype(f32)     # GLRenderer.cpp:627
    emm = (inverse(mv) @ translation(0, 0, 0) @ scaling(*(f32(1) / extent)) @ translation(0.5, 0.5, 0.5)).astype(f32)
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
    p.level_zero_wse = wse                             # GLVolumePool.cpp:801 (global volume)
    p.sample_rate_modifier = sample_rate
    p.trans_scale, p.gradient_scale, p.isoval = trans_scale, gradient_scale, isoval
    p.render_mode, p.lighting = render_mode, int(lighting)
    a = cam.ambient[:3] * cam.ambient[3]
    d = cam.diffuse[:3] * cam.diffuse[3]
    s = cam.specular[:3] * cam.specular[3]
    p.light_ambient[:] = a.tolist()
    p.light_diffuse[:] = d.tolist()
    p.light_specular[:] = s.tolist()
    ld = (np.append(cam.light_dir, f32(0)) @ emm_global)[:3]
    ld = ld / f32(np.sqrt(f32(np.dot(ld, ld))))
    p.model_light_dir[:] = ld.tolist()
    p.model_eye_pos[:] = (np.append(cam.eye, f32(1)) @ emm_global)[:3].tolist()   # sic: world eye through eye->model
    p.model_pos_scale_m1[:] = pos_scale_m1
    p.model_pos_offset[:] = pos_offset
    vs = sc / sc.min()
    p.domain_scale[:] = (f32(1) / vs).tolist()
    p.iso_color[:] = cam.iso_color.tolist()
    p.compose_light_dir[:] = cam.light_dir.tolist()
    p.new_view = int(new_view)
    p.missing_strategy, p.rehash_count, p.hash_table_size = strategy, rehash, hash_size
    p.clip_enabled = 0
    if clip_plane is not None:
        # GLGridLeaper::FillBBoxVBO (GLGridLeaper.cpp:515-526): world-space plane -> model space through
        # PLANE::operator*(inverse(rotation*translation)) (Vectors.h:1460-1477: plane4 * (M^-1)^T, then all four
        # components / |normal|); Clipper::BoxPlane keeps n.p + d <= 0.  Model space = centred box of size `extent`.
        m = inverse((cam.rotation @ cam.translation).astype(f32))
        q = (np.asarray(clip_plane, f32) @ inverse(m).T).astype(f32)
        q = (q / f32(np.sqrt(f32(np.dot(q[:3], q[:3]))))).astype(f32)
        n = q[:3] * extent
        p.clip_enabled = 1
        p.clip_plane[:] = [float(n[0]), float(n[1]), float(n[2]), float(q[3] - f32(0.5) * (n[0] + n[1] + n[2]))]
    p.box_min[:] = box_min
    p.box_max[:] = box_max
    p.count_samples = int(count_samples)
    return p


# ---------------------------------------------------------------- transfer functions
def tf1d_std(n, center=0.5, inv_gradient=0.5):
    """TransferFunction1D::SetStdFunction (TransferFunction1D.cpp:74-112) -> float table [n][4]."""
    t = np.zeros((n, 4), f32)
    ic = int((n - 1) * f32(min(max(0.0, center), 1.0)))
    ig = int((n - 1) * f32(min(max(0.0, inv_gradient), 1.0)))
    start = 0 if ig // 2 > ic else ic - ig // 2
    end = n if ig // 2 + ic > n else ic + ig // 2
    for i in range(start, end):
        x = f32(float(i - ic + ig // 2) / float(ig))
        t[i, :] = f32(3) * x * x - f32(2) * x * x * x
    t[end:, :] = 1
    return t


def tf_bytes(table):
    """TransferFunction1D::GetByteArray (TransferFunction1D.cpp:311-328): truncation, not rounding."""
    return (np.clip(np.asarray(table, f32), 0, 1) * f32(255)).astype(np.uint8)


def tf_nonzero_limits(table):
    """ComputeNonZeroLimits (TransferFunction1D.cpp:362-370)."""
    nz = np.nonzero(np.asarray(table)[:, 3] != 0)[0]
    return (len(table), 0) if nz.size == 0 else (int(nz[0]), int(nz[-1]))


# ---------------------------------------------------------------- pool
def largest_single_brick_lod(plan):
    for l in range(plan.lod_count):
        if plan.lod_bricks[l][0] * plan.lod_bricks[l][1] * plan.lod_bricks[l][2] == 1:
            return l
    return plan.lod_count - 1


def pool_layout(vol, max_brick, overlap, lod_count):
    lay = np.zeros((lod_count, 3), f32)
    off = np.zeros(lod_count, np.uint32)
    bricks = np.zeros((lod_count, 3), np.uint32)
    total = C.c_uint32()
    O.lib().orc_pool_layout((C.c_uint32 * 3)(*vol), (C.c_uint32 * 3)(*max_brick), C.c_uint32(overlap),
                            C.c_uint32(lod_count), lay.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                            bricks.ctypes.data_as(C.c_void_p), C.byref(total))
    return lay, off, bricks, total.value


class OraclePool:
    """GLVolumePool bookkeeping + a numpy pool texture, fed from an oracle octree (O.build_octree)."""

    def __init__(self, oct, capacity):
        self.oct = oct
        p = oct["plan"]
        self.vol = tuple(int(v) for v in oct["vol"])
        self.bpv = oct["bpv"]
        self.overlap = oct["overlap"]
        self.max_brick = tuple(int(v) for v in oct["sizes"].max(axis=0))   # GetMaxUsedBrickSizes
        self.lod_count = largest_single_brick_lod(p) + 1
        self.capacity = tuple(capacity)
        self.layout, self.lod_offset, self.lod_bricks, self.total = pool_layout(self.vol, self.max_brick, self.overlap, self.lod_count)
        dims = [self.capacity[a] * self.max_brick[a] for a in range(3)]
        self.data = np.zeros((dims[2], dims[1], dims[0]), np.uint8 if self.bpv == 1 else np.uint16)
        self.meta = np.zeros(self.total, np.uint32)
        nslots = capacity[0] * capacity[1] * capacity[2]
        self.slot_brick = -np.ones(nslots, np.int64)
        self.slot_time = np.zeros(nslots, np.uint64)
        self.slot_orig = np.zeros(nslots, np.uint64)
        self.time = 2
        self.insert_pos = 0
        self.order = np.arange(nslots)
        # min/max in POOL index order: m_vMinMaxScalar[i] = MaxMinForKey(IndexFrom4D(GetVectorBrickID(i)))
        # (GLVolumePool.cpp:232-239)
        self.minmax = np.zeros((self.total, 2), np.float64)
        for i in range(self.total):
            x, y, z, lod = self.index_to_xyzl(i)
            b = p.lod_bricks[lod]
            if x >= b[0] or y >= b[1] or z >= b[2]:
                raise ValueError("pool layout and octree layout disagree at LOD %d" % lod)
            self.minmax[i] = oct["minmax"][self.toc_index(x, y, z, lod)]
        self.vis = None

    # brick ids ----------------------------------------------------------------
    def brick_index(self, x, y, z, lod):
        b = self.lod_bricks[lod]
        return int(self.lod_offset[lod]) + x + y * int(b[0]) + z * int(b[0]) * int(b[1])

    def toc_index(self, x, y, z, lod):
        p = self.oct["plan"]
        b = p.lod_bricks[lod]
        return int(p.lod_offset[lod]) + x + y * int(b[0]) + z * int(b[0]) * int(b[1])

    def index_to_xyzl(self, idx):
        lod = int(np.searchsorted(self.lod_offset, idx, side="right") - 1)
        r = idx - int(self.lod_offset[lod])
        b = self.lod_bricks[lod]
        return r % int(b[0]), (r // int(b[0])) % int(b[1]), r // (int(b[0]) * int(b[1])), lod

    # uploads ------------------------------------------------------------------
    def _write_slot(self, slot, brick):
        cx, cy, _ = self.capacity
        s = (slot % cx, (slot // cx) % cy, slot // (cx * cy))
        o = [s[a] * self.max_brick[a] for a in range(3)]
        self.data[o[2]:o[2] + brick.shape[0], o[1]:o[1] + brick.shape[1], o[0]:o[0] + brick.shape[2]] = brick

    def _upload(self, idx, slot, time):
        """UploadBrick(iBrickID, ..., iInsertPos, iTimeOfCreation), GLVolumePool.cpp:673-717."""
        if self.slot_time[slot] > 1:                      # ContainsVisibleBrick
            self.meta[self.slot_brick[slot]] = 0
        self.slot_brick[slot] = idx
        self.slot_time[slot] = time
        self.meta[idx] = slot + 3
        x, y, z, lod = self.index_to_xyzl(idx)
        self._write_slot(slot, O.brick(self.oct, self.toc_index(x, y, z, lod)))

    def upload_first_brick(self):
        """UploadFirstBrick, :763-766: coarsest brick into the LAST slot, pinned forever."""
        self._upload(int(self.lod_offset[-1]), len(self.slot_brick) - 1, np.iinfo(np.uint64).max)

    def prepare_for_paging(self):
        self.order = np.argsort(self.slot_time, kind="stable")     # PoolSorter (stable tie-break = slot order)
        self.insert_pos = 0

    def upload_requests(self, requests):
        """UploadBricks -> UploadBricksToBrickPoolT (:1720-1789,1362-1393); visibility is up to date."""
        paged = 0
        if len(requests):
            self.prepare_for_paging()
            for x, y, z, lod in requests:
                if self.insert_pos >= len(self.slot_brick) - 1:
                    break
                slot = int(self.order[self.insert_pos])
                self._upload(self.brick_index(int(x), int(y), int(z), int(lod)), slot, self.time)
                self.time += 1
                self.insert_pos += 1
                paged += 1
        return paged

    # visibility ---------------------------------------------------------------
    def recompute_visibility(self, mode, smin, smax=0.0, gmin=0.0, gmax=0.0):
        """RecomputeVisibility (:1580-1718) incl. the slot FlagEmpty/Restore of ForBrickPool (:992-1018)."""
        v = Visibility(mode, smin, smax, gmin, gmax)
        self.vis = v
        for slot in range(len(self.slot_brick)):
            idx = self.slot_brick[slot]
            if idx < 0:
                continue
            lo, hi = self.minmax[idx]
            if mode == 2:
                cd = smin >= lo and smin <= hi
            else:
                cd = smax >= lo and smin <= hi
            contained = self.slot_time[slot] > 1
            if cd:
                if not contained:
                    self.slot_time[slot] = self.slot_orig[slot]      # Restore
                self.meta[idx] = slot + 3
            else:
                if contained:
                    self.slot_orig[slot] = self.slot_time[slot]      # FlagEmpty
                    self.slot_time[slot] = 1
                self.meta[idx] = 2
        # everything that is not resident-and-visible is recomputed by orc_visibility
        keep = self.meta.copy()
        keep[keep < 3] = 0
        counts = np.zeros(4, np.uint32)
        O.lib().orc_visibility((C.c_uint32 * 3)(*self.vol), (C.c_uint32 * 3)(*self.max_brick), C.c_uint32(self.overlap),
                               C.c_uint32(self.lod_count), C.byref(v), self.minmax.ctypes.data_as(C.c_void_p), None,
                               keep.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p))
        self.meta = keep
        return counts

    def c_pool(self):
        cp = OrcPool()
        cp.data = self.data.ctypes.data
        cp.bpv = self.bpv
        cp.volume_size[:] = self.vol
        cp.max_brick[:] = self.max_brick
        cp.overlap = self.overlap
        cp.capacity[:] = self.capacity
        cp.lod_count = self.lod_count
        cp.meta = self.meta.ctypes.data
        return cp


class OracleFrame:
    """Resume buffers + request table of one view."""

    def __init__(self, w, h, hash_size=509):
        self.pos = np.zeros((h, w, 4), f32)
        self.aux = np.zeros((h, w, 4), f32)
        self.image = np.zeros((h, w, 4), f32)
        self.hash = np.zeros(hash_size, np.uint32)
        self.stats = np.zeros(4, np.uint64)     # samples, GetBrick calls, rays, dropped reports


def raycast(params, pool, tf_rgba8, frame, rect=None, clear_hash=True):
    """One sub-frame.  rect = (x0, y0, x1, y1) restricts it to a pixel rectangle of the full-size buffers (crops of
    frames too large to render whole on the CPU); clear_hash=False accumulates the requests of several rectangles."""
    tf = np.ascontiguousarray(tf_rgba8, np.uint8)
    tf_w, tf_h = (tf.shape[0], 1) if tf.ndim == 2 else (tf.shape[1], tf.shape[0])
    cp = pool.c_pool()
    x0, y0, x1, y1 = (0, 0, params.width, params.height) if rect is None else rect
    rc = O.lib().orc_raycast_rect(C.byref(params), C.byref(cp), tf.ctypes.data_as(C.c_void_p), C.c_uint32(tf_w), C.c_uint32(tf_h),
                                  frame.pos.ctypes.data_as(C.c_void_p), frame.aux.ctypes.data_as(C.c_void_p),
                                  frame.image.ctypes.data_as(C.c_void_p), frame.hash.ctypes.data_as(C.c_void_p),
                                  frame.stats.ctypes.data_as(C.c_void_p), C.c_uint32(x0), C.c_uint32(y0), C.c_uint32(x1),
                                  C.c_uint32(y1), C.c_int(1 if clear_hash else 0))
    assert rc == 0
    return frame.image


def gen_sphere(vol, dtype, z0=0, nz=None):
    """The reference's CT_SPHERE test volume (UVFReader/DataSource.h:361-372) as a numpy array [z][y][x]."""
    vol = tuple(int(v) for v in vol)
    nz = vol[2] - z0 if nz is None else int(nz)
    out = np.empty((nz, vol[1], vol[0]), dtype)
    O.lib().orc_gen_sphere(out.ctypes.data_as(C.c_void_p), (C.c_uint64 * 3)(*vol), C.c_uint32(out.dtype.itemsize),
                           C.c_uint64(int(z0)), C.c_uint64(nz))
    return out


class ResidentPool:
    """A host pool with EVERY pool-LOD brick resident (slot i = brick i, the coarsest brick in the last slot), filled
    from a brick store in ToC order: the state bench.py / the full-size tests give the device pool.  Duck-types the
    part of OraclePool that raycast() needs."""

    def __init__(self, vol, bpv, overlap, max_brick, capacity, lod_count, store, offsets, sizes, slots, meta):
        self.vol, self.bpv, self.overlap = tuple(int(v) for v in vol), int(bpv), int(overlap)
        self.max_brick, self.capacity, self.lod_count = tuple(int(v) for v in max_brick), tuple(int(v) for v in capacity), int(lod_count)
        self.layout, self.lod_offset, self.lod_bricks, self.total = pool_layout(self.vol, self.max_brick, self.overlap, self.lod_count)
        dims = [self.capacity[a] * self.max_brick[a] for a in range(3)]
        self.data = np.zeros((dims[2], dims[1], dims[0]), np.uint8 if self.bpv == 1 else np.uint16)
        store = np.ascontiguousarray(store).view(np.uint8).reshape(-1)
        cap3, mb3 = (C.c_uint32 * 3)(*self.capacity), (C.c_uint32 * 3)(*self.max_brick)
        place = O.lib().orc_pool_place
        base = store.ctypes.data
        for i in range(len(slots)):
            bs = (C.c_uint32 * 3)(*[int(v) for v in sizes[i]])
            place(C.c_void_p(self.data.ctypes.data), cap3, mb3, C.c_uint32(self.bpv), C.c_uint32(int(slots[i])),
                  C.c_void_p(base + int(offsets[i])), bs)
        self.meta = np.ascontiguousarray(meta, np.uint32)

    def c_pool(self):
        return OraclePool.c_pool(self)


def read_requests(pool, frame, cap=4096):
    out = np.zeros((cap, 4), np.uint32)
    n = O.lib().orc_read_requests(frame.hash.ctypes.data_as(C.c_void_p), C.c_uint32(frame.hash.size),
                                  (C.c_uint32 * 3)(*[int(v) for v in pool.lod_bricks[0]]),
                                  out.ctypes.data_as(C.c_void_p), C.c_uint32(cap))
    return out[:n].copy()


def render_converged(params, pool, tf_rgba8, max_subframes=200):
    """paint() until converged (GLGridLeaper.cpp:872-890,914-1154).  Returns (image, subframes, paged)."""
    frame = OracleFrame(params.width, params.height, params.hash_table_size)
    params.new_view = 1
    paged = 0
    total_samples = 0
    for sub in range(max_subframes):
        raycast(params, pool, tf_rgba8, frame)
        total_samples += int(frame.stats[0])
        params.new_view = 0
        req = read_requests(pool, frame)
        if len(req) == 0:
            return frame.image.copy(), sub + 1, paged, total_samples
        paged += pool.upload_requests(req)
    raise RuntimeError("oracle render did not converge")


# ---------------------------------------------------------------- stereo (TEST INFRASTRUCTURE, restatement)
def perspective_off_center(left, right, bottom, top, n, f):
    """MATRIX4::MatrixPerspectiveOffCenter (Basics/Vectors.h:1286-1291), row-vector storage like perspective()"""
    left, right, bottom, top, n, f = [f32(v) for v in (left, right, bottom, top, n, f)]
    m = np.zeros((4, 4), f32)
    m[0, 0] = f32(2) * n / (right - left)
    m[2, 0] = (right + left) / (right - left)
    m[1, 1] = f32(2) * n / (top - bottom)
    m[2, 1] = (top + bottom) / (top - bottom)
    m[2, 2] = -(f + n) / (f - n)
    m[3, 2] = f32(-2) * (f * n) / (f - n)
    m[2, 3] = f32(-1)
    return m


def stereo_matrices(cam, aspect, focal_length=1.0, eye_dist=0.02):
    """FLOATMATRIX4::BuildStereoLookAtAndProjection (Basics/Vectors.h:1215-1248) -> ((view L, proj L), (view R, proj R))"""
    radians = f32(3.14159265358979323846 / 180.0) * f32(cam.fov) / f32(2)
    wd2 = f32(cam.znear) * f32(math.tan(float(radians)))
    shift = f32(eye_dist) * (f32(cam.znear) / f32(focal_length))
    a = f32(aspect)
    pl = perspective_off_center(-a * wd2 + shift, a * wd2 + shift, -wd2, wd2, cam.znear, cam.zfar)
    pr = perspective_off_center(-a * wd2 - shift, a * wd2 - shift, -wd2, wd2, cam.znear, cam.zfar)
    look = look_at(cam.eye, cam.at, cam.up)
    vl = (translation(f32(eye_dist), 0, 0) @ look).astype(f32)
    vr = (translation(-f32(eye_dist), 0, 0) @ look).astype(f32)
    return (vl, pl), (vr, pr)


def render_converged_stereo(p_left, p_right, pool, tf_rgba8, max_subframes=200):
    """GLGridLeaper::Render3DRegion with m_bDoStereoRendering (GLGridLeaper.cpp:922-958): per sub-frame the table is cleared
    once, both eyes march (each with its own resume buffers) and the table is read once.  -> (left, right, subframes)"""
    fl = OracleFrame(p_left.width, p_left.height, p_left.hash_table_size)
    fr = OracleFrame(p_right.width, p_right.height, p_right.hash_table_size)
    fr.hash = fl.hash
    p_left.new_view = p_right.new_view = 1
    for sub in range(max_subframes):
        raycast(p_left, pool, tf_rgba8, fl)
        raycast(p_right, pool, tf_rgba8, fr, clear_hash=False)
        p_left.new_view = p_right.new_view = 0
        req = read_requests(pool, fl)
        if len(req) == 0:
            return fl.image.copy(), fr.image.copy(), sub + 1
        pool.upload_requests(req)
    raise RuntimeError("oracle stereo render did not converge")


def stereo_compose(eye0, eye1, mode, swap=False, split=0.5, af_id=0):
    """GLRenderer::EndFrame's stereo pass (GLRenderer.cpp:770-812) with Compose-Anaglyphs-FS.glsl (mode 0),
    Compose-Scanline-FS.glsl (1), Compose-SBS-FS.glsl (2), Compose-AF-FS.glsl (3); images [h][w][4], GL_NEAREST"""
    left, right = (eye1, eye0) if swap else (eye0, eye1)
    h, w = left.shape[:2]
    if mode == 0:
        g = np.array([0.3, 0.59, 0.11], f32)
        gl = (left[..., :3].astype(np.float64) * g).sum(axis=2).astype(f32)
        gr = (right[..., :3].astype(np.float64) * g).sum(axis=2).astype(f32)
        return np.stack([gl, gr * f32(0.5), gr, np.maximum(left[..., 3], right[..., 3])], axis=2).astype(f32)
    if mode == 1:
        out = left.copy()
        out[1::2] = right[1::2]
        return out
    if mode == 2:
        out = np.empty_like(left)
        x = np.arange(w)
        tx = (x.astype(f32) + f32(0.5)) / f32(w)
        lhs = tx < f32(split)
        t = np.where(lhs, 2.0 * x + 1.0, 2.0 * x + 1.0 - 2.0 * float(f32(split)) * w)
        sx = np.clip(np.floor(t).astype(np.int64), 0, w - 1)
        out[:, lhs] = left[:, sx[lhs]]
        out[:, ~lhs] = right[:, sx[~lhs]]
        return out
    return (left if af_id == 0 else right).copy()


def psnr(a, b, peak=1.0):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)


# ---------------------------------------------------------------- synthetic volumes (UVFReader/DataSource.h)
def synth_sphere(size, dtype=np.uint8):
    """UVFReader/DataSource.h:361-372 'sphere': v = T(max(0,(0.5-|0.5-p/size|)*max*2))."""
    sx, sy, sz = size
    z, y, x = np.meshgrid(np.arange(sz), np.arange(sy), np.arange(sx), indexing="ij")
    fx, fy, fz = f32(0.5) - x.astype(f32) / f32(sx), f32(0.5) - y.astype(f32) / f32(sy), f32(0.5) - z.astype(f32) / f32(sz)
    d = np.sqrt(fx * fx + fy * fy + fz * fz).astype(f32)
    mx = f32(np.iinfo(dtype).max)
    return np.maximum(f32(0), (f32(0.5) - d) * mx * f32(2)).astype(dtype)


def synth_noise(size, dtype=np.uint8, seed=1, maxval=None):
    rng = np.random.default_rng(seed)
    hi = int(np.iinfo(dtype).max if maxval is None else maxval)
    return rng.integers(0, hi + 1, size=size[::-1]).astype(dtype)


def synth_blobs(size, dtype=np.uint8, seed=3):
    """smooth multi-blob field (sum of gaussians), a harder target than the sphere for LOD/visibility tests"""
    sx, sy, sz = size
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.linspace(0, 1, sz, dtype=f32), np.linspace(0, 1, sy, dtype=f32),
                          np.linspace(0, 1, sx, dtype=f32), indexing="ij")
    v = np.zeros((sz, sy, sx), f32)
    for _ in range(6):
        c = rng.random(3).astype(f32) * f32(0.6) + f32(0.2)
        r = f32(0.08 + 0.12 * rng.random())
        v += np.exp(-((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) / (r * r)).astype(f32)
    v = np.clip(v / v.max(), 0, 1)
    return (v * f32(np.iinfo(dtype).max)).astype(dtype)


# ---------------------------------------------------------------- classic per-brick ray-caster (R12 / R13)
class ClassicBrick(C.Structure):                     # tvk_classic_brick (include/tvk_cuda.h)
    _fields_ = [("center", C.c_float * 3), ("extension", C.c_float * 3), ("eye_to_tex", C.c_float * 16),
                ("ray_step", C.c_float), ("voxels", C.c_uint32 * 3), ("store_offset", C.c_uint64)]


class ClassicParams(C.Structure):                    # tvk_classic_params
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("inv_model_view", C.c_float * 16),
                ("normal_matrix", C.c_float * 9), ("inv_projection", C.c_float * 16), ("step_scale", C.c_float),
                ("trans_scale", C.c_float), ("lighting", C.c_uint32), ("light_ambient", C.c_float * 3),
                ("light_diffuse", C.c_float * 3), ("light_specular", C.c_float * 3), ("light_dir", C.c_float * 3),
                ("domain_scale", C.c_float * 3), ("bytes_per_voxel", C.c_uint32), ("render_mode", C.c_uint32),
                ("gradient_scale", C.c_float), ("isoval", C.c_float), ("proj_param", C.c_float * 2),
                ("iso_color", C.c_float * 3), ("clip_enabled", C.c_uint32), ("clip_plane", C.c_float * 4)]


def classic_plan(oc, cam, width, height, tf_limits, lighting=False, sample_rate=1.0, scale=(1, 1, 1), trans_scale=1.0,
                 capture=False, lod=None, render_mode=0, gradient_scale=1.0, iso_value=None, clip_plane=None):
    """What AbstrRenderer::PlanFrame / BuildSubFrameBrickList (AbstrRenderer.cpp:1125-1212, 999-1100) and
    GLRaycaster::SetBrickDepShaderVars / ComputeEyeToTextureMatrix hand to the shaders for the FINAL sub-frame of a
    view: (ClassicParams, [ClassicBrick...] front to back, lod).  oc: oracle octree (tight store offsets).
    render_mode 0/1/2 = 1D TF / 2D TF / isosurface (AbstrRenderer::ContainsData, :953-995): tf_limits = the scalar
    non-zero limits of the active table with the 1D table's size (a TOC file's gradient limits are -/+DBL_MAX, so the
    2D gradient test always passes); iso_value = the raw isovalue (m_fIsovalue)."""
    plan = oc["plan"]
    dom0 = np.array([int(v) for v in plan.lod_size[0]], f32)
    sc = np.asarray(scale, f32)
    aspect = f32(width) / f32(height)
    view = look_at(cam.eye, cam.at, cam.up)
    proj = perspective(cam.fov, aspect, cam.znear, cam.zfar)
    mv = (cam.rotation @ cam.translation @ view).astype(f32)
    # --- ComputeMinLODForCurrentView (:789-803) + CullingLOD::GetLODLevel (CullingLOD.cpp:128-141)
    extent = dom0 * sc
    extent = (extent / extent.max()).astype(f32)
    wse = f32((extent / dom0).max())
    zc = f32((np.array([0, 0, 0, 1], f32) @ mv)[2])
    fz = max(f32(cam.znear), f32(-zc))
    lf = lod_factor(cam.fov, height)
    level = int(math.floor(f32(f32(math.log(float(f32(lf * fz / wse)))) / f32(math.log(2.0)))))
    min_lod = min(max(level, 0), plan.lod_count - 1)
    if lod is None:
        lod = 0 if capture else min_lod
    # --- frustum planes (CullingLOD::Update, :89-126), model matrix = identity
    mvp = (mv @ proj).astype(f32)
    planes = []
    for col, sign in ((0, -1), (0, 1), (1, -1), (1, 1), (2, 1), (2, -1)):
        planes.append(np.array([sign * mvp[r, col] + mvp[r, 3] for r in range(4)], f32))
    # --- brick metadata of this LOD (UVFDataset::ComputeMetadataTOC, uvfDataset.cpp:242-288)
    nb = [int(v) for v in plan.lod_bricks[lod]]
    dl = np.array([int(v) for v in plan.lod_size[lod]], f32)
    asp = np.array([float(v) for v in plan.lod_aspect[lod]], f32)
    nds = (dl * asp).astype(f32)
    maxval = f32(nds.max())
    nds = (nds / maxval).astype(f32)
    overlap, brick = oc["overlap"], oc["brick"]
    first = int(plan.lod_offset[lod])
    vscale = sc.copy()
    corrected = (sc * dom0 / f32(dom0.max())).astype(f32)
    vscale = (vscale / f32(corrected.max())).astype(f32)
    tight = np.concatenate([[0], np.cumsum(oc["sizes"].astype(np.uint64).prod(axis=1) * oc["bpv"])]).astype(np.uint64)
    maxv = float(oc["minmax"][:int(np.prod([int(v) for v in plan.lod_bricks[0]])), 1].max())
    lo_i, hi_i, tf_n = tf_limits
    resc = maxv / float(tf_n - 1)
    out = []
    corner = np.zeros(3, f32)
    ext_axis = []
    for a in range(3):                                # extents along an axis depend on the brick coordinate only
        e = []
        for c in range(nb[a]):
            last = c >= nb[a] - 1
            rem = int(dl[a]) % (brick - 2 * overlap)
            size = (2 * overlap + rem) if (last and rem) else brick
            e.append(f32(f32(size - 2 * overlap) * asp[a] / maxval))
        ext_axis.append(e)
    inv_mv = inverse(mv)
    for x in range(nb[0]):
        for y in range(nb[1]):
            for z in range(nb[2]):
                idx = first + x + y * nb[0] + z * nb[0] * nb[1]
                ext = np.array([ext_axis[0][x], ext_axis[1][y], ext_axis[2][z]], f32)
                cor = np.array([sum(ext_axis[0][:x], f32(0)), sum(ext_axis[1][:y], f32(0)), sum(ext_axis[2][:z], f32(0))], f32)
                center = ((cor + ext / f32(2)) - nds * f32(0.5)).astype(f32)
                b_ext, b_cen = (ext * vscale).astype(f32), (center * vscale).astype(f32)
                # RegionNeedsBrick (:874-920): frustum, then data range
                half = f32(0.5) * b_ext
                vis = all(float(pl[0] * b_cen[0] + pl[1] * b_cen[1] + pl[2] * b_cen[2] + pl[3]) >
                          -float(half[0] * abs(pl[0]) + half[1] * abs(pl[1]) + half[2] * abs(pl[2])) for pl in planes)
                if not vis:
                    continue
                if clip_plane is not None:            # AbstrRenderer::Clipped (:923-948)
                    cpl = np.asarray(clip_plane, f32)
                    world = (cam.rotation @ cam.translation).astype(f32)
                    all_clipped = True
                    for k in range(8):
                        sgn = np.array([0.5 if k & 4 else -0.5, 0.5 if k & 2 else -0.5, 0.5 if k & 1 else -0.5], f32)
                        w4 = np.append((b_cen + sgn * b_ext).astype(f32), f32(1)) @ world
                        wp = (w4[:3] / w4[3]).astype(f32)
                        if not (float(cpl[0] * wp[0] + cpl[1] * wp[1] + cpl[2] * wp[2] + cpl[3]) >= 0):
                            all_clipped = False
                            break
                    if all_clipped:
                        continue
                mn, mx = oc["minmax"][idx]
                if render_mode == 2:
                    contains = float(iso_value) <= mx
                else:
                    contains = hi_i * resc >= mn and lo_i * resc <= mx
                if not contains:
                    continue                          # bIsEmpty: stays in the reference's list but is never drawn
                vox = oc["sizes"][idx].astype(np.uint32)
                tmin = (f32(overlap) / vox.astype(f32)).astype(f32)
                tmax = ((f32(1) - tmin) * asp).astype(f32)
                # brick_distance (:808-841)
                dist = None
                for sx in (-1, 1):
                    for sy in (-1, 1):
                        for sz in (-1, 1):
                            cpt = (b_cen + np.array([sx, sy, sz], f32) * b_ext * f32(0.4999)).astype(f32)
                            e4 = (np.append(cpt, f32(1)) @ mv)[:3]
                            dd = f32(np.sqrt(f32(np.dot(e4, e4))))
                            dist = dd if dist is None else min(dist, dd)
                p1, p2 = (b_cen + b_ext / f32(2)).astype(f32), (b_cen - b_ext / f32(2)).astype(f32)
                tm = (inv_mv @ translation(*(-p1)) @ scaling(*((tmin - tmax) / (p2 - p1))) @ translation(*tmax)).astype(f32)
                cb = ClassicBrick()
                cb.center[:] = b_cen.tolist(); cb.extension[:] = b_ext.tolist()
                cb.eye_to_tex[:] = tm.reshape(-1).tolist()
                cb.ray_step = float((b_ext * (f32(1) / vox.astype(f32)) * f32(0.5) * f32(1.0) / f32(sample_rate)).min())
                cb.voxels[:] = [int(v) for v in vox]
                cb.store_offset = int(tight[idx])
                out.append((float(dist), idx, cb))
    out.sort(key=lambda t: (t[0], t[1]))
    p = ClassicParams()
    p.width, p.height = width, height
    p.inv_model_view[:] = inv_mv.reshape(-1).tolist()
    p.normal_matrix[:] = inv_mv.T[:3, :3].reshape(-1).tolist()
    p.inv_projection[:] = inverse(proj).reshape(-1).tolist()
    p.step_scale = float(f32(1.0) / f32(sample_rate) * f32((dom0 / dl).max()))
    p.trans_scale = trans_scale
    p.lighting = int(lighting)
    p.light_ambient[:] = (cam.ambient[:3] * cam.ambient[3]).tolist()
    p.light_diffuse[:] = (cam.diffuse[:3] * cam.diffuse[3]).tolist()
    p.light_specular[:] = (cam.specular[:3] * cam.specular[3]).tolist()
    p.light_dir[:] = cam.light_dir.tolist()
    p.domain_scale[:] = (f32(1) / sc).tolist()
    p.bytes_per_voxel = oc["bpv"]
    p.render_mode = int(render_mode)
    p.gradient_scale = float(gradient_scale)
    if iso_value is not None:
        p.isoval = float(f32(float(iso_value) / float(1 << (8 * oc["bpv"]))))       # GetNormalizedIsovalue
    zn, zf = f32(cam.znear), f32(cam.zfar)
    p.proj_param[:] = [float(f32(zf / f32(zf - zn))), float(f32(f32(zf * zn) / f32(zn - zf)))]
    p.iso_color[:] = cam.iso_color.tolist()
    if clip_plane is not None:                        # GLRaycaster::RenderBox (:327-337)
        q = (np.asarray(clip_plane, f32) @ (cam.rotation @ cam.translation).astype(f32).T).astype(f32)
        ln = f32(np.sqrt(f32(q[0] * q[0] + q[1] * q[1] + q[2] * q[2])))
        if ln > 0:
            p.clip_enabled = 1
            p.clip_plane[:] = (q / ln).tolist()
    return p, [t[2] for t in out], lod


def classic_render(p, bricks, store_tight, tf_rgba8):
    """orc_raycast_classic: (image [H][W][4] float32, samples)"""
    from . import oracle as O
    arr = (ClassicBrick * max(len(bricks), 1))(*bricks)
    img = np.zeros((p.height, p.width, 4), f32)
    n = C.c_uint64()
    tf = np.ascontiguousarray(tf_rgba8, np.uint8)          # [n][4] (1D) or [h][w][4] (2D)
    tf_w, tf_h = (tf.shape[0], 1) if tf.ndim == 2 else (tf.shape[1], tf.shape[0])
    st = np.ascontiguousarray(store_tight, np.uint8)
    rc = O.lib().orc_raycast_classic(C.byref(p), arr, C.c_uint32(len(bricks)), st.ctypes.data_as(C.c_void_p),
                                     tf.ctypes.data_as(C.c_void_p), C.c_uint32(tf_w), C.c_uint32(tf_h),
                                     img.ctypes.data_as(C.c_void_p), C.byref(n))
    if rc:
        raise ValueError("orc_raycast_classic failed")
    return img, int(n.value)


class ClearViewParams(C.Structure):
    _fields_ = [("cv_isoval", C.c_float), ("cv_color", C.c_float * 3), ("size", C.c_float), ("context_scale", C.c_float),
                ("border_scale", C.c_float), ("pick_pos", C.c_float * 3)]


def classic_render_cv(p, cv, bricks, store_tight):
    """orc_raycast_classic_cv (isosurface mode with ClearView): (image, hit buffers [4][H][W][4], samples)"""
    from . import oracle as O
    arr = (ClassicBrick * max(len(bricks), 1))(*bricks)
    img = np.zeros((p.height, p.width, 4), f32)
    hits = np.zeros((4, p.height, p.width, 4), f32)
    n = C.c_uint64()
    st = np.ascontiguousarray(store_tight, np.uint8)
    rc = O.lib().orc_raycast_classic_cv(C.byref(p), C.byref(cv), arr, C.c_uint32(len(bricks)), st.ctypes.data_as(C.c_void_p),
                                        hits.ctypes.data_as(C.c_void_p), img.ctypes.data_as(C.c_void_p), C.byref(n))
    if rc:
        raise ValueError("orc_raycast_classic_cv failed")
    return img, hits, int(n.value)


# ---------------------------------------------------------------------------------------------------------------
# 2-D slice / MIP views: GLRenderer::Render2DView (GLRenderer.cpp:1057-1176), RenderSlice (:922-1014) and the shaders
# 1D-slice-FS.glsl, 2D-slice-FS.glsl, MIP-slice-FS.glsl, Transfer-MIP-FS.glsl.  TEST INFRASTRUCTURE (restatement; the GL
# rasteriser's interpolation of the quad is replaced by the texture coordinate at the pixel centre).  Pinned to the shader
# text by oracle/classicref.py (tests/test_glsl_classic_reference.py).
# ---------------------------------------------------------------------------------------------------------------
def slice_params(vol_dims, scale, lod_size, brick_voxels, bpv, width, height, window_mode, slice_index, mip, flip, render_mode,
                 max_value, max_grad):
    """Render2DView's set-up for one 2-D region covering a width x height target.  vol_dims: LOD-0 domain; lod_size /
    brick_voxels: domain and brick size (ghost voxels included) of the finest single-brick LOD."""
    f = np.float32
    mn = [(f(brick_voxels[a]) - f(lod_size[a]) + f(1.0)) / (f(2.0) * f(brick_voxels[a])) for a in range(3)]
    mx = [f(brick_voxels[a]) / f(brick_voxels[a]) - mn[a] for a in range(3)]
    aspect = [float(f(scale[a])) * float(vol_dims[a]) for a in range(3)]
    win = [1.0 / width, 1.0 / height]
    wm = max(win)
    win = [w / wm for w in win]
    ah = 1 if window_mode == 0 else 0
    av = 1 if window_mode == 2 else 2
    if flip[0]:
        mn[ah], mx[ah] = mx[ah], mn[ah]
    if flip[1]:
        mn[av], mx[av] = mx[av], mn[av]
    q = [aspect[ah] * win[0], aspect[av] * win[1]]
    qm = max(q)
    return dict(width=width, height=height, window_mode=window_mode, use_mip=int(bool(mip)), slice=f(float(slice_index) / float(vol_dims[window_mode])),
                voxels=[int(v) for v in brick_voxels], bpv=bpv, min_coords=mn, max_coords=mx, quad=[f(q[0] / qm), f(q[1] / qm)],
                render_mode=(1 if (render_mode == 1 and not mip) else 0), trans_scale=f(((1 << (8 * bpv)) - 1) / max_value),
                gradient_scale=f(1.0) if max_grad == 0 else f(1.0) / f(max_grad))


def _slice_sample(brick, c):
    """texture3D: GL_LINEAR, clamp to edge, UNORM; brick [z][y][x], c = (cx, cy, cz) arrays"""
    f = np.float32
    nz, ny, nx = brick.shape
    b = brick.astype(f)
    def taps(cc, n):
        u = cc.astype(f) * f(n) + f(-0.5)
        fl = np.floor(u).astype(f)
        return (u - fl).astype(f), np.clip(fl.astype(np.int64), 0, n - 1), np.clip(fl.astype(np.int64) + 1, 0, n - 1)
    ta, x0, x1 = taps(c[0], nx)
    tb, y0, y1 = taps(c[1], ny)
    tg, z0, z1 = taps(c[2], nz)
    lerp = lambda t, a0, a1: (t * (a1 - a0).astype(f) + a0).astype(f)   # one rounding for the product-sum is an fma; numpy rounds twice:
    def lerp_fma(t, a0, a1):                                            # evaluate in float64 and round once (exact for these magnitudes)
        return (t.astype(np.float64) * (a1 - a0).astype(f).astype(np.float64) + a0.astype(np.float64)).astype(f)
    c00, c10 = lerp_fma(ta, b[z0, y0, x0], b[z0, y0, x1]), lerp_fma(ta, b[z0, y1, x0], b[z0, y1, x1])
    c01, c11 = lerp_fma(ta, b[z1, y0, x0], b[z1, y0, x1]), lerp_fma(ta, b[z1, y1, x0], b[z1, y1, x1])
    c0, c1 = lerp_fma(tb, c00, c10), lerp_fma(tb, c01, c11)
    return (lerp_fma(tg, c0, c1) * (f(1.0) / f(255.0) if brick.dtype == np.uint8 else f(1.0) / f(65535.0))).astype(f)


def slice_texcoords(p):
    """the texture coordinate of every pixel centre on the quad GLRenderer::RenderSlice draws (GLRenderer.cpp:922-1014): which
    volume axes run along the screen axes in each window mode, flipView already folded into min / max coords.
    -> (inside [h][w] bool, [cx, cy, cz] float32 arrays); pinned to the reference's glTexCoord3d / glVertex3d calls by
    tests/test_slice_mip_host_pins.py"""
    f = np.float32
    w, h = p["width"], p["height"]
    px, py = np.meshgrid(np.arange(w), np.arange(h))
    nx = ((f(2.0) * (px.astype(f) + f(0.5))) / f(w) - f(1.0)).astype(f)
    ny = ((f(2.0) * (py.astype(f) + f(0.5))) / f(h) - f(1.0)).astype(f)
    qx, qy = f(p["quad"][0]), f(p["quad"][1])
    inside = (np.abs(nx) <= qx) & (np.abs(ny) <= qy)
    sx = np.where(inside, ((nx + qx) / (f(2.0) * qx)).astype(f), f(0))
    sy = np.where(inside, ((ny + qy) / (f(2.0) * qy)).astype(f), f(0))
    wm = p["window_mode"]
    ah, av, ad = (1 if wm == 0 else 0), (1 if wm == 2 else 2), wm
    mn, mx = [f(v) for v in p["min_coords"]], [f(v) for v in p["max_coords"]]
    fma = lambda t, d, a0: (t.astype(np.float64) * np.float64(d) + np.float64(a0)).astype(f)
    c = [None, None, None]
    c[ah] = fma(sx, f(mx[ah] - mn[ah]), mn[ah])
    c[av] = fma(sy, f(mx[av] - mn[av]), mn[av])
    c[ad] = np.full(sx.shape, f(p["slice"]), f)
    return inside, c


def slice_view(p, brick, tf_rgba8):
    """-> float32 image [h][w][4].  tf_rgba8: [n][4] (1-D) or [rows][n][4] (2-D, slice mode with render_mode 1)"""
    f = np.float32
    w, h = p["width"], p["height"]
    inside, c = slice_texcoords(p)
    sx = c[0]
    wm = p["window_mode"]
    ad = wm
    tf = np.asarray(tf_rgba8)
    tfw = tf.shape[-2]
    nearest = lambda s, n: np.clip((s * f(n)).astype(f).astype(np.int64), 0, n - 1)
    img = np.zeros((h, w, 4), f)
    img[..., 3] = 1.0
    if not p["use_mip"]:
        val = _slice_sample(brick, c)
        if p["render_mode"] == 1:
            hx, hy, hz = [f(1.0) / f(v) for v in p["voxels"]]
            s = lambda dx, dy, dz: _slice_sample(brick, [(c[0] + f(dx)).astype(f), (c[1] + f(dy)).astype(f), (c[2] + f(dz)).astype(f)])
            gx = ((s(-hx, 0, 0) - s(hx, 0, 0)) / f(2.0)).astype(f)
            gy = ((s(0, -hy, 0) - s(0, hy, 0)) / f(2.0)).astype(f)
            gz = ((s(0, 0, -hz) - s(0, 0, hz)) / f(2.0)).astype(f)
            gm = np.sqrt((gz.astype(np.float64) * gz + (gy.astype(np.float64) * gy + (gx * gx).astype(f)).astype(f)).astype(f)).astype(f)
            t = tf[nearest((f(1.0) - gm * f(p["gradient_scale"])).astype(f), tf.shape[0]), nearest(val * f(p["trans_scale"]), tfw)]
        else:
            t = tf.reshape(-1, tfw, 4)[0][nearest(val * f(p["trans_scale"]), tfw)]
        t = t.astype(f) * (f(1.0) / f(255.0))
        rgb = t[..., :3] * t[..., 3:4]
        img[..., :3] = np.where(inside[..., None], rgb, f(0))
    else:
        m = np.zeros(sx.shape, f)
        ns = p["voxels"][ad]
        for i in range(ns):
            c[ad] = np.full(sx.shape, f(np.float64(i) / np.float64(ns)), f)
            m = np.maximum(m, _slice_sample(brick, c))
        m = np.where(inside, m, f(0))
        t = tf.reshape(-1, tfw, 4)[0][nearest(m * f(p["trans_scale"]), tfw)].astype(f) * (f(1.0) / f(255.0))
        img[..., :3] = t[..., :3]
    return img


# ---------------------------------------------------------------------------------------------------------------
# Full-resolution MIP of the capture mode: GLRaycaster::RenderHQMIPInLoop (GLRaycaster.cpp:494-529), GLRaycaster-MIP-Rot-FS.glsl,
# Transfer-MIP-FS.glsl, on the plan the host made (model-view, projection, brick list).  TEST INFRASTRUCTURE (numpy
# restatement; the rasterised front / back faces of a brick are the analytic entry / exit of the pixel-centre ray).
# ---------------------------------------------------------------------------------------------------------------
def _fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def _xform(m, x, y, z, w):
    """(x, y, z, w) * M in the kernel's order of operations; m: 16 floats, FLOATMATRIX4 memory order"""
    f = np.float32
    m = [f(v) for v in m]
    return [_fma(w, m[12 + j], _fma(z, m[8 + j], _fma(y, m[4 + j], (np.asarray(x, f) * m[j]).astype(f)))) for j in range(4)]


def hq_mip(p, bricks, store, tf_rgba8):
    """p: the plan's parameters (width, height, inv_model_view, inv_projection, bytes_per_voxel, trans_scale); bricks: objects
    with center / extension / eye_to_tex / ray_step / voxels / store_offset; store: the brick store as bytes -> [h][w][4]"""
    f = np.float32
    w, h = int(p.width), int(p.height)
    px, py = np.meshgrid(np.arange(w), np.arange(h))
    nx = _fma(((px.astype(f) + f(0.5)) / f(w)).astype(f), f(2.0), f(-1.0))
    ny = _fma(((py.astype(f) + f(0.5)) / f(h)).astype(f), f(2.0), f(-1.0))
    one, mone = np.ones_like(nx), -np.ones_like(nx)
    n4, f4 = _xform(p.inv_projection[:], nx, ny, mone, one), _xform(p.inv_projection[:], nx, ny, one, one)
    e0 = [(n4[k] / n4[3]).astype(f) for k in range(3)]
    ed = [((f4[k] / f4[3]).astype(f) - e0[k]).astype(f) for k in range(3)]
    O = _xform(p.inv_model_view[:], e0[0], e0[1], e0[2], one)
    D = _xform(p.inv_model_view[:], ed[0], ed[1], ed[2], np.zeros_like(nx))
    vmax = np.zeros((h, w), f)
    covered = np.zeros((h, w), bool)
    bpv = int(p.bytes_per_voxel)
    dt = np.uint8 if bpv == 1 else np.uint16
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for b in bricks:
            s0 = np.full((h, w), -np.inf, f)
            s1 = np.full((h, w), np.inf, f)
            miss = np.zeros((h, w), bool)
            for k in range(3):
                lo, hi = f(b.center[k]) - f(b.extension[k]) / f(2.0), f(b.center[k]) + f(b.extension[k]) / f(2.0)
                zero = D[k] == 0
                miss |= zero & ((O[k] < lo) | (O[k] > hi))
                inv = (f(1.0) / D[k]).astype(f)
                ta, tb = ((lo - O[k]) * inv).astype(f), ((hi - O[k]) * inv).astype(f)
                ta, tb = np.minimum(ta, tb), np.maximum(ta, tb)
                s0 = np.where(zero, s0, np.maximum(s0, ta))
                s1 = np.where(zero, s1, np.minimum(s1, tb))
            se = np.maximum(s0, f(0.0))
            s1 = np.minimum(s1, f(1.0))
            act = ~miss & (s1 > se)
            if not act.any():
                continue
            covered |= act
            en = [_fma(ed[k], se, e0[k]).astype(np.float16).astype(f) for k in range(3)]
            ex = [_fma(ed[k], s1, e0[k]) for k in range(3)]
            te = _xform(b.eye_to_tex[:], en[0], en[1], en[2], one)
            tx = _xform(b.eye_to_tex[:], ex[0], ex[1], ex[2], one)
            dd = [(ex[k] - en[k]).astype(f) for k in range(3)]
            ln = np.sqrt(_fma(dd[2], dd[2], _fma(dd[1], dd[1], (dd[0] * dd[0]).astype(f)))).astype(f)
            q = (ln / f(b.ray_step)).astype(f)
            steps = np.where(act, q.astype(np.int64) + 1, 0)
            it = [((tx[k] - te[k]) / q).astype(f) for k in range(3)]
            vox = [int(v) for v in b.voxels]
            n = vox[0] * vox[1] * vox[2]
            data = np.frombuffer(store, np.uint8, n * bpv, int(b.store_offset)).view(dt).reshape(vox[2], vox[1], vox[0])
            pt = [te[k].copy() for k in range(3)]
            for i in range(int(steps.max())):
                m = steps > i
                if not m.any():
                    break
                val = _slice_sample(data, [pt[0][m], pt[1][m], pt[2][m]])
                vmax[m] = np.maximum(vmax[m], val)
                for k in range(3):
                    pt[k] = (pt[k] + it[k]).astype(f)
    tf = np.asarray(tf_rgba8).reshape(-1, 4)
    idx = np.clip(((vmax * f(p.trans_scale)).astype(f) * f(len(tf))).astype(f).astype(np.int64), 0, len(tf) - 1)
    img = np.zeros((h, w, 4), f)
    img[..., 3] = 1.0
    rgb = tf[idx][..., :3].astype(f) * (f(1.0) / f(255.0))
    img[..., :3] = np.where(covered[..., None], rgb, f(0))
    return img, vmax


def hq_mip_lod(domain, win, lod_count, mip_lod=True):
    """AbstrRenderer::PlanHQMIPFrame (AbstrRenderer.cpp:1220-1235): the domain is halved while its smallest side is not below the
    largest window side, then one level back (pinned to the reference's statements: tests/test_slice_mip_host_pins.py)"""
    vc = [int(v) for v in domain]
    lod = 0
    if mip_lod:
        while min(vc) >= max(win):
            vc = [v // 2 for v in vc]
            lod += 1
    if lod > 0:
        lod = min(int(lod_count) - 1, lod - 1)
    return lod


def hq_mip_matrices(window_mode, flip, angle_deg, ortho, width, height, cam):
    """GLRenderer::RenderHQMIPPreLoop (:1256-1283) + the orthographic projection of Render2DView (:1183-1197) /
    GLRaycaster::RenderHQMIPPreLoop (GLRaycaster.cpp:481-492) -> (model-view, projection), row-vector convention"""
    f = np.float32
    rot_dir = np.eye(4, dtype=f)
    if window_mode == 0:
        rot_dir = (rotation_x(-math.pi / 2.0) @ rotation_y(-math.pi / 2.0)).astype(f)
    elif window_mode == 1:
        rot_dir = rotation_x(-math.pi / 2.0)
    flip_y = scaling(-1, 1, 1) if flip[0] else np.eye(4, dtype=f)
    flip_x = scaling(1, -1, 1) if flip[1] else np.eye(4, dtype=f)
    mip_rot = (rot_dir @ np.eye(4, dtype=f) @ flip_x @ flip_y @ rotation_y(math.pi * float(f(angle_deg)) / 180.0)).astype(f)
    if ortho:
        wa = np.array([1.0 / width, 1.0 / height])
        wa = wa / wa.max()
        r2 = max(f(1.0), f(1.414213) * f(wa[0] / wa[1])) if wa[0] < wa[1] else f(1.414213)
        l, r = f(-0.5) * r2 / f(wa[0]), f(0.5) * r2 / f(wa[0])
        bt, tp = f(-0.5) * r2 / f(wa[1]), f(0.5) * r2 / f(wa[1])
        zn, zf = f(-100.0), f(100.0)
        proj = np.eye(4, dtype=f)
        proj[0, 0] = f(2) / (r - l); proj[3, 0] = -(r + l) / (r - l)       # noqa: E702
        proj[1, 1] = f(2) / (tp - bt); proj[3, 1] = -(tp + bt) / (tp - bt)  # noqa: E702
        proj[2, 2] = -f(2) / (zf - zn); proj[3, 2] = -(zf + zn) / (zf - zn)  # noqa: E702
        return mip_rot, proj
    return (mip_rot @ look_at(cam.eye, cam.at, cam.up)).astype(f), perspective(cam.fov, f(width) / f(height), cam.znear, cam.zfar)
