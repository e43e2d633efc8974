"""ctypes mirror of include/tvk_uvf.h: UVF v5 container writer / reader (host code in libtvk_cuda.so).

write_uvf() takes the arrays the K2 builder hands out (TvkContext.build_toc / build_bricks / build_minmax /
build_hist1d / build_hist2d); UvfFile reads a UVF's TOC block, histograms, max/min and key/value blocks."""
import ctypes as C

import numpy as np

from . import _lib as L

COMP_NONE, COMP_ZLIB, COMP_LZMA, COMP_LZ4, COMP_BZIP = 0, 1, 2, 3, 4


class UvfVolume(C.Structure):
    _fields_ = [("component_type", C.c_uint32), ("component_count", C.c_uint64), ("volume_size", C.c_uint64 * 3),
                ("aspect", C.c_double * 3), ("brick_size", C.c_uint64 * 3), ("overlap", C.c_uint32),
                ("compression", C.c_uint32), ("compression_level", C.c_uint32)]


class UvfWriteDesc(C.Structure):
    _fields_ = [("vol", UvfVolume), ("toc_block_id", C.c_char_p), ("total_bricks", C.c_uint64),
                ("brick_offsets", C.c_void_p), ("brick_sizes", C.c_void_p), ("store", C.c_void_p), ("minmax", C.c_void_p),
                ("hist1d", C.c_void_p), ("hist1d_size", C.c_uint64),
                ("hist2d", C.c_void_p), ("hist2d_x", C.c_uint64), ("hist2d_y", C.c_uint64),
                ("max_grad_magnitude", C.c_float),
                ("kv_keys", C.POINTER(C.c_char_p)), ("kv_values", C.POINTER(C.c_char_p)), ("kv_count", C.c_uint32)]


class UvfInfo(C.Structure):
    _fields_ = [("uvf_version", C.c_uint64), ("block_count", C.c_uint32), ("block_semantics", C.c_uint32 * 16),
                ("toc_block", C.c_int32), ("vol", UvfVolume), ("octree_version", C.c_uint32), ("lod_count", C.c_uint32),
                ("total_bricks", C.c_uint64), ("store_bytes", C.c_uint64), ("hist1d_size", C.c_uint64),
                ("hist2d_x", C.c_uint64), ("hist2d_y", C.c_uint64), ("max_grad_magnitude", C.c_float),
                ("maxmin_bricks", C.c_uint64), ("kv_count", C.c_uint32), ("checksum_ok", C.c_uint32),
                ("rdb_block", C.c_int32), ("rdb_bits", C.c_uint32), ("rdb_signed", C.c_uint32),
                ("rdb_domain", C.c_uint64 * 3), ("rdb_brick", C.c_uint64 * 3), ("rdb_overlap", C.c_uint64 * 3),
                ("rdb_lod_levels", C.c_uint64), ("timesteps", C.c_uint32)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def md5(data):
    buf = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    out = (C.c_uint8 * 16)()
    L.load().tvk_md5(_p(buf) if buf.size else None, C.c_uint64(buf.size), out)
    return bytes(out)


def write_uvf(path, volume_size, bpv, brick, overlap, offsets, sizes, store, minmax, hist1d=None, hist2d=None,
              max_grad=0.0, kv=(("Source Endianness", "little"), ("Source Type", "integer")), compression=COMP_NONE,
              level=1, aspect=(1.0, 1.0, 1.0), toc_block_id=None):
    lib = L.load()
    offsets = np.ascontiguousarray(offsets, np.uint64)
    sizes = np.ascontiguousarray(sizes, np.uint32)
    store = np.ascontiguousarray(store).view(np.uint8).reshape(-1)
    minmax = np.ascontiguousarray(minmax, np.float64)
    d = UvfWriteDesc()
    d.vol.component_type = 1 if bpv == 2 else 0
    d.vol.component_count = 1
    d.vol.volume_size[:] = [int(v) for v in volume_size]
    d.vol.aspect[:] = [float(v) for v in aspect]
    d.vol.brick_size[:] = [int(brick)] * 3
    d.vol.overlap = int(overlap)
    d.vol.compression = int(compression)
    d.vol.compression_level = int(level)
    d.toc_block_id = toc_block_id.encode() if toc_block_id else None
    d.total_bricks = offsets.size
    d.brick_offsets, d.brick_sizes, d.store, d.minmax = _p(offsets).value, _p(sizes).value, _p(store).value, _p(minmax).value
    keep = [offsets, sizes, store, minmax]
    if hist1d is not None:
        h1 = np.ascontiguousarray(hist1d, np.uint64)
        d.hist1d, d.hist1d_size = _p(h1).value, h1.size
        keep.append(h1)
    if hist2d is not None:
        h2 = np.ascontiguousarray(hist2d, np.uint64)
        d.hist2d, d.hist2d_x, d.hist2d_y = _p(h2).value, h2.shape[0], h2.shape[1]
        keep.append(h2)
    d.max_grad_magnitude = float(max_grad)
    kv = list(kv) + ([("Source Bitwidth", str(8 * bpv))] if not any(k == "Source Bitwidth" for k, _ in kv) else [])
    keys = (C.c_char_p * len(kv))(*[k.encode() for k, _ in kv])
    vals = (C.c_char_p * len(kv))(*[v.encode() for _, v in kv])
    d.kv_keys, d.kv_values, d.kv_count = keys, vals, len(kv)
    err = C.create_string_buffer(256)
    rc = lib.tvk_uvf_write(path.encode(), C.byref(d), err, C.c_size_t(256))
    if rc:
        raise L.TvkError("tvk_uvf_write: " + err.value.decode())


class UvfFile:
    def __init__(self, path, verify=True, timestep=0):
        self.lib = L.load()
        self.h = C.c_void_p()
        err = C.create_string_buffer(256)
        rc = self.lib.tvk_uvf_open_timestep(path.encode(), C.c_int(1 if verify else 0), C.c_uint32(int(timestep)), C.byref(self.h), err,
                                            C.c_size_t(256))
        if rc:
            raise L.TvkError("tvk_uvf_open(%s): %s" % (path, err.value.decode()))
        self.info = UvfInfo()
        self.lib.tvk_uvf_get_info(self.h, C.byref(self.info))

    def close(self):
        if self.h:
            self.lib.tvk_uvf_close(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc, what):
        if rc:
            self.lib.tvk_uvf_last_error.restype = C.c_char_p
            raise L.TvkError("%s failed (%d): %s" % (what, rc, (self.lib.tvk_uvf_last_error(self.h) or b"").decode()))

    def toc(self):
        n = int(self.info.total_bricks)
        off, ln = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        comp, sizes = np.zeros(n, np.uint32), np.zeros((n, 3), np.uint32)
        self._ck(self.lib.tvk_uvf_get_toc(self.h, _p(off), _p(ln), _p(comp), _p(sizes), C.c_uint64(n)), "tvk_uvf_get_toc")
        return off, ln, comp, sizes

    def read_bricks(self, first=0, count=None):
        """tightly packed, decompressed brick payloads of bricks [first, first+count)"""
        count = int(self.info.total_bricks) - first if count is None else count
        _, _, _, sizes = self.toc()
        bpv = 2 if self.info.vol.component_type == 1 else 1
        nbytes = int(sizes[first:first + count].astype(np.uint64).prod(axis=1).sum()) * bpv
        out = np.zeros(nbytes, np.uint8)
        self._ck(self.lib.tvk_uvf_read_bricks(self.h, C.c_uint64(first), C.c_uint64(count), _p(out), None), "tvk_uvf_read_bricks")
        return out

    def read_bricks_raw(self, first=0, count=None):
        """the same bricks as the file stores them (still compressed), packed; lengths = toc()[1]"""
        count = int(self.info.total_bricks) - first if count is None else count
        _, ln, _, _ = self.toc()
        out = np.zeros(int(ln[first:first + count].sum()), np.uint8)
        self._ck(self.lib.tvk_uvf_read_bricks_raw(self.h, C.c_uint64(first), C.c_uint64(count), _p(out), None), "tvk_uvf_read_bricks_raw")
        return out

    def read_rdb_volume(self):
        """finest level of a legacy RasterDataBlock file as a [z][y][x] array"""
        i = self.info
        dt = np.uint8 if i.rdb_bits == 8 else np.uint16
        out = np.zeros((int(i.rdb_domain[2]), int(i.rdb_domain[1]), int(i.rdb_domain[0])), dt)
        self._ck(self.lib.tvk_uvf_read_rdb_volume(self.h, _p(out), C.c_uint64(out.nbytes)), "tvk_uvf_read_rdb_volume")
        return out

    def minmax(self):
        n = int(self.info.maxmin_bricks)
        mm = np.zeros((n, 4), np.float64)
        self._ck(self.lib.tvk_uvf_get_minmax(self.h, _p(mm), C.c_uint64(n)), "tvk_uvf_get_minmax")
        return mm

    def hist1d(self):
        h = np.zeros(int(self.info.hist1d_size), np.uint64)
        if h.size:
            self._ck(self.lib.tvk_uvf_get_hist1d(self.h, _p(h), C.c_uint64(h.size)), "tvk_uvf_get_hist1d")
        return h

    def hist2d(self):
        h = np.zeros((int(self.info.hist2d_x), int(self.info.hist2d_y)), np.uint64)
        if h.size:
            self._ck(self.lib.tvk_uvf_get_hist2d(self.h, _p(h), C.c_uint64(h.size)), "tvk_uvf_get_hist2d")
        return h

    def kv(self):
        out = []
        for i in range(int(self.info.kv_count)):
            k, v = C.c_char_p(), C.c_char_p()
            self._ck(self.lib.tvk_uvf_get_kv(self.h, C.c_uint32(i), C.byref(k), C.byref(v)), "tvk_uvf_get_kv")
            out.append((k.value.decode(), v.value.decode()))
        return out
