"""Thin object wrapper over the kernel-level C ABI (include/tvk_cuda.h).

Only marshals numpy arrays into the C calls; all arithmetic happens in
libtvk_cuda.so.  The AbstrRenderer-shaped host lives in csrc/host/ (C++) and is
exposed through imagevis3d_b200.renderer.
"""
import ctypes as C

import numpy as np

from . import _lib as L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class TvkContext:
    def __init__(self, device=0):
        self.lib = L.load()
        self.h = C.c_void_p()
        rc = self.lib.tvk_create(C.c_int(device), C.byref(self.h))
        if rc != L.TVK_OK:
            raise L.TvkError("tvk_create(device=%d) failed with status %d (no CUDA device?)" % (device, rc))
        self.device = device
        self.total_bricks = 0
        self.build_info = None

    def close(self):
        if self.h:
            self.lib.tvk_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        L.check(self.h, rc, what)

    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); None = own stream"""
        self._ck(self.lib.tvk_set_stream(self.h, C.c_void_p(cuda_stream)), "tvk_set_stream")

    def synchronize(self):
        self._ck(self.lib.tvk_device_synchronize(self.h), "tvk_device_synchronize")

    # ------------------------------------------------------------------ pool
    def set_march_occupancy(self, ctas_per_sm):
        """resident ray-march CTAs per SM (0 = full occupancy); fewer leaves room for concurrent paging"""
        self._ck(self.lib.tvk_set_march_occupancy(self.h, C.c_uint32(int(ctas_per_sm))), "tvk_set_march_occupancy")

    def pool_create(self, volume_size, max_brick_size, overlap, bytes_per_voxel, capacity, lod_count,
                    sampler=L.SAMPLER_EXACT):
        d = L.PoolDesc()
        d.volume_size[:] = [int(v) for v in volume_size]
        d.max_brick_size[:] = [int(v) for v in max_brick_size]
        d.overlap = overlap
        d.bytes_per_voxel = bytes_per_voxel
        d.pool_capacity[:] = [int(v) for v in capacity]
        d.lod_count = lod_count
        d.sampler = sampler
        self._ck(self.lib.tvk_pool_create(self.h, C.byref(d)), "tvk_pool_create")
        n = C.c_uint32()
        self._ck(self.lib.tvk_pool_total_bricks(self.h, C.byref(n)), "tvk_pool_total_bricks")
        self.total_bricks = n.value
        self.pool_desc = d
        return n.value

    def pool_upload_host(self, slots, bricks):
        """bricks: list of C-contiguous numpy arrays [z][y][x] (host memory)."""
        n = len(slots)
        slot = np.asarray(slots, np.uint32)
        src = (C.c_void_p * n)(*[b.ctypes.data for b in bricks])
        sizes = np.asarray([[b.shape[2], b.shape[1], b.shape[0]] for b in bricks], np.uint32).reshape(-1)
        self._ck(self.lib.tvk_pool_upload_bricks(self.h, C.c_uint32(n), _p(slot), src, _p(sizes), C.c_int(0)),
                 "tvk_pool_upload_bricks")
        self._ck(self.lib.tvk_pool_upload_wait(self.h), "tvk_pool_upload_wait")

    def pool_upload_device(self, slots, dptrs, sizes3):
        n = len(slots)
        slot = np.asarray(slots, np.uint32)
        src = (C.c_void_p * n)(*[int(p) for p in dptrs])
        sizes = np.ascontiguousarray(sizes3, np.uint32).reshape(-1)
        self._ck(self.lib.tvk_pool_upload_bricks(self.h, C.c_uint32(n), _p(slot), src, _p(sizes), C.c_int(1)),
                 "tvk_pool_upload_bricks")
        self._ck(self.lib.tvk_pool_upload_wait(self.h), "tvk_pool_upload_wait")

    def set_metadata(self, meta):
        meta = np.ascontiguousarray(meta, np.uint32)
        self._ck(self.lib.tvk_pool_set_metadata(self.h, _p(meta), C.c_uint32(meta.size)), "tvk_pool_set_metadata")

    def patch_metadata(self, idx, val):
        idx = np.ascontiguousarray(idx, np.uint32)
        val = np.ascontiguousarray(val, np.uint32)
        self._ck(self.lib.tvk_pool_patch_metadata(self.h, C.c_uint32(idx.size), _p(idx), _p(val)),
                 "tvk_pool_patch_metadata")

    def get_metadata(self):
        meta = np.zeros(self.total_bricks, np.uint32)
        self._ck(self.lib.tvk_pool_get_metadata(self.h, _p(meta), C.c_uint32(meta.size)), "tvk_pool_get_metadata")
        return meta

    # ------------------------------------------------------------ visibility
    def set_minmax(self, minmax, grad_minmax=None):
        mm = np.ascontiguousarray(minmax, np.float64)
        g = None if grad_minmax is None else np.ascontiguousarray(grad_minmax, np.float64)
        self._ck(self.lib.tvk_visibility_set_minmax(self.h, _p(mm), _p(g) if g is not None else None,
                                                    C.c_uint32(mm.shape[0])), "tvk_visibility_set_minmax")

    def recompute_visibility(self, mode, smin, smax=0.0, gmin=0.0, gmax=0.0):
        v = L.Visibility(mode, smin, smax, gmin, gmax)
        counts = np.zeros(4, np.uint32)
        self._ck(self.lib.tvk_visibility_recompute(self.h, C.byref(v), _p(counts)), "tvk_visibility_recompute")
        return counts

    # ------------------------------------------------------------------- TFs
    def set_tf1d(self, rgba):
        rgba = np.ascontiguousarray(rgba, np.uint8)
        self._ck(self.lib.tvk_set_tf1d(self.h, _p(rgba), C.c_uint32(rgba.shape[0])), "tvk_set_tf1d")

    def set_tf2d(self, rgba):
        rgba = np.ascontiguousarray(rgba, np.uint8)   # [h][w][4]
        self._ck(self.lib.tvk_set_tf2d(self.h, _p(rgba), C.c_uint32(rgba.shape[1]), C.c_uint32(rgba.shape[0])),
                 "tvk_set_tf2d")

    # --------------------------------------------------------------- raycast
    def raycast(self, params):
        self._ck(self.lib.tvk_raycast_subframe(self.h, C.byref(params)), "tvk_raycast_subframe")
        self._wh = (params.width, params.height)

    def read_requests(self, cap=4096):
        out = np.zeros((cap, 4), np.uint32)
        n = C.c_uint32()
        self._ck(self.lib.tvk_read_requests(self.h, _p(out), C.c_uint32(cap), C.byref(n)), "tvk_read_requests")
        return out[:n.value].copy()

    def read_image(self):
        w, h = self._wh
        img = np.zeros((h, w, 4), np.float32)
        self._ck(self.lib.tvk_read_image(self.h, _p(img)), "tvk_read_image")
        return img

    def read_image_rgba8(self, flip_rows=False, out=None):
        """glReadPixels(GL_RGBA, GL_UNSIGNED_BYTE): converted on the device, [h][w][4] uint8."""
        w, h = self._wh
        img = np.empty((h, w, 4), np.uint8) if out is None else out
        self._ck(self.lib.tvk_read_image_rgba8(self.h, _p(img), C.c_int(int(flip_rows))), "tvk_read_image_rgba8")
        return img

    def read_resume(self):
        w, h = self._wh
        pos = np.zeros((h, w, 4), np.float32)
        col = np.zeros((h, w, 4), np.float32)
        self._ck(self.lib.tvk_read_resume(self.h, _p(pos), _p(col)), "tvk_read_resume")
        return pos, col

    def stats(self, reset=False):
        s = L.Stats()
        self._ck(self.lib.tvk_get_stats(self.h, C.byref(s), C.c_int(1 if reset else 0)), "tvk_get_stats")
        return s

    def probe_lod(self, depth, lod_factor, wse):
        depth = np.ascontiguousarray(depth, np.float32)
        out = np.zeros(depth.size, np.uint32)
        self._ck(self.lib.tvk_probe_lod(self.h, _p(depth), C.c_uint32(depth.size), C.c_float(lod_factor),
                                        C.c_float(wse), _p(out)), "tvk_probe_lod")
        return out

    def probe_brick_index(self, xyz, lod):
        xyz = np.ascontiguousarray(xyz, np.float32)
        lod = np.ascontiguousarray(lod, np.uint32)
        idx = np.zeros(lod.size, np.uint32)
        ser = np.zeros(lod.size, np.uint32)
        self._ck(self.lib.tvk_probe_brick_index(self.h, _p(xyz), _p(lod), C.c_uint32(lod.size), _p(idx), _p(ser)),
                 "tvk_probe_brick_index")
        return idx, ser

    # --------------------------------------------------------------- builder
    @staticmethod
    def build_plan(volume_size, brick=64, overlap=2, bpv=1):
        d = L.BuildDesc()
        d.volume_size[:] = [int(v) for v in volume_size]
        d.brick_size, d.overlap, d.bytes_per_voxel = brick, overlap, bpv
        info = L.BuildInfo()
        rc = L.load().tvk_build_plan(C.byref(d), C.byref(info))
        if rc != L.TVK_OK:
            raise L.TvkError("tvk_build_plan failed with status %d" % rc)
        return info

    # ---------------------------------------------------------- re-bricking (DynamicBrickingDS)
    @staticmethod
    def rebrick_desc(lod_sizes, src_brick, tgt_brick, overlap=2, bpv=1):
        d = L.RebrickDesc()
        d.lod_count = len(lod_sizes)
        for l, s in enumerate(lod_sizes):
            d.lod_size[l][:] = [int(v) for v in s]
        d.src_brick[:] = [int(v) for v in src_brick]
        d.tgt_brick[:] = [int(v) for v in tgt_brick]
        d.overlap, d.bytes_per_voxel = overlap, bpv
        return d

    @staticmethod
    def rebrick_plan(desc):
        """-> RebrickInfo (layout only, no device needed); TvkError when the reference would refuse the sizes"""
        info = L.RebrickInfo()
        rc = L.load().tvk_rebrick_plan(C.byref(desc), C.byref(info))
        if rc != L.TVK_OK:
            raise L.TvkError("tvk_rebrick_plan failed with status %d (target core must divide source core, overlap 2)" % rc)
        return info

    @staticmethod
    def rebrick_toc(desc):
        """-> (offsets u64[n], sizes u32[n,3], src_index u64[n], src_offset u32[n,3]) of the target bricks"""
        info = TvkContext.rebrick_plan(desc)
        n = int(info.total_bricks)
        off, sz = np.zeros(n, np.uint64), np.zeros((n, 3), np.uint32)
        si, so = np.zeros(n, np.uint64), np.zeros((n, 3), np.uint32)
        rc = L.load().tvk_rebrick_toc(C.byref(desc), _p(off), _p(sz), _p(si), _p(so), C.c_uint64(n))
        if rc != L.TVK_OK:
            raise L.TvkError("tvk_rebrick_toc failed with status %d" % rc)
        return off, sz, si, so

    def rebrick(self, desc, src_store, src_offsets, src_sizes):
        """src_store: numpy uint8 array holding every source brick (brick i at src_offsets[i], src_sizes[i] voxels).
        -> (target store as numpy uint8 [256-byte aligned bricks], offsets, sizes, minmax f64[n,2])"""
        import torch
        info = self.rebrick_plan(desc)
        off, sz, _, _ = self.rebrick_toc(desc)
        dev = torch.device("cuda", self.device)
        src = torch.from_numpy(np.ascontiguousarray(src_store)).to(dev)
        dst = torch.zeros(int(info.total_bytes), dtype=torch.uint8, device=dev)
        so = np.ascontiguousarray(src_offsets, np.uint64)
        ss = np.ascontiguousarray(src_sizes, np.uint32).reshape(-1)
        mm = np.zeros((int(info.total_bricks), 2), np.float64)
        self._ck(self.lib.tvk_rebrick(self.h, C.byref(desc), C.c_void_p(src.data_ptr()), _p(so), _p(ss), C.c_uint64(so.size),
                                      C.c_void_p(dst.data_ptr()), _p(mm)), "tvk_rebrick")
        return dst.cpu().numpy(), off, sz, mm

    def build_octree(self, raw=None, brick=64, overlap=2, median=False, clamp=False, raw_dptr=None,
                     volume_size=None, bpv=None):
        d = L.BuildDesc()
        if raw is not None:
            raw = np.ascontiguousarray(raw)
            volume_size = raw.shape[::-1]
            bpv = raw.dtype.itemsize
            ptr, on_dev = C.c_void_p(raw.ctypes.data), 0
        else:
            ptr, on_dev = C.c_void_p(int(raw_dptr)), 1
        d.volume_size[:] = [int(v) for v in volume_size]
        d.brick_size, d.overlap, d.bytes_per_voxel = brick, overlap, bpv
        d.use_median, d.clamp_to_edge = int(median), int(clamp)
        info = L.BuildInfo()
        self._ck(self.lib.tvk_build_octree(self.h, C.byref(d), ptr, C.c_int(on_dev), C.byref(info)), "tvk_build_octree")
        self.build_info = info
        self.build_bpv = bpv
        return info

    # ---- streamed build: LOD 0 fed in z-slabs of whole brick rows (tvk_build_stream_*)
    def _desc(self, volume_size, bpv, brick, overlap, median=False):
        d = L.BuildDesc()
        d.volume_size[:] = [int(v) for v in volume_size]
        d.brick_size, d.overlap, d.bytes_per_voxel = brick, overlap, bpv
        d.use_median, d.clamp_to_edge = int(median), 0
        return d

    def build_stream_begin(self, volume_size, bpv, brick=64, overlap=2, median=False, store_ptr=None):
        """store_ptr: None = device store owned by the context; else caller-owned device / mapped pinned host memory"""
        self._stream_desc = self._desc(volume_size, bpv, brick, overlap, median)
        info = L.BuildInfo()
        self._ck(self.lib.tvk_build_stream_begin(self.h, C.byref(self._stream_desc), C.c_void_p(int(store_ptr) if store_ptr else None),
                                                 C.byref(info)), "tvk_build_stream_begin")
        self.build_info, self.build_bpv = info, bpv
        return info

    def build_stream_rows(self, bz0, bz1):
        """raw rows (z0, count) the slab for LOD-0 brick rows [bz0, bz1) must hold"""
        z0, rows = C.c_uint32(), C.c_uint32()
        self._ck(self.lib.tvk_build_stream_rows(C.byref(self._stream_desc), C.c_uint32(bz0), C.c_uint32(bz1), C.byref(z0), C.byref(rows)),
                 "tvk_build_stream_rows")
        return int(z0.value), int(rows.value)

    def build_stream_slab(self, slab_dptr, z0, rows, bz0, bz1):
        self._ck(self.lib.tvk_build_stream_slab(self.h, C.c_void_p(int(slab_dptr)), C.c_uint32(z0), C.c_uint32(rows), C.c_uint32(bz0),
                                                C.c_uint32(bz1)), "tvk_build_stream_slab")

    def build_stream_end(self):
        info = L.BuildInfo()
        self._ck(self.lib.tvk_build_stream_end(self.h, C.byref(info)), "tvk_build_stream_end")
        self.build_info = info
        return info

    def build_toc(self):
        n = int(self.build_info.total_bricks)
        off = np.zeros(n, np.uint64)
        sizes = np.zeros((n, 3), np.uint32)
        self._ck(self.lib.tvk_build_get_toc(self.h, _p(off), _p(sizes), C.c_uint64(n)), "tvk_build_get_toc")
        return off, sizes

    def build_minmax(self):
        n = int(self.build_info.total_bricks)
        mm = np.zeros((n, 2), np.float64)
        self._ck(self.lib.tvk_build_get_minmax(self.h, _p(mm), C.c_uint64(n)), "tvk_build_get_minmax")
        return mm

    def build_bricks(self, offset=0, nbytes=None):
        nbytes = int(self.build_info.total_bytes) - offset if nbytes is None else nbytes
        out = np.zeros(nbytes, np.uint8)
        self._ck(self.lib.tvk_build_get_bricks(self.h, _p(out), C.c_uint64(offset), C.c_uint64(nbytes)),
                 "tvk_build_get_bricks")
        return out

    def build_store_ptr(self):
        p = C.c_void_p()
        self._ck(self.lib.tvk_build_store_device_ptr(self.h, C.byref(p)), "tvk_build_store_device_ptr")
        return p.value

    def build_hist1d(self):
        cap = 1 << (8 * self.build_bpv)
        bins = np.zeros(cap, np.uint64)
        n = C.c_uint32()
        self._ck(self.lib.tvk_build_hist1d(self.h, _p(bins), C.c_uint32(cap), C.byref(n)), "tvk_build_hist1d")
        return bins[:n.value].copy()

    def build_hist2d(self, bins_x, max_nonzero):
        bins = np.zeros((bins_x, 256), np.uint64)
        mg = C.c_float()
        self._ck(self.lib.tvk_build_hist2d(self.h, C.c_uint32(bins_x), C.c_double(max_nonzero), _p(bins), C.byref(mg)),
                 "tvk_build_hist2d")
        return bins, mg.value

    def build_last_ms(self):
        """CUDA-event time of the builder kernels of the last build_octree (ms)."""
        ms = C.c_float()
        self._ck(self.lib.tvk_build_last_ms(self.h, C.byref(ms)), "tvk_build_last_ms")
        return ms.value

    def build_hist2d_last_ms(self):
        ms = C.c_float()
        self._ck(self.lib.tvk_build_hist2d_last_ms(self.h, C.byref(ms)), "tvk_build_hist2d_last_ms")
        return ms.value

    def build_release(self):
        self._ck(self.lib.tvk_build_release(self.h), "tvk_build_release")

    # ---------------------------------------------------------- quantisation
    def quantize(self, data=None, to8bit=False, src_dptr=None, dtype=None, count=None, dst_dptr=None):
        """RAWConverter's quantize() on the device.  data: numpy array (copied up), or src_dptr/dtype/count
        for a volume already in HBM.  Returns dict(out_bits, generated, hist, device_ms, data): `data` is the
        quantised stream fetched back unless dst_dptr (a caller-owned device buffer of 2*count bytes) is given."""
        import torch
        if data is not None:
            data = np.ascontiguousarray(data).reshape(-1)
            dtype, count = data.dtype, data.size
            src, on_dev = _p(data), 0
        else:
            src, on_dev = C.c_void_p(int(src_dptr)), 1
        code = L.DTYPE_CODES[np.dtype(dtype).name]
        own = None
        if dst_dptr is None:
            own = torch.empty(2 * int(count) + 16, dtype=torch.uint8, device="cuda")
            dst_dptr = own.data_ptr()
        hist = np.zeros(4096, np.uint64)
        r = L.QuantizeResult()
        self._ck(self.lib.tvk_quantize(self.h, C.c_uint32(code), C.c_uint64(int(count)), src, C.c_int(on_dev),
                                       C.c_int(1 if to8bit else 0), C.c_void_p(int(dst_dptr)), _p(hist), C.byref(r)), "tvk_quantize")
        out = dict(out_bits=int(r.out_bits), generated=bool(r.generated), device_ms=float(r.device_ms),
                   hist=hist[:r.hist_size].copy() if r.hist_size else None, bin_count=int(r.bin_count), data=None)
        if own is not None and r.generated:
            nb = int(count) * (r.out_bits // 8)
            raw = own[:nb].cpu().numpy()
            out["data"] = raw.copy() if r.out_bits == 8 else raw.view(np.uint16).copy()
        return out

    # ------------------------------------------------------------ compositor
    def composite_ordered(self, dptrs, pixels, out_dptr):
        n = len(dptrs)
        arr = (C.c_void_p * n)(*[int(p) for p in dptrs])
        self._ck(self.lib.tvk_composite_ordered(self.h, arr, C.c_uint32(n), C.c_uint64(pixels), C.c_void_p(int(out_dptr))),
                 "tvk_composite_ordered")

    # ------------------------------------------- sort-last exchange over peer memory
    def exchange_create(self, world, rank, width, height):
        """returns the 64-byte CUDA IPC handle of this rank's exchange region"""
        h = C.create_string_buffer(64)
        self._ck(self.lib.tvk_exchange_create(self.h, C.c_uint32(world), C.c_uint32(rank), C.c_uint32(width), C.c_uint32(height), h),
                 "tvk_exchange_create")
        self._xchg_world = world
        return h.raw

    def exchange_open_peers(self, handles):
        buf = b"".join(handles)
        assert len(buf) == 64 * self._xchg_world
        self._ck(self.lib.tvk_exchange_open_peers(self.h, C.c_char_p(buf)), "tvk_exchange_open_peers")

    def exchange_local_base(self):
        p = C.c_void_p()
        self._ck(self.lib.tvk_exchange_local_base(self.h, C.byref(p)), "tvk_exchange_local_base")
        return p.value

    def exchange_set_peers(self, bases):
        arr = (C.c_void_p * len(bases))(*[int(b) for b in bases])
        self._ck(self.lib.tvk_exchange_set_peers(self.h, arr), "tvk_exchange_set_peers")

    def exchange_connect(self, handles=None, bases=None):
        """handles: list of 64-byte IPC handles (None entries allowed where bases gives a pointer); bases: list of pointers
        valid in this process (None where the handle is to be opened)"""
        w = self._xchg_world
        hb = None
        if handles is not None:
            hb = b"".join(h if h is not None else b"\0" * 64 for h in handles)
            assert len(hb) == 64 * w
        arr = None
        if bases is not None:
            assert len(bases) == w
            arr = (C.c_void_p * w)(*[int(b) if b else None for b in bases])
        self._ck(self.lib.tvk_exchange_connect(self.h, C.c_char_p(hb) if hb is not None else None, arr), "tvk_exchange_connect")

    def exchange_peer_base(self, r):
        p = C.c_void_p()
        self._ck(self.lib.tvk_exchange_peer_base(self.h, C.c_uint32(r), C.byref(p)), "tvk_exchange_peer_base")
        return p.value

    def exchange_begin_frame(self, frame_id):
        self._ck(self.lib.tvk_exchange_begin_frame(self.h, C.c_uint32(frame_id)), "tvk_exchange_begin_frame")

    def exchange_composite(self, order, frame_id):
        if len(order) != self._xchg_world:                   # tvk_exchange_composite reads exactly `world` entries
            raise ValueError("exchange_composite: order has %d entries, the exchange has %d members" % (len(order), self._xchg_world))
        arr = (C.c_uint32 * len(order))(*[int(o) for o in order])
        self._ck(self.lib.tvk_exchange_composite(self.h, arr, C.c_uint32(frame_id)), "tvk_exchange_composite")

    def exchange_final(self, frame_id):
        p = C.c_void_p()
        self._ck(self.lib.tvk_exchange_final(self.h, C.c_uint32(frame_id), C.byref(p)), "tvk_exchange_final")
        return p.value

    def exchange_read_final_rgba8(self, frame_id, out, flip_rows=False):
        """rank 0: the finished sort-last frame as RGBA8 into `out` ([h][w][4] uint8 numpy array or a pinned torch tensor's
        data pointer wrapped in numpy)"""
        self._ck(self.lib.tvk_exchange_read_final_rgba8(self.h, C.c_uint32(frame_id), _p(out), C.c_int(int(flip_rows))),
                 "tvk_exchange_read_final_rgba8")
        return out

    def exchange_timed_out(self):
        t = C.c_uint32()
        self._ck(self.lib.tvk_exchange_status(self.h, C.byref(t)), "tvk_exchange_status")
        return bool(t.value)

    def exchange_destroy(self):
        self._ck(self.lib.tvk_exchange_destroy(self.h), "tvk_exchange_destroy")

    def image_device_ptr(self):
        p = C.c_void_p()
        self._ck(self.lib.tvk_image_device_ptr(self.h, C.byref(p)), "tvk_image_device_ptr")
        return p.value
