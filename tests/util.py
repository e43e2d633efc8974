"""Helpers shared by the parity tests (test infrastructure)."""
import os
import struct
import zlib

import numpy as np

from oracle import host as H
from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    z = np.load(os.path.join(GOLDEN, "builder_cases.npz"))
    out = []
    for i in range(int(z["n_cases"])):
        k = "c%d_" % i
        out.append({n: z[k + n] for n in ("raw", "params", "lod_size", "lod_bricks", "lod_aspect", "sizes", "minmax",
                                           "store", "hist1d", "hist2d", "max_grad")})
    return out


def write_png(path, img8):
    h, w, c = img8.shape
    raw = b"".join(b"\x00" + img8[y].tobytes() for y in range(h))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6 if c == 4 else 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


def cuda_store_tight(ctx):
    """Brick store of the CUDA builder re-packed tightly in ToC order (the oracle's layout)."""
    off, sizes = ctx.build_toc()
    blob = ctx.build_bricks()
    bpv = ctx.build_bpv
    parts = []
    for o, s in zip(off, sizes):
        n = int(s[0]) * int(s[1]) * int(s[2]) * bpv
        parts.append(blob[int(o):int(o) + n])
    return np.concatenate(parts), sizes


def mirror_pool(ctx, pool, sampler=0):
    """Creates the CUDA pool with the oracle pool's geometry and copies its residency state."""
    ctx.pool_create(pool.vol, pool.max_brick, pool.overlap, pool.bpv, pool.capacity, pool.lod_count, sampler)
    assert ctx.total_bricks == pool.total
    sync_pool(ctx, pool, set())


def sync_pool(ctx, pool, already):
    """Uploads every resident slot whose (slot, brick) pair is not in `already`, then the metadata."""
    slots, bricks = [], []
    for slot, idx in enumerate(pool.slot_brick):
        if idx < 0 or (slot, int(idx)) in already:
            continue
        x, y, z, lod = pool.index_to_xyzl(int(idx))
        slots.append(slot)
        bricks.append(np.ascontiguousarray(O.brick(pool.oct, pool.toc_index(x, y, z, lod))))
        already.add((slot, int(idx)))
    if slots:
        ctx.pool_upload_host(slots, bricks)
    ctx.set_metadata(pool.meta)
    return already


def to_cuda_params(p):
    """oracle RaycastParams -> product RaycastParams (same C layout)."""
    import ctypes as C
    from imagevis3d_b200 import _lib as L
    q = L.RaycastParams()
    assert C.sizeof(q) == C.sizeof(p)
    C.memmove(C.byref(q), C.byref(p), C.sizeof(p))
    return q


def image_errors(a, b):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return H.psnr(a, b), float(d.max())


def lockstep_render(tvk, params, pool, tf_rgba8, sampler=0, max_subframes=100):
    """Renders one view to convergence with the oracle and the CUDA path side by side.  Paging
    decisions are taken by the oracle's pool manager from the ORACLE's requests and mirrored to
    the device, so that both always march through identical pool states.  Returns a dict."""
    mirror = set()
    tvk.pool_create(pool.vol, pool.max_brick, pool.overlap, pool.bpv, pool.capacity, pool.lod_count, sampler)
    sync_pool(tvk, pool, mirror)
    if tf_rgba8.ndim == 2:
        tvk.set_tf1d(tf_rgba8)
    else:
        tvk.set_tf2d(tf_rgba8)
    frame = H.OracleFrame(params.width, params.height, params.hash_table_size)
    params.new_view = 1
    tvk.stats(reset=True)
    req_mismatch, req_total, subframes, samples_o, rays_o = 0, 0, 0, 0, 0
    req_mismatch_nodrop, drops = 0, 0
    for sub in range(max_subframes):
        H.raycast(params, pool, tf_rgba8, frame)
        samples_o += int(frame.stats[0])
        rays_o += int(frame.stats[2])
        drops += int(frame.stats[3])
        tvk.raycast(to_cuda_params(params))
        params.new_view = 0
        subframes += 1
        req_o = {tuple(int(v) for v in r) for r in H.read_requests(pool, frame)}
        req_c = {tuple(int(v) for v in r) for r in tvk.read_requests()}
        req_total += len(req_o | req_c)
        req_mismatch += len(req_o ^ req_c)
        if int(frame.stats[3]) == 0:
            # no report was dropped for want of a free table entry: the table holds EVERY brick the rays asked for,
            # whatever the order of the atomics -- the two request sets must then be the same set
            req_mismatch_nodrop += len(req_o ^ req_c)
        if not req_o and not req_c:
            break
        pool.upload_requests(H.read_requests(pool, frame) if req_o else np.array(sorted(req_c), np.uint32))
        sync_pool(tvk, pool, mirror)
    else:
        raise AssertionError("did not converge")
    img_c = tvk.read_image()
    pos_c, aux_c = tvk.read_resume()
    st = tvk.stats()
    return dict(cuda=img_c, oracle=frame.image.copy(), subframes=subframes, req_total=req_total,
                req_mismatch=req_mismatch, req_mismatch_nodrop=req_mismatch_nodrop, req_drops=drops,
                samples_cuda=int(st.samples), samples_oracle=samples_o,
                pos_cuda=pos_c, pos_oracle=frame.pos.copy(), rays_cuda=int(st.rays), rays_oracle=rays_o)


def resident_pool(tvk, vol, brick, tf_rgba8, lo, hi, sampler=0):
    """Builds the octree of `vol` with the CUDA builder and makes EVERY pool brick resident (slot i = brick i,
    coarsest brick in the last slot), with 1D-TF visibility for the value range [lo, hi]."""
    dims = vol.shape[::-1]
    info = tvk.build_octree(vol, brick, 2)
    off, sizes = tvk.build_toc()
    mm = tvk.build_minmax()
    lod_count = next(l for l in range(info.lod_count)
                     if info.lod_bricks[l][0] * info.lod_bricks[l][1] * info.lod_bricks[l][2] == 1) + 1
    total = int(info.lod_offset[lod_count - 1]) + 1
    nbx, nby = int(info.lod_bricks[0][0]) + 1, int(info.lod_bricks[0][1]) + 1
    cap = (nbx, nby, -(-total // (nbx * nby)))
    mb = tuple(int(v) for v in sizes[:total].max(axis=0))
    assert tvk.pool_create(dims, mb, 2, vol.dtype.itemsize, cap, lod_count, sampler) == total
    store = tvk.build_store_ptr()
    slots = np.arange(total, dtype=np.uint32)
    slots[total - 1] = cap[0] * cap[1] * cap[2] - 1
    tvk.pool_upload_device(slots, [store + int(o) for o in off[:total]], sizes[:total])
    tvk.set_metadata(slots + 3)
    tvk.set_minmax(mm[:total])
    tvk.set_tf1d(tf_rgba8)
    tvk.recompute_visibility(0, lo, hi)
    tvk.build_release()


def write_rdb_uvf(path, vol, brick=256, overlap=4, block_id="Qvis data volume converted by ImageVis3D"):
    """TEST HELPER: writes a legacy (UVF version 2) file with one single-LOD RasterDataBlock, the layout
    RasterDataBlock::CopyHeaderToFile / ComputeBricks give it (Tuvok/IO/UVF/RasterDataBlock.cpp:210-260,321-349;
    DemoData/c60.uvf is such a file) plus an MD5 checksum -- so that the legacy reader can be tested where
    /root/reference (and with it DemoData) does not exist."""
    import hashlib
    import struct
    vol = np.ascontiguousarray(vol)
    dz, dy, dx = vol.shape
    bits = 8 * vol.dtype.itemsize

    def axis(size):
        out = []
        if size <= brick:
            return [size]
        while True:
            if size + overlap <= brick:
                out.append(size)
                break
            out.append(brick)
            size = size + overlap - brick
            if size <= overlap:
                break
        return out
    ax = [axis(dx), axis(dy), axis(dz)]
    payload = bytearray()
    step = brick - overlap
    for kz, sz in enumerate(ax[2]):
        for ky, sy in enumerate(ax[1]):
            for kx, sx in enumerate(ax[0]):
                payload += vol[kz * step:kz * step + sz, ky * step:ky * step + sy, kx * step:kx * step + sx].tobytes()
    q = lambda *v: struct.pack("<%dQ" % len(v), *v)        # noqa: E731
    hdr = q(3) + q(1, 2, 3) + struct.pack("<16d", *np.eye(4).reshape(-1)) + q(dx, dy, dz) + q(brick, brick, brick)
    hdr += q(overlap, overlap, overlap) + q(2, 2, 2) + q(0, 0, 0) + q(1) + q(1) + q(1) + q(0) + q(bits) + q(bits) + b"\x00" + q(0)
    bid = block_id.encode()
    block = q(len(bid)) + bid + q(1) + q(0) + q(0) + hdr + bytes(payload)          # last block: offsetToNext = 0
    body = q(0) + block                                                             # offsetToFirstDataBlock + blocks
    digest = hashlib.md5(body).digest()
    with open(path, "wb") as f:
        f.write(b"UVF-DATA" + b"\x00" + q(2) + q(2) + q(16) + digest + body)


def lz4_compress_block(data, min_match=4):
    """TEST HELPER: a small greedy LZ4 BLOCK compressor (raw block format, what lz4Compress / LZ4_compress write for a brick):
    valid for any LZ4 decoder -- the last five bytes are literals and no match starts within the last twelve."""
    src = bytes(data)
    n = len(src)
    out = bytearray()
    table = {}
    anchor = i = 0

    def emit(lit_from, lit_to, off, mlen):
        lit = lit_to - lit_from
        token_l = 15 if lit >= 15 else lit
        token_m = 0 if mlen is None else (15 if mlen - 4 >= 15 else mlen - 4)
        out.append((token_l << 4) | token_m)
        if lit >= 15:
            r = lit - 15
            while r >= 255:
                out.append(255)
                r -= 255
            out.append(r)
        out.extend(src[lit_from:lit_to])
        if mlen is not None:
            out.extend((off & 0xFF, off >> 8))
            if mlen - 4 >= 15:
                r = mlen - 4 - 15
                while r >= 255:
                    out.append(255)
                    r -= 255
                out.append(r)

    limit = n - 12
    while i < limit:
        key = src[i:i + 4]
        j = table.get(key)
        table[key] = i
        if j is not None and i - j <= 65535:
            m = 4
            while i + m < n - 5 and src[j + m] == src[i + m]:
                m += 1
            emit(anchor, i, i - j, m)
            i += m
            anchor = i
        else:
            i += 1
    emit(anchor, n, 0, None)
    return bytes(out)
