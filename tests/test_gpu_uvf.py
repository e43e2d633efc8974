"""GPU: the UVF file boundary on both sides of the hot path.
 * K2 builder output written by CudaGridLeaper::SaveUVF must equal, BYTE FOR BYTE, the UVF the reference wrote
   for the same raw volume (tests/golden/ref_*.uvf.gz, from oracle/_ref/ref_driver writeuvf);
 * CudaGridLeaper::LoadDataset on a reference-written (zlib / lz4) UVF pages those bricks and renders the same
   image as the renderer that built the octree from the raw volume itself."""
import numpy as np
import pytest

from oracle import host as H
from tests.test_uvf import CASES, case_volume, golden_file

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ref_u8_none", "ref_u8_zlib", "ref_u16_zlib"])
def test_builder_to_uvf_is_byte_identical_to_reference(name, tmp_path):
    from imagevis3d_b200.renderer import Renderer
    vol, dt, brick, overlap, comp, level = CASES[name]
    r = Renderer(0)
    r.loadRawVolume(case_volume(name), brick_size=brick, overlap=overlap)
    out = str(tmp_path / "ours.uvf")
    r.saveUVF(out, compression=comp, level=level)
    ours, ref = open(out, "rb").read(), open(golden_file(name, tmp_path), "rb").read()
    assert len(ours) == len(ref) and ours == ref
    r.cleanup()


@pytest.mark.parametrize("name", ["ref_u8_zlib", "ref_u8_lz4", "ref_u16_zlib"])
def test_load_reference_uvf_renders_like_raw(name, tmp_path):
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    vol, dt, brick, overlap, comp, level = CASES[name]
    raw = case_volume(name)
    imgs = []
    for src in ("uvf", "raw"):
        r = Renderer(0)
        if src == "uvf":
            r.loadDataset(golden_file(name, tmp_path))
        else:
            r.loadRawVolume(raw, brick_size=brick, overlap=overlap, hist2d=True)
        info = r.datasetInfo()
        assert list(info.domain) == list(vol)
        tf = H.tf1d_std(info.tf_size, 0.4, 0.4)
        tf[:, 3] *= 0.2
        r.set1DTransFloat(tf)
        r.setLightingEnabled(True)
        r.resize((160, 120))
        r.setRendererTarget(RT_HEADLESS)
        r.setRotation4x4(H.rotation_y(0.5) @ H.rotation_x(0.2))
        r.paintUntilConverged()
        imgs.append((r.frameFloat().copy(), info.tf_size, tuple(info.range), info.max_grad_magnitude))
        r.cleanup()
    assert imgs[0][1:] == imgs[1][1:]
    assert imgs[0][0].max() > 0.05 and np.array_equal(imgs[0][0], imgs[1][0])


def test_load_legacy_rasterdatablock_file(tmp_path):
    """renderer.loadDataset on a legacy single-brick RDB file (c60.uvf's layout, rebuilt from tests/golden/c60.npz):
    the volume is re-bricked to 64^3 by the K2 builder and renders like the raw volume loaded directly."""
    import os
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    from tests.util import write_rdb_uvf
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c60.npz"))
    vol = np.ascontiguousarray(z["brick"])
    path = str(tmp_path / "c60_like.uvf")
    write_rdb_uvf(path, vol, 256, 4)
    imgs = []
    for src in ("uvf", "raw"):
        r = Renderer(0)
        if src == "uvf":
            r.loadDataset(path)
        else:
            r.loadRawVolume(vol, brick_size=64, hist2d=True)
        info = r.datasetInfo()
        assert list(info.domain) == [128, 128, 128] and info.tf_size == 256
        r.set1DTrans(np.ascontiguousarray(z["tf1d_rgba8"]))
        r.resize((256, 256))
        r.setRendererTarget(RT_HEADLESS)
        r.paintUntilConverged()
        imgs.append(r.frameFloat().copy())
        r.cleanup()
    assert imgs[0].max() > 0.3 and np.array_equal(imgs[0], imgs[1])


@pytest.mark.parametrize("kind", ["gridleaper", "raycaster"])
def test_batch_renderer_script(kind, tmp_path):
    """the BatchRenderer flow (loadDataset -> auto-loaded .1dt -> paint until converged -> captureSingleFrame) on a
    c60-like legacy UVF, both renderer types; the PNG must decode to the renderer's own 8-bit frame."""
    import os
    import struct
    import zlib
    from imagevis3d_b200 import batch
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    from tests.util import write_rdb_uvf
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c60.npz"))
    path = str(tmp_path / "c60.uvf")
    write_rdb_uvf(path, np.ascontiguousarray(z["brick"]), 256, 4)
    tfb = np.ascontiguousarray(z["tf1d_rgba8"]).astype(np.float32) / 255.0
    with open(str(tmp_path / "c60.1dt"), "w") as f:                     # TransferFunction1D text format
        f.write("%d\n" % len(tfb) + "".join("%.9g %.9g %.9g %.9g\n" % tuple(r) for r in tfb))
    out = str(tmp_path / "render.png")
    assert batch.main([path, out, "--size", "320x240", "--renderer", kind, "--rotate-y", "25"]) == 0
    raw = open(out, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    w, h = struct.unpack(">II", raw[16:24])
    assert (w, h) == (320, 240)
    idat = b""
    o = 8
    while o < len(raw):
        n, typ = struct.unpack(">I4s", raw[o:o + 8])
        if typ == b"IDAT":
            idat += raw[o + 8:o + 8 + n]
        o += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 4 * w)
    assert (rows[:, 0] == 0).all()                                       # filter type 0 on every row
    img = rows[:, 1:].reshape(h, w, 4)
    assert img[..., 3].max() == 255 and 0.05 < (img[..., 3] > 0).mean() < 0.9


def _gpu_inflate(tvk, blocks, sizes, codec="lz4"):
    import ctypes as C
    n = len(blocks)
    bufs = [np.frombuffer(b, np.uint8).copy() for b in blocks]
    outs = [np.zeros(s, np.uint8) for s in sizes]
    src = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
    dst = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    fn = tvk.lib.tvk_zlib_inflate if codec == "zlib" else tvk.lib.tvk_lz4_inflate
    rc = fn(tvk.h, C.c_uint32(n), src, (C.c_uint32 * n)(*[len(b) for b in blocks]), dst, (C.c_uint32 * n)(*sizes))
    return rc, outs


def test_lz4_inflate_on_the_device(tvk):
    """tvk_lz4_inflate (one warp per block): random, run-length (matches that overlap their own output, offsets 1..40), smooth
    volume data and an incompressible block -- against the bytes that went in."""
    from tests.util import lz4_compress_block
    rng = np.random.default_rng(3)
    datas = [rng.integers(0, 4, 70000, dtype=np.uint8).tobytes(),                       # short alphabet: many matches
             bytes(20000),                                                              # one long run (offset 1)
             (bytes(range(37)) * 900),                                                  # period 37 > 32: the chunked copy
             (bytes(range(7)) * 3000) + rng.integers(0, 256, 999, dtype=np.uint8).tobytes(),
             H.synth_blobs((64, 64, 40), np.uint8).tobytes(),
             rng.integers(0, 256, 50000, dtype=np.uint8).tobytes(), b"abc", b""]
    blocks = [lz4_compress_block(d) for d in datas]
    assert len(blocks[1]) < 200 and len(blocks[2]) < 400                                 # they really are matches
    rc, outs = _gpu_inflate(tvk, blocks, [len(d) for d in datas])
    assert rc == 0
    for d, o in zip(datas, outs):
        assert o.tobytes() == d


def test_lz4_inflate_rejects_corrupt_blocks(tvk):
    from tests.util import lz4_compress_block
    good = lz4_compress_block(bytes(range(50)) * 100)
    bad_offset = bytearray(good)
    bad_offset[len(bytes(range(50))) + 2:len(bytes(range(50))) + 4] = b"\xff\xff"       # a match before the start of the output
    for blk, size in ((bytes(bad_offset), 5000), (good[:len(good) // 2], 5000), (good, 9000)):
        rc, _ = _gpu_inflate(tvk, [blk], [size])
        assert rc != 0


def test_reference_lz4_file_inflates_on_the_device(tvk, tmp_path):
    """the bricks of a reference-written LZ4 UVF, taken from the file still compressed and inflated by the kernel, equal the
    host decode of the same file"""
    from imagevis3d_b200 import uvf as U
    path = golden_file("ref_u8_lz4", tmp_path)
    f = U.UvfFile(path)
    off, ln, comp, sizes = f.toc()
    assert (comp == 3).all()                     # TVK_UVF_COMP_LZ4
    want = f.read_bricks()
    raw = f.read_bricks_raw()
    blocks, pos = [], 0
    for k in ln:
        blocks.append(raw[pos:pos + int(k)].tobytes())
        pos += int(k)
    nbytes = [int(s[0]) * int(s[1]) * int(s[2]) for s in sizes]
    rc, outs = _gpu_inflate(tvk, blocks, nbytes)
    assert rc == 0
    assert np.array_equal(np.concatenate(outs), want)
    f.close()


def test_paging_from_a_compressed_store_renders_the_same(tmp_path):
    """renderer.loadDataset(keep_compressed=True): the pinned store holds the LZ4 bytes of the file, paging moves them
    over PCIe compressed and the device inflates them -- same frame, bit for bit, as the host-inflated load."""
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    path = golden_file("ref_u8_lz4", tmp_path)
    imgs = []
    for keep in (False, True):
        r = Renderer(0)
        r.loadDataset(path, gpu_mem_bytes=24 * 36 ** 3, keep_compressed=keep)      # a pool smaller than the dataset: real paging
        assert r.storeIsCompressed() == keep
        info = r.datasetInfo()
        tf = H.tf1d_std(info.tf_size, 0.4, 0.4)
        tf[:, 3] *= 0.2
        r.set1DTransFloat(tf)
        r.setLightingEnabled(True)
        r.resize((160, 120))
        r.setRendererTarget(RT_HEADLESS)
        r.setRotation4x4(H.rotation_y(0.5) @ H.rotation_x(0.2))
        n = r.paintUntilConverged()
        assert n > 1 and r.perf()["paged_bricks"] > 0
        imgs.append(r.frameFloat().copy())
        r.cleanup()
    assert imgs[0].max() > 0.05 and np.array_equal(imgs[0], imgs[1])


def test_zlib_inflate_on_the_device(tvk):
    """tvk_zlib_inflate (one warp per stream, DEFLATE on the device): stored, fixed-code and dynamic-code blocks, long and
    short match distances, several blocks per stream, brick-sized volume data -- against the bytes that went in."""
    import zlib
    rng = np.random.default_rng(7)
    vol8 = H.synth_blobs((64, 64, 64), np.uint8).tobytes()
    vol16 = H.synth_blobs((64, 64, 64), np.uint16).tobytes()                             # a whole 64^3 uint16 brick: 512 KB
    datas = [rng.integers(0, 4, 70000, dtype=np.uint8).tobytes(), bytes(20000), bytes(range(37)) * 900,
             (bytes(range(7)) * 3000) + rng.integers(0, 256, 999, dtype=np.uint8).tobytes(), vol8, vol16,
             rng.integers(0, 256, 50000, dtype=np.uint8).tobytes(), b"abc", b"a" * 300]
    blocks, sizes, want = [], [], []
    for d in datas:
        for level in (0, 1, 6, 9):                                                       # 0: stored blocks
            blocks.append(zlib.compress(d, level))
            sizes.append(len(d))
            want.append(d)
    co = zlib.compressobj(9, zlib.DEFLATED, 15, 9, zlib.Z_FIXED)                         # fixed Huffman codes only
    blocks.append(co.compress(vol8) + co.flush()); sizes.append(len(vol8)); want.append(vol8)      # noqa: E702
    co = zlib.compressobj(6)                                                             # several blocks, sync flushes in between
    parts = b"".join(co.compress(vol16[i:i + 50000]) + co.flush(zlib.Z_FULL_FLUSH) for i in range(0, len(vol16), 50000)) + co.flush()
    blocks.append(parts); sizes.append(len(vol16)); want.append(vol16)                   # noqa: E702
    rc, outs = _gpu_inflate(tvk, blocks, sizes, "zlib")
    assert rc == 0, tvk.lib.tvk_last_error(tvk.h)
    for k, (d, o) in enumerate(zip(want, outs)):
        assert o.tobytes() == d, k


def test_zlib_inflate_rejects_corrupt_streams(tvk):
    import zlib
    d = H.synth_blobs((40, 40, 40), np.uint8).tobytes()
    good = zlib.compress(d, 6)
    flipped = bytearray(good)
    flipped[len(good) // 2] ^= 0x5A
    assert _gpu_inflate(tvk, [good[:len(good) // 2]], [len(d)], "zlib")[0] != 0          # truncated
    assert _gpu_inflate(tvk, [good], [len(d) + 100], "zlib")[0] != 0                     # the ToC promises more bytes than the stream holds
    assert _gpu_inflate(tvk, [good], [len(d) - 100], "zlib")[0] != 0                     # ... or fewer
    assert _gpu_inflate(tvk, [b"\x00\x00" + good[2:]], [len(d)], "zlib")[0] != 0         # not a zlib header
    rc, outs = _gpu_inflate(tvk, [bytes(flipped)], [len(d)], "zlib")                     # a flipped byte: an error, or different bytes --
    assert rc != 0 or outs[0].tobytes() != d                                             # never a crash (bounds are checked)
    assert _gpu_inflate(tvk, [good], [len(d)], "zlib")[0] == 0


@pytest.mark.parametrize("name", ["ref_u8_zlib", "ref_u16_zlib"])
def test_reference_zlib_file_inflates_on_the_device(tvk, tmp_path, name):
    """the bricks of a reference-written zlib UVF, taken from the file still compressed and inflated by the kernel, equal the
    host decode of the same file"""
    from imagevis3d_b200 import uvf as U
    path = golden_file(name, tmp_path)
    f = U.UvfFile(path)
    off, ln, comp, sizes = f.toc()
    assert (comp == 1).all()                     # TVK_UVF_COMP_ZLIB
    want = f.read_bricks()
    raw = f.read_bricks_raw()
    blocks, pos = [], 0
    for k in ln:
        blocks.append(raw[pos:pos + int(k)].tobytes())
        pos += int(k)
    bpv = 2 if "u16" in name else 1
    nbytes = [int(s[0]) * int(s[1]) * int(s[2]) * bpv for s in sizes]
    rc, outs = _gpu_inflate(tvk, blocks, nbytes, "zlib")
    assert rc == 0
    assert np.array_equal(np.concatenate(outs), np.asarray(want).view(np.uint8).reshape(-1))
    f.close()


def test_paging_from_a_zlib_store_renders_the_same(tmp_path):
    """renderer.loadDataset(keep_compressed=True) on a zlib file: the pinned store holds the file's zlib streams, paging moves
    them over PCIe compressed and the device inflates them -- same frame, bit for bit, as the host-inflated load."""
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    path = golden_file("ref_u8_zlib", tmp_path)
    imgs = []
    for keep in (False, True):
        r = Renderer(0)
        r.loadDataset(path, gpu_mem_bytes=24 * 36 ** 3, keep_compressed=keep)
        assert r.storeIsCompressed() == keep
        info = r.datasetInfo()
        tf = H.tf1d_std(info.tf_size, 0.4, 0.4)
        tf[:, 3] *= 0.2
        r.set1DTransFloat(tf)
        r.setLightingEnabled(True)
        r.resize((160, 120))
        r.setRendererTarget(RT_HEADLESS)
        r.setRotation4x4(H.rotation_y(0.5) @ H.rotation_x(0.2))
        n = r.paintUntilConverged()
        assert n > 1 and r.perf()["paged_bricks"] > 0
        imgs.append(r.frameFloat().copy())
        r.cleanup()
    assert imgs[0].max() > 0.05 and np.array_equal(imgs[0], imgs[1])


def test_batch_renderer_script_views(tmp_path):
    """the other things the script can write: an isosurface with the ClearView lens (raycaster), a stereo frame, a slice view
    and the capture mode's MIP -- every call must produce a PNG of the requested size with something in it"""
    import os
    import struct
    from imagevis3d_b200 import batch
    from tests.util import write_rdb_uvf
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c60.npz"))
    path = str(tmp_path / "c60.uvf")
    write_rdb_uvf(path, np.ascontiguousarray(z["brick"]), 256, 4)
    tfb = np.ascontiguousarray(z["tf1d_rgba8"]).astype(np.float32) / 255.0
    with open(str(tmp_path / "c60.1dt"), "w") as f:
        f.write("%d\n" % len(tfb) + "".join("%.9g %.9g %.9g %.9g\n" % tuple(r) for r in tfb))
    for k, extra in enumerate((["--renderer", "raycaster", "--iso", "60", "--clearview", "110"], ["--stereo", "rb", "--eye-dist", "0.05"],
                               ["--slice", "axial:64"], ["--mip", "coronal:30"], ["--iso", "60"])):
        out = str(tmp_path / ("view%d.png" % k))
        assert batch.main([path, out, "--size", "200x160", "--rotate-y", "25"] + extra) == 0
        raw = open(out, "rb").read()
        assert raw[:8] == b"\x89PNG\r\n\x1a\n" and struct.unpack(">II", raw[16:24]) == (200, 160) and len(raw) > 2000, extra
