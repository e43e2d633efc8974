"""CPU: the UVF v5 writer / reader (include/tvk_uvf.h, host code) against the reference.
 * a UVF written by the REFERENCE (tests/golden/ref_*.uvf, made by tests/golden/make_golden_uvf.py from
   oracle/_ref/ref_driver writeuvf = the block assembly of RAWConverter::ConvertRAWDataset) is read by our
   reader and compared with the oracle's octree; and our writer, fed the oracle's arrays, must reproduce that
   file BYTE FOR BYTE (uncompressed and zlib);
 * where oracle/_ref exists (build container) the reference's own UVF::Open(verify) + TOCBlock::GetData read
   the files our writer produced."""
import os
import subprocess

import numpy as np
import pytest

from imagevis3d_b200 import uvf as U
from oracle import host as H
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"ref_u8_none": ((50, 41, 33), np.uint8, 16, 2, 0, 0), "ref_u8_zlib": ((50, 41, 33), np.uint8, 16, 2, 1, 1),
         "ref_u16_zlib": ((37, 20, 45), np.uint16, 20, 2, 1, 1), "ref_u8_lz4": ((50, 41, 33), np.uint8, 16, 2, 3, 1),
         "ref_u8_lzma": ((50, 41, 33), np.uint8, 16, 2, 2, 4), "ref_u16_bzip": ((37, 20, 45), np.uint16, 20, 2, 4, 5)}


def case_volume(name):
    vol, dt, brick, overlap, comp, level = CASES[name]
    raw = H.synth_blobs(vol, dt, seed=7) if comp else H.synth_noise(vol, dt, seed=4)
    if dt == np.uint16:
        raw = (H.synth_blobs(vol, np.uint8, seed=7).astype(np.uint16) + 300)   # < 4096: no quantisation in the reference flow
    return raw


def golden_file(name, tmp_path):
    """the fixtures are committed gzip-compressed (their 2-D histogram blocks are mostly zeros)"""
    import gzip
    out = os.path.join(str(tmp_path), name + ".uvf")
    with gzip.open(os.path.join(GOLD, name + ".uvf.gz"), "rb") as f:
        open(out, "wb").write(f.read())
    return out


def _oracle_parts(raw, brick, overlap):
    oc = O.build_octree(raw, brick, overlap)
    h1 = O.hist1d(oc)
    h2, mg = O.hist2d(oc, len(h1), oc["minmax"][:, 1].max())
    tight = np.concatenate([[0], np.cumsum(oc["sizes"].astype(np.uint64).prod(axis=1) * oc["bpv"])]).astype(np.uint64)
    return oc, h1, h2, mg, tight


def test_md5_known_answers():
    import hashlib
    for msg in (b"", b"a", b"abc", b"message digest", bytes(range(256)) * 37):
        assert U.md5(msg) == hashlib.md5(msg).digest()


@pytest.mark.parametrize("name", sorted(CASES))
def test_reader_on_reference_written_file(name, tmp_path):
    vol, dt, brick, overlap, comp, level = CASES[name]
    raw = case_volume(name)
    oc, h1, h2, mg, tight = _oracle_parts(raw, brick, overlap)
    with U.UvfFile(golden_file(name, tmp_path), verify=True) as f:
        i = f.info
        assert i.uvf_version == 5 and i.checksum_ok == 1 and i.toc_block == 0 and i.octree_version == 2
        assert list(i.block_semantics[:i.block_count]) == [9, 5, 6, 7, 4]     # TOC, Hist1D, Hist2D, MaxMin, KeyValue
        assert list(i.vol.volume_size) == list(vol) and list(i.vol.brick_size) == [brick] * 3 and i.vol.overlap == overlap
        assert i.total_bricks == oc["plan"].total_bricks and i.lod_count == oc["plan"].lod_count
        off, ln, c, sizes = f.toc()
        assert (sizes == oc["sizes"]).all()
        if comp:
            assert (c == comp).any()
        assert (f.read_bricks() == oc["store"]).all()
        mm = f.minmax()
        assert (mm[:, :2] == oc["minmax"]).all() and (mm[:, 2] == -np.finfo(np.float64).max).all()
        assert (f.hist1d() == h1).all() and (f.hist2d() == h2).all()
        assert np.float32(i.max_grad_magnitude) == np.float32(mg)
        assert f.kv() == [("Source Endianness", "little"), ("Source Type", "integer"), ("Source Bitwidth", str(8 * raw.dtype.itemsize))]


@pytest.mark.parametrize("name", ["ref_u8_none", "ref_u8_zlib", "ref_u16_zlib"])
def test_writer_reproduces_reference_file_byte_for_byte(name, tmp_path):
    vol, dt, brick, overlap, comp, level = CASES[name]
    raw = case_volume(name)
    oc, h1, h2, mg, tight = _oracle_parts(raw, brick, overlap)
    out = str(tmp_path / "ours.uvf")
    U.write_uvf(out, vol, raw.dtype.itemsize, brick, overlap, tight[:-1], oc["sizes"], oc["store"], oc["minmax"], h1, h2, mg,
                compression=comp, level=level)
    ours, ref = open(out, "rb").read(), open(golden_file(name, tmp_path), "rb").read()
    assert len(ours) == len(ref)
    assert ours == ref


def test_reader_rejects_corruption(tmp_path):
    from imagevis3d_b200._lib import TvkError
    data = bytearray(open(golden_file("ref_u8_none", tmp_path), "rb").read())
    data[len(data) // 2] ^= 0x40
    bad = str(tmp_path / "bad.uvf")
    open(bad, "wb").write(bytes(data))
    with pytest.raises(TvkError):
        U.UvfFile(bad, verify=True)
    U.UvfFile(bad, verify=False).close()          # like the reference, unverified opens do not look at the payload
    open(bad, "wb").write(b"not a uvf file at all, not even close ........................")
    with pytest.raises(TvkError):
        U.UvfFile(bad)
    with pytest.raises(TvkError):
        U.UvfFile(str(tmp_path / "missing.uvf"))


def test_reader_survives_hostile_sizes(tmp_path):
    """Sizes found in the file are checked against the file before anything is allocated from them, and no C++
    exception crosses the C ABI: a brick-table entry of 2^62 bytes, an entry pointing behind the end of the file, a
    max/min block announcing 2^40 bricks and a RasterDataBlock header with an absurd LOD index all come back as
    errors (or as an unusable block) -- the process must not abort or hang."""
    import struct
    from imagevis3d_b200._lib import TvkError
    good = bytearray(open(golden_file("ref_u8_zlib", tmp_path), "rb").read())
    f = U.UvfFile(golden_file("ref_u8_zlib", tmp_path), verify=False)
    off, length, comp, sizes = f.toc()
    n = len(off)
    f.close()
    # the ToC (36-byte entries: u64 offset, u64 length, u32 compression, u64 valid length, 2 x u32) ends where the
    # first brick starts; entry k's length field sits at table + 36*k + 8
    table = int(off.min()) - 36 * n
    assert struct.unpack_from("<Q", good, table + 8)[0] == int(length[0])
    for value, field in ((1 << 62, 8), (len(good) + 12345, 0)):
        bad = bytearray(good)
        struct.pack_into("<Q", bad, table + 36 * (n // 2) + field, value)
        path = str(tmp_path / ("hostile_%d.uvf" % field))
        open(path, "wb").write(bytes(bad))
        with pytest.raises(TvkError):
            U.UvfFile(path, verify=False)
    # max/min block: brick count 2^40
    pat = b"Brick Max/Min Values"
    at = bytes(good).find(pat)
    assert at > 0
    bad = bytearray(good)
    struct.pack_into("<Q", bad, at + len(pat) + 24, 1 << 40)        # after id: semantics, compression, offset-to-next, then n
    path = str(tmp_path / "hostile_mm.uvf")
    open(path, "wb").write(bytes(bad))
    with pytest.raises(TvkError):
        U.UvfFile(path, verify=False)
    # legacy RasterDataBlock: LOD group count 2^60 (the header loop must not run 2^60 times)
    from tests.util import write_rdb_uvf
    path = str(tmp_path / "rdb.uvf")
    write_rdb_uvf(path, H.synth_blobs((40, 30, 20), np.uint8, seed=3), brick=32, overlap=4)
    raw = bytearray(open(path, "rb").read())
    at = bytes(raw).find(b"Qvis data volume converted by ImageVis3D") + 40 + 24
    groups = at + 8 + 24 + 128 + 24 + 24 + 24 + 24                    # dims, semantics, transform, domain, brick, overlap, decimation
    assert struct.unpack_from("<3Q", raw, groups) == (0, 0, 0)
    struct.pack_into("<Q", raw, groups, 1 << 60)
    open(path, "wb").write(bytes(raw))
    g = U.UvfFile(path, verify=False)                               # opens; the undecodable grid block is skipped
    with pytest.raises(TvkError):
        g.read_rdb_volume()
    g.close()


@pytest.mark.skipif(O.ref_driver_path() is None, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("comp", [0, 1])
def test_reference_reads_our_file(comp, tmp_path):
    vol, brick, overlap = (70, 33, 29), 32, 2
    raw = H.synth_blobs(vol, np.uint8, seed=9)
    oc, h1, h2, mg, tight = _oracle_parts(raw, brick, overlap)
    out = str(tmp_path / "ours.uvf")
    U.write_uvf(out, vol, 1, brick, overlap, tight[:-1], oc["sizes"], oc["store"], oc["minmax"], h1, h2, mg, compression=comp, level=3)
    dump = str(tmp_path / "dump.bin")
    subprocess.check_call([O.ref_driver_path(), "readuvf", out, dump])       # UVF::Open(verify = true) checks our MD5
    r = O.parse_ref_convert(dump, 1)
    assert (r["store"] == oc["store"]).all() and (r["minmax"] == oc["minmax"]).all()
    assert (r["hist1d"] == h1).all() and (r["hist2d"] == h2).all() and np.float32(r["max_grad"]) == np.float32(mg)


@pytest.mark.parametrize("vol,brick,overlap,dt", [((128, 128, 128), 256, 4, np.uint8), ((100, 70, 45), 32, 4, np.uint8),
                                                  ((61, 33, 90), 24, 2, np.uint16)])
def test_legacy_rasterdatablock_reader(vol, brick, overlap, dt, tmp_path):
    """legacy single-LOD RasterDataBlock files (DemoData/c60.uvf is one): header fields and the finest level
    stitched back into the flat volume; multi-brick layouts included."""
    from tests.util import write_rdb_uvf
    raw = H.synth_noise(vol, dt, seed=sum(vol))
    path = str(tmp_path / "legacy.uvf")
    write_rdb_uvf(path, raw, brick, overlap)
    with U.UvfFile(path, verify=True) as f:
        i = f.info
        assert i.uvf_version == 2 and i.toc_block == -1 and i.rdb_block == 0 and i.checksum_ok == 1
        assert list(i.rdb_domain) == list(vol) and list(i.rdb_brick) == [brick] * 3 and list(i.rdb_overlap) == [overlap] * 3
        assert i.rdb_bits == 8 * raw.dtype.itemsize and i.rdb_signed == 0 and i.rdb_lod_levels == 1
        assert (f.read_rdb_volume() == raw).all()


@pytest.mark.skipif(not os.path.exists("/root/reference/DemoData/c60.uvf"), reason="needs /root/reference (build container)")
def test_reads_the_demo_dataset_itself():
    z = np.load(os.path.join(GOLD, "c60.npz"))          # decoded by the reference's own UVF reader (make_golden.py)
    with U.UvfFile("/root/reference/DemoData/c60.uvf", verify=True) as f:
        assert f.info.uvf_version == 2 and list(f.info.rdb_domain) == [128, 128, 128] and list(f.info.rdb_brick) == [256] * 3
        assert (f.read_rdb_volume() == z["brick"]).all() and (f.hist1d() == z["hist1d"]).all()


@pytest.mark.skipif(O.ref_driver_path() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_reference_reads_the_synthesised_legacy_file(tmp_path):
    """pins the test-side RasterDataBlock writer: the reference's UVF::Open(verify) + RasterDataBlock::GetData decode it"""
    from tests.util import write_rdb_uvf
    raw = H.synth_noise((40, 36, 28), np.uint8, seed=3)
    path, dump = str(tmp_path / "legacy.uvf"), str(tmp_path / "dump.bin")
    write_rdb_uvf(path, raw, 256, 4)
    # uvfinfo needs the histogram / max-min blocks c60.uvf has; a bare RDB file makes it report "missing blocks"
    # AFTER the open + checksum verification succeeded -- which is the part pinned here
    rc = subprocess.run([O.ref_driver_path(), "uvfinfo", path, dump], capture_output=True, text=True)
    assert "open failed" not in rc.stderr and "missing blocks" in rc.stderr


@pytest.mark.skipif(O.ref_driver_path() is None, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("vol,dt,brick,comp", [((1, 1, 1), np.uint8, 16, 0), ((1, 1, 1), np.uint16, 16, 1), ((5, 4, 3), np.uint8, 16, 1),
                                               ((13, 1, 1), np.uint16, 8, 0), ((12, 12, 12), np.uint8, 16, 1), ((17, 16, 15), np.uint8, 16, 0)])
def test_degenerate_volumes_byte_identical_to_the_reference(vol, dt, brick, comp, tmp_path):
    """one-voxel, one-brick and one-voxel-thick volumes: the reference writes the file (ref_driver writeuvf), ours must be
    the same bytes, and both must read back to the same bricks"""
    raw = H.synth_noise(vol, dt, seed=sum(vol) + brick, maxval=200 if dt == np.uint8 else 3000)
    rawf, ref = str(tmp_path / "v.raw"), str(tmp_path / "ref.uvf")
    raw.tofile(rawf)
    subprocess.check_call([O.ref_driver_path(), "writeuvf", rawf, str(vol[0]), str(vol[1]), str(vol[2]), "u16" if raw.dtype.itemsize == 2 else "u8",
                           str(brick), "2", "0", str(comp), "1", ref])
    oc, h1, h2, mg, tight = _oracle_parts(raw, brick, 2)
    out = str(tmp_path / "ours.uvf")
    U.write_uvf(out, vol, raw.dtype.itemsize, brick, 2, tight[:-1], oc["sizes"], oc["store"], oc["minmax"], h1, h2, mg, compression=comp, level=1)
    a, b = open(out, "rb").read(), open(ref, "rb").read()
    assert len(a) == len(b) and a == b
    with U.UvfFile(ref, verify=True) as f:
        assert f.info.total_bricks == oc["plan"].total_bricks and (f.read_bricks() == oc["store"]).all()
        assert (f.minmax()[:, :2] == oc["minmax"]).all()
