"""GPU: K2 builder kernels against the golden vectors of the reference and the oracle (bit-exact)."""
import numpy as np
import pytest

from oracle import host as H
from oracle import oracle as O
from tests.util import cuda_store_tight, golden_cases

pytestmark = pytest.mark.gpu


def _compare(tvk, raw, brick, overlap, median, ref_store, ref_sizes, ref_minmax, ref_h1=None, ref_h2=None, ref_mg=None):
    info = tvk.build_octree(raw, brick, overlap, median=bool(median))
    store, sizes = cuda_store_tight(tvk)
    assert (sizes == ref_sizes).all()
    assert store.size == ref_store.size
    bad = np.nonzero(store != ref_store)[0]
    assert bad.size == 0, "first mismatching byte %d of %d" % (bad[0], store.size)
    assert (tvk.build_minmax() == ref_minmax).all()
    if ref_h1 is not None:
        h1 = tvk.build_hist1d()
        assert h1.size == ref_h1.size and (h1 == ref_h1).all()
        h2, mg = tvk.build_hist2d(len(h1), float(ref_minmax[:, 1].max()))
        assert np.float32(mg) == np.float32(ref_mg)
        assert (h2 == ref_h2).all()
    tvk.build_release()
    return info


@pytest.mark.parametrize("case", range(7))
def test_builder_matches_reference_golden(tvk, case):
    g = golden_cases()[case]
    brick, overlap, median = [int(v) for v in g["params"]]
    _compare(tvk, g["raw"], brick, overlap, median, g["store"], g["sizes"], g["minmax"], g["hist1d"], g["hist2d"], g["max_grad"])


@pytest.mark.parametrize("vol,dtype,brick,overlap,median", [
    ((61, 47, 33), np.uint8, 12, 2, False), ((49, 9, 5), np.uint8, 12, 2, True), ((97, 97, 97), np.uint16, 20, 2, False),
    ((130, 121, 64), np.uint8, 64, 2, False), ((200, 150, 100), np.uint16, 64, 2, True), ((50, 26, 14), np.uint16, 16, 4, False)])
def test_builder_matches_oracle(tvk, vol, dtype, brick, overlap, median):
    raw = H.synth_noise(vol, dtype, seed=sum(vol), maxval=255 if dtype == np.uint8 else 4095)
    oc = O.build_octree(raw, brick, overlap, median)
    h1 = O.hist1d(oc)
    h2, mg = O.hist2d(oc, len(h1), oc["minmax"][:, 1].max())
    _compare(tvk, raw, brick, overlap, median, oc["store"], oc["sizes"], oc["minmax"], h1, h2, mg)


@pytest.mark.parametrize("kind", ["noise", "smooth", "flat", "flat_odd", "noise_unaligned"])
def test_builder_u16_full_range_histogram(tvk, kind):
    """Full-range 16-bit volumes: packed shared-memory histogram inside the LOD 0 -> 1 pass (all 65536 bins); flat
    volumes take its one-atomic-per-warp path, odd sizes give the lanes of one warp different row counts; volumes whose
    rows are not 16-byte multiples keep the windowed histogram kernel (two 32768-bin passes)."""
    vol = (128, 97, 73) if kind == "flat_odd" else (128, 96, 72)
    if kind == "noise_unaligned":
        vol = (127, 96, 75)       # rows are not 16-byte multiples: separate reduction + two-window histogram kernels
    if kind in ("noise", "noise_unaligned"):
        raw = H.synth_noise(vol, np.uint16, seed=5, maxval=65535)
    elif kind == "smooth":
        raw = (H.synth_sphere(vol, np.uint8).astype(np.uint32) * 257).astype(np.uint16)
    else:
        raw = np.full(vol[::-1], 40000, np.uint16)
        raw[:, :, :7] = 3
    oc = O.build_octree(raw, 64, 2, False)
    tvk.build_octree(raw, 64, 2)
    store, sizes = cuda_store_tight(tvk)
    assert (store == oc["store"]).all() and (tvk.build_minmax() == oc["minmax"]).all()
    h1 = tvk.build_hist1d()
    ref = np.bincount(raw.reshape(-1), minlength=65536)
    ref = ref[:np.nonzero(ref)[0][-1] + 1]
    assert h1.size == ref.size and (h1 == ref).all()
    tvk.build_release()


def test_builder_u16_packed_histogram_spills(tvk):
    """The LOD 0 -> 1 pass counts 16-bit volumes in packed 16-bit shared-memory counters; a counter that reaches
    0x8000 moves that many counts to the global bin.  Large flat regions (both halves of one word, the top bin,
    zero) make every CTA spill many times; the histogram must still be Histogram1DDataBlock::Compute's
    (IO/UVF/Histogram1DDataBlock.cpp:60-160), i.e. a plain count of the voxels."""
    vol = (512, 256, 192)
    rng = np.random.default_rng(11)
    raw = rng.integers(0, 65536, size=vol[::-1], dtype=np.uint16)
    raw[:40] = 40000
    raw[40:80, :, ::2] = 40001            # alternating with noise: no run-length merging
    raw[80:110] = 65535
    raw[110:130, :128] = 0
    raw[130:150] = (np.arange(512) // 3 + 32760).astype(np.uint16)    # straddles the old window boundary
    tvk.build_octree(raw, 64, 2)
    h1 = tvk.build_hist1d()
    ref = np.bincount(raw.reshape(-1), minlength=65536)
    assert h1.size == 65536 and (h1 == ref).all()
    mm = tvk.build_minmax()
    assert mm[:, 1].max() == 65535 and mm[:, 0].min() == 0
    tvk.build_release()


def test_builder_rejects_unsupported(tvk):
    from imagevis3d_b200._lib import TvkError
    raw = np.zeros((8, 8, 8), np.uint8)
    with pytest.raises(TvkError):
        tvk.build_octree(raw, 4, 2)


@pytest.mark.parametrize("vol,dtype,brick,overlap,median", [
    ((70, 45, 38), np.uint8, 16, 2, False), ((37, 29, 23), np.uint8, 12, 2, True), ((50, 26, 40), np.uint16, 16, 2, False),
    ((9, 7, 5), np.uint8, 12, 2, False), ((8, 8, 6), np.uint8, 16, 2, False), ((130, 121, 70), np.uint16, 64, 2, False)])
def test_builder_clamp_to_edge_matches_oracle(tvk, vol, dtype, brick, overlap, median):
    """bClampToEdge (ExtendedOctreeConverter.cpp:376-462): bricks, min/max and both histograms equal to the oracle's, which
    tests/test_oracle_builder.py pins to the compiled reference on every voxel the reference defines in that mode."""
    raw = H.synth_noise(vol, dtype, seed=sum(vol) + 7, maxval=255 if dtype == np.uint8 else 60000)
    raw[raw == 0] = 1                                          # a zero in a ghost cell would be a border that was not clamped
    oc = O.build_octree(raw, brick, overlap, median, clamp=True)
    tvk.build_octree(raw, brick, overlap, median=median, clamp=True)
    store, sizes = cuda_store_tight(tvk)
    assert (sizes == oc["sizes"]).all()
    bad = np.nonzero(store != oc["store"])[0]
    assert bad.size == 0, "first mismatching byte %d of %d" % (bad[0], store.size)
    assert store.view(dtype).min() > 0
    assert (tvk.build_minmax() == oc["minmax"]).all()
    h1 = tvk.build_hist1d()
    ref1 = O.hist1d(oc)
    assert h1.size == ref1.size and (h1 == ref1).all()
    h2, mg = tvk.build_hist2d(len(h1), float(oc["minmax"][:, 1].max()))
    r2, rmg = O.hist2d(oc, len(h1), oc["minmax"][:, 1].max())
    assert np.float32(mg) == np.float32(rmg) and (h2 == r2).all()
    tvk.build_release()


def test_builder_full_size_properties(tvk):
    """512^3 uint8 (BASELINE config 2 size): size-independent properties instead of a CPU re-run."""
    vol = (512, 512, 512)
    raw = H.synth_sphere(vol, np.uint8)
    info = tvk.build_octree(raw, 64, 2)
    assert info.total_bricks == 895 and info.lod_count == 10 and info.lod_offset[4] + 1 == 890
    h1 = tvk.build_hist1d()
    assert h1.sum() == 512 ** 3                                   # every voxel counted once
    assert (h1 == np.bincount(raw.reshape(-1), minlength=256)[:len(h1)]).all()
    mm = tvk.build_minmax()
    assert mm[:729, 1].max() == raw.max() and mm[:, 0].min() == 0
    off, sizes = tvk.build_toc()
    # brick (4,4,4) of LOD 0 is the raw sub-volume [238..302)^3
    i = 4 + 4 * 9 + 4 * 81
    n = int(np.prod(sizes[i]))
    b = tvk.build_bricks(int(off[i]), n).reshape(64, 64, 64)
    assert (b == raw[238:302, 238:302, 238:302]).all()
    # mean filter: LOD 1 inner voxel == floor(mean of the 8 children)
    j = int(info.lod_offset[1])
    b1 = tvk.build_bricks(int(off[j]), int(np.prod(sizes[j]))).reshape(64, 64, 64)
    blk = raw[:120, :120, :120].astype(np.uint32).reshape(60, 2, 60, 2, 60, 2).sum(axis=(1, 3, 5)) // 8
    assert (b1[2:62, 2:62, 2:62] == blk).all()
    tvk.build_release()


@pytest.mark.parametrize("vol,dtype,median", [((130, 121, 64), np.uint8, False), ((200, 150, 100), np.uint16, True), ((96, 96, 96), np.uint16, False)])
def test_fused_lod0_pass_is_bit_exact(vol, dtype, median, monkeypatch):
    """TVK_BUILD_FUSED=1: bricks + min/max + 1-D histogram + LOD 1 from one read of the raw volume (opt-in path)."""
    from imagevis3d_b200.context import TvkContext
    monkeypatch.setenv("TVK_BUILD_FUSED", "1")
    t = TvkContext(0)
    raw = H.synth_noise(vol, dtype, seed=sum(vol), maxval=255 if dtype == np.uint8 else 50000)
    oc = O.build_octree(raw, 64, 2, median)
    t.build_octree(raw, 64, 2, median=median)
    store, sizes = cuda_store_tight(t)
    assert (store == oc["store"]).all() and (t.build_minmax() == oc["minmax"]).all()
    h1 = t.build_hist1d()
    ref = np.bincount(raw.reshape(-1), minlength=1 << (8 * raw.dtype.itemsize))
    ref = ref[:np.nonzero(ref)[0][-1] + 1]
    assert h1.size == ref.size and (h1 == ref).all()
    t.close()


def _hist2d_volumes():
    rng = np.random.default_rng(11)
    out = []
    out.append(("u8 noise", H.synth_noise((132, 37, 21), np.uint8, seed=3, maxval=255)))
    out.append(("u8 sphere", H.synth_sphere((64, 64, 64), np.uint8)))
    z, y, x = np.meshgrid(np.arange(20), np.arange(24), np.arange(128), indexing="ij")
    out.append(("u8 ramp (every voxel ties with the maximum)", ((x + 2 * y + z) % 256).astype(np.uint8)))
    out.append(("u8 flat", np.full((9, 10, 32), 77, np.uint8)))
    out.append(("u16 noise", H.synth_noise((136, 30, 17), np.uint16, seed=9, maxval=65535)))
    out.append(("u16 12-bit blobs", (H.synth_blobs((96, 52, 40), np.uint16) >> 4).astype(np.uint16)))
    sm = rng.integers(0, 4096, (12, 16, 64)).astype(np.uint16)
    sm[:, :, 20:40] = 4000                                   # runs of equal bins inside a thread's four voxels
    out.append(("u16 runs", sm))
    # rows of several 256-voxel chunks, the last one a single octet: x halos at chunk borders come from memory
    out.append(("u16 wide noise", rng.integers(0, 65536, (5, 9, 520)).astype(np.uint16)))
    out.append(("u8 wide noise", rng.integers(0, 256, (5, 11, 776)).astype(np.uint8)))
    out.append(("u16 sphere (the bench generator: S nearly constant, many ties with the maximum)", H.synth_sphere((72, 80, 64), np.uint16)))
    flat = np.zeros((20, 18, 512), np.uint16)
    flat[4:9, 3:12, 100:300] = 51234                         # uniform warps next to edges of a constant block
    out.append(("u16 constant block in empty space", flat))
    # the maximum sits on the last voxel of a row that ends inside the warp (idle lanes to its right)
    out.append(("u16 ramp, one octet per row", np.arange(64, dtype=np.uint16).reshape(1, 8, 8)))
    return out


@pytest.mark.parametrize("case", range(12))
def test_hist2d_four_voxels_per_thread_is_bit_exact(case, monkeypatch):
    """The vectorised 2-D histogram kernels (8 voxels per thread with fp32 squares where rows are whole octets, 4 voxels
    per thread with TVK_HIST2D_V4=1 / rows of whole quads) against the oracle AND the one-voxel-per-thread kernels
    (TVK_HIST2D_SCALAR=1)."""
    from imagevis3d_b200.context import TvkContext
    name, raw = _hist2d_volumes()[case]
    raw = np.ascontiguousarray(raw)
    assert raw.shape[2] % 4 == 0
    oc = O.build_octree(raw, 64, 2, False)
    bins_x = 256 if raw.dtype == np.uint8 else 4096
    maxv = float(oc["minmax"][:, 1].max())
    h_ref, mg_ref = O.hist2d(oc, bins_x, maxv)
    t = TvkContext(0)
    t.build_octree(raw, 64, 2)
    h_vec, mg_vec = t.build_hist2d(bins_x, maxv)
    monkeypatch.setenv("TVK_HIST2D_V4", "1")
    h_v4, mg_v4 = t.build_hist2d(bins_x, maxv)
    assert np.float32(mg_v4) == np.float32(mg_ref) and (h_v4 == h_ref).all(), name
    monkeypatch.setenv("TVK_HIST2D_SCALAR", "1")
    h_sc, mg_sc = t.build_hist2d(bins_x, maxv)
    t.close()
    assert h_vec.sum() == raw.size, name
    assert np.float32(mg_vec) == np.float32(mg_ref) == np.float32(mg_sc), name
    assert (h_sc == h_ref).all(), name
    bad = np.argwhere(h_vec != h_ref)
    assert bad.size == 0, (name, bad[:5], h_vec[tuple(bad[0])], h_ref[tuple(bad[0])])


def _streamed(tvk, raw, brick, overlap, median, splits, host_store=False):
    """Feeds `raw` through tvk_build_stream_* with the given brick-row split points; -> (tight store, sizes, minmax, hist1d)"""
    import torch
    dims = raw.shape[::-1]
    pinned = None
    if host_store:
        info = tvk.build_stream_begin(dims, raw.dtype.itemsize, brick, overlap, median)     # plan only, to learn the size
        nbytes = int(info.total_bytes)
        tvk.build_release()
        pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        info = tvk.build_stream_begin(dims, raw.dtype.itemsize, brick, overlap, median, store_ptr=pinned.data_ptr())
    else:
        info = tvk.build_stream_begin(dims, raw.dtype.itemsize, brick, overlap, median)
    nbz = int(info.lod_bricks[0][2])
    cuts = [0] + [c for c in splits if 0 < c < nbz] + [nbz]
    for b0, b1 in zip(cuts[:-1], cuts[1:]):
        z0, rows = tvk.build_stream_rows(b0, b1)
        slab = torch.from_numpy(np.ascontiguousarray(raw[z0:z0 + rows]).view(np.uint8).reshape(-1).copy()).cuda()
        tvk.build_stream_slab(slab.data_ptr(), z0, rows, b0, b1)
        del slab
    tvk.build_stream_end()
    off, sizes = tvk.build_toc()
    if host_store:
        tvk.synchronize()
        blob = pinned.numpy()
        bpv = raw.dtype.itemsize
        store = np.concatenate([blob[int(o):int(o) + int(s[0]) * int(s[1]) * int(s[2]) * bpv] for o, s in zip(off, sizes)])
    else:
        store, sizes = cuda_store_tight(tvk)
    mm, h1 = tvk.build_minmax(), tvk.build_hist1d()
    tvk.build_release()
    return store, sizes, mm, h1


@pytest.mark.parametrize("vol,dtype,brick,overlap,median,splits", [
    ((128, 96, 200), np.uint16, 24, 2, False, [1, 4, 5]),          # fused reduction + histogram per slab
    ((128, 96, 200), np.uint8, 24, 2, True, [3]),
    ((127, 61, 133), np.uint16, 20, 2, False, [2, 7]),              # unaligned rows: scalar reduction, plain histogram
    ((64, 64, 61), np.uint8, 64, 2, False, []),                      # one brick row: the whole volume as one slab
    ((130, 70, 129), np.uint16, 64, 2, False, [1, 2]),               # last brick row thinner than the overlap + odd depth
    ((96, 80, 97), np.uint8, 12, 4, False, [1, 2, 3, 5, 8, 13])])
def test_streamed_build_equals_one_shot(tvk, vol, dtype, brick, overlap, median, splits):
    """tvk_build_stream_*: LOD 0 fed in z-slabs gives the same store, min/max and histogram as tvk_build_octree, byte for byte."""
    raw = H.synth_noise(vol, dtype, seed=sum(vol), maxval=255 if dtype == np.uint8 else 65535)
    tvk.build_octree(raw, brick, overlap, median=bool(median))
    ref_store, ref_sizes = cuda_store_tight(tvk)
    ref_mm, ref_h1 = tvk.build_minmax(), tvk.build_hist1d()
    tvk.build_release()
    store, sizes, mm, h1 = _streamed(tvk, raw, brick, overlap, median, splits)
    assert (sizes == ref_sizes).all() and store.size == ref_store.size
    bad = np.nonzero(store != ref_store)[0]
    assert bad.size == 0, "first mismatching byte %d of %d" % (bad[0], store.size)
    assert (mm == ref_mm).all()
    assert h1.size == ref_h1.size and (h1 == ref_h1).all()


def test_streamed_build_into_pinned_host_store(tvk):
    """The brick store may be caller-owned pinned host memory: bricks are written over PCIe as they are cut (how a volume
    larger than HBM is bricked)."""
    raw = H.synth_noise((128, 96, 150), np.uint16, seed=11, maxval=65535)
    oc = O.build_octree(raw, 24, 2, False)
    store, sizes, mm, h1 = _streamed(tvk, raw, 24, 2, False, [2, 5], host_store=True)
    assert (sizes == oc["sizes"]).all() and (store == oc["store"]).all() and (mm == oc["minmax"]).all()
    assert (h1 == O.hist1d(oc)).all()


def test_streamed_build_rejects_gaps_and_short_slabs(tvk):
    import torch
    from imagevis3d_b200._lib import TvkError
    raw = H.synth_noise((64, 64, 100), np.uint8, seed=2, maxval=255)
    tvk.build_stream_begin((64, 64, 100), 1, 24, 2)
    z0, rows = tvk.build_stream_rows(0, 2)
    slab = torch.from_numpy(raw[z0:z0 + rows].reshape(-1).copy()).cuda()
    with pytest.raises(TvkError):
        tvk.build_stream_slab(slab.data_ptr(), z0, rows, 1, 2)          # does not start at brick row 0
    with pytest.raises(TvkError):
        tvk.build_stream_slab(slab.data_ptr(), z0, rows - 1, 0, 2)      # a row short
    tvk.build_stream_slab(slab.data_ptr(), z0, rows, 0, 2)
    with pytest.raises(TvkError):
        tvk.build_stream_end()                                          # rows 2.. missing
    tvk.build_release()
