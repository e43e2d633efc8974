"""GPU: brick indexing, LOD selection, request serialisation and visibility flags -- bit-exact."""
import ctypes as C

import numpy as np
import pytest

from oracle import host as H
from oracle import oracle as O
from tests.util import mirror_pool

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_pool():
    vol = H.synth_blobs((150, 131, 97), np.uint8)
    oc = O.build_octree(vol, 20, 2)
    pool = H.OraclePool(oc, (5, 5, 5))
    pool.upload_first_brick()
    return pool


def test_lod_selection_bit_exact(tvk, small_pool):
    mirror_pool(tvk, small_pool)
    lf = float(H.lod_factor(50.0, 1080))
    wse = float(np.float32(1.0 / 150.0))
    rng = np.random.default_rng(5)
    depth = -np.exp(rng.uniform(np.log(1e-3), np.log(50.0), 200000)).astype(np.float32)
    # exact power-of-two boundaries of lf*(-d)/wse and their float neighbours, plus junk
    k = np.arange(-2, 8)
    edge = (-(2.0 ** k) * wse / lf).astype(np.float32)
    edges = np.concatenate([np.nextafter(edge, np.float32(-np.inf)), edge, np.nextafter(edge, np.float32(0))])
    depth = np.concatenate([depth, edges, np.array([0.0, 1.0, 501.0, np.nan, -np.inf], np.float32)])
    want = np.zeros(depth.size, np.uint32)
    O.lib().orc_probe_lod(depth.ctypes.data_as(C.c_void_p), C.c_uint32(depth.size), C.c_float(lf), C.c_float(wse),
                          C.c_uint32(small_pool.lod_count - 1), want.ctypes.data_as(C.c_void_p))
    got = tvk.probe_lod(depth, lf, wse)
    assert (got == want).all()
    assert got.max() == small_pool.lod_count - 1 and got.min() == 0


def test_brick_index_and_serialisation_bit_exact(tvk, small_pool):
    mirror_pool(tvk, small_pool)
    rng = np.random.default_rng(11)
    n = 300000
    xyz = rng.uniform(-0.05, 1.05, (n, 3)).astype(np.float32)
    xyz[:1000] = rng.integers(0, 9, (1000, 3)).astype(np.float32) / np.float32(8)     # brick faces
    xyz[1000:1100] = 1.0
    lod = rng.integers(0, small_pool.lod_count, n).astype(np.uint32)
    want_i = np.zeros(n, np.uint32)
    want_s = np.zeros(n, np.uint32)
    O.lib().orc_probe_brick_index((C.c_uint32 * 3)(*small_pool.vol), (C.c_uint32 * 3)(*small_pool.max_brick),
                                  C.c_uint32(small_pool.overlap), C.c_uint32(small_pool.lod_count),
                                  xyz.ctypes.data_as(C.c_void_p), lod.ctypes.data_as(C.c_void_p), C.c_uint32(n),
                                  want_i.ctypes.data_as(C.c_void_p), want_s.ctypes.data_as(C.c_void_p))
    got_i, got_s = tvk.probe_brick_index(xyz, lod)
    assert (got_i == want_i).all() and (got_s == want_s).all()
    assert got_i.max() < small_pool.total


@pytest.mark.parametrize("mode,lo,hi", [(0, 40.0, 90.0), (0, 0.0, 255.0), (0, 200.0, 255.0), (0, 300.0, 400.0),
                                       (2, 128.0, 0.0), (2, 1.0, 0.0), (1, 10.0, 20.0)])
def test_visibility_flags_bit_exact(tvk, small_pool, mode, lo, hi):
    pool = H.OraclePool(small_pool.oct, small_pool.capacity)
    pool.upload_first_brick()
    # make a handful of bricks resident so that the resident / EMPTY transitions are covered
    pool.upload_requests([(0, 0, 0, 0), (3, 3, 2, 0), (1, 1, 1, 1), (0, 0, 0, 2), (4, 2, 1, 0)])
    mirror_pool(tvk, pool)
    tvk.set_minmax(pool.minmax)
    want_counts = pool.recompute_visibility(mode, lo, hi, -1e308, 1e308)
    got_counts = tvk.recompute_visibility(mode, lo, hi, -1e308, 1e308)
    assert (tvk.get_metadata() == pool.meta).all()
    assert list(got_counts) == list(want_counts)
    kinds = set(np.unique(pool.meta[pool.meta < 3]).tolist())
    assert kinds  # at least one non-resident class present
