"""CPU: the coarse-to-fine sub-frame ladder (row R13) -- ComputeMaxLODForCurrentView of the product's host, through the C ABI,
against THE REFERENCE'S FUNCTION BODY (AbstrRenderer.cpp:681-787, cut out and run by oracle/ref_extract.py) on random
state sequences, and scripted cases that need no reference."""
import ctypes as C

import numpy as np
import pytest

from imagevis3d_b200 import _lib as L
from oracle import ref_extract as R


class LodState(C.Structure):
    _fields_ = [("target", C.c_int32), ("msec_passed", C.c_float * 2), ("max_ms_per_frame", C.c_float), ("lod_not_ok_counter", C.c_uint32),
                ("perf_lod_skip", C.c_uint64), ("low_res_intermediate", C.c_int32), ("decrease_sampling_rate", C.c_int32),
                ("decrease_screen_res", C.c_int32), ("brick_list_size", C.c_uint64), ("bricks_rendered", C.c_uint64),
                ("max_lod_index", C.c_uint64), ("min_lod_for_view", C.c_uint64), ("lod_limit_x", C.c_uint64),
                ("start_lod_offset", C.c_uint64), ("current_lod_offset", C.c_uint64)]


def _run(st):
    s = LodState()
    (s.target, s.msec_passed[0], s.msec_passed[1], s.max_ms_per_frame, s.lod_not_ok_counter, s.perf_lod_skip, s.low_res_intermediate,
     s.decrease_sampling_rate, s.decrease_screen_res, s.brick_list_size, s.bricks_rendered, s.max_lod_index, s.min_lod_for_view,
     s.lod_limit_x) = st
    fn = L.load().tvkr_compute_max_lod_for_view
    fn.restype = C.c_int
    assert fn(C.byref(s)) == 0
    return (float(s.msec_passed[0]), float(s.msec_passed[1]), int(s.lod_not_ok_counter), int(s.perf_lod_skip), int(s.decrease_sampling_rate),
            int(s.decrease_screen_res), int(s.start_lod_offset), int(s.current_lod_offset))


def test_ladder_scripted():
    # very first interactive frame (no timings yet): start at the coarsest LOD
    assert _run((0, -1.0, -1.0, 100.0, 0, 0, 0, 0, 0, 10, 10, 5, 1, 0))[6:] == (5, 5)
    # headless / capture: straight to the LOD the view needs
    assert _run((2, -1.0, -1.0, 100.0, 0, 0, 0, 0, 0, 10, 10, 5, 1, 0))[6:] == (1, 1)
    assert _run((1, 50.0, 50.0, 100.0, 0, 3, 0, 0, 0, 10, 10, 5, 1, 0))[6:] == (1, 1)
    # fast first and second sub-frames: the skip grows by one, the start LOD gets one finer
    r = _run((0, 20.0, 30.0, 100.0, 2, 1, 0, 0, 0, 10, 10, 5, 1, 0))
    assert r[2:4] == (0, 2) and r[6:] == (3, 3) and r[:2] == (-1.0, -1.0)
    # too slow: three frames of grace, then one LOD coarser
    assert _run((0, 500.0, 30.0, 100.0, 1, 2, 0, 0, 0, 10, 10, 5, 1, 0))[2:4] == (2, 2)
    r = _run((0, 500.0, 30.0, 100.0, 3, 2, 0, 0, 0, 10, 10, 5, 1, 0))
    assert r[2:4] == (0, 1) and r[6:] == (4, 4)
    # already coarsest and "use all means": resolution first, then sampling rate
    assert _run((0, 500.0, 30.0, 100.0, 3, 0, 1, 0, 0, 10, 10, 5, 1, 0))[4:6] == (0, 1)
    assert _run((0, 500.0, 30.0, 100.0, 3, 0, 1, 0, 1, 10, 10, 5, 1, 0))[4:6] == (1, 1)
    # the LOD limit caps how coarse a view may start
    assert _run((0, -1.0, -1.0, 100.0, 0, 0, 0, 0, 0, 10, 10, 5, 1, 2))[6:] == (3, 3)


@pytest.mark.skipif(not R.available(), reason="/root/reference is not mounted")
def test_ladder_equals_reference_body():
    rng = np.random.default_rng(11)
    states = []
    for _ in range(4000):
        maxidx = int(rng.integers(0, 9))
        minlod = int(rng.integers(0, maxidx + 1))
        lst = int(rng.integers(0, 50))
        states.append((int(rng.integers(0, 3)), float(rng.choice([-1.0, 5.0, 99.0, 100.0, 101.0, 4000.0])),
                       float(rng.choice([-1.0, 5.0, 100.0, 250.0])), 100.0, int(rng.integers(0, 5)), int(rng.integers(0, maxidx + 2)),
                       int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.integers(0, 2)), lst,
                       lst if rng.random() < 0.7 else int(rng.integers(0, lst + 1)), maxidx, minlod, int(rng.integers(0, maxidx + 1))))
    want = R.max_lod_for_view(states)
    assert len(want) == len(states)
    for st, w in zip(states, want):
        assert _run(st) == w, st
