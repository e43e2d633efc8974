"""CPU: pool sizing -- BrickPool::ComputePoolCapacity through the C ABI (no device needed) against the
reference's own statements (GPUMemMan::GetVolumePool, cut out of GPUMemMan.cpp and run by oracle/ref_extract.py) and, where
/root/reference is absent, against cases worked out by hand from those statements."""
import ctypes as C
import itertools

import pytest

from imagevis3d_b200 import _lib as L
from oracle import ref_extract as R


def _capacity(max_brick, bpv, total, gpu_mem, max3d):
    cap = (C.c_uint32 * 3)()
    fn = L.load().tvkr_compute_pool_capacity
    fn.restype = C.c_int
    rc = fn((C.c_uint32 * 3)(*max_brick), C.c_uint32(bpv), C.c_uint64(total), C.c_uint64(gpu_mem), C.c_uint32(max3d), cap)
    assert rc == 0
    return tuple(int(v) for v in cap)


# (max brick, bytes/voxel, bricks in the dataset, pool bytes, 3-D texture limit) -> slots per axis, by hand:
HAND = [
    # 49,597 bricks of 64^3 u16 (configs[2]) with 1 GiB: memory bound.  voxels = 2^29, cbrt = 812, round(812/64)=13 -> 832;
    # y: (2^29 / 832^2) / 64 = 12 -> 768; z: (2^29 / (832*768)) / 64 = 13 -> 832
    (((64, 64, 64), 2, 49597, 1 << 30, 16384), (13, 12, 13)),
    # 890 bricks of 64^3 u8 (configs[1]) with plenty of memory: dataset bound.  cbrt(890) = 9 -> 9 x ceil(890/81) = 11 x ceil(890/99) = 9
    (((64, 64, 64), 1, 890, 64 << 30, 16384), (9, 11, 9)),
    # a single brick
    (((64, 64, 64), 1, 1, 64 << 30, 16384), (1, 1, 1)),
]


@pytest.mark.parametrize("args,want", HAND)
def test_pool_capacity_hand_cases(args, want):
    assert _capacity(*args) == want


@pytest.mark.skipif(not R.available(), reason="/root/reference is not mounted")
def test_pool_capacity_equals_reference_statements():
    cases = 0
    for mb, bpv, total, mem, max3d in itertools.product(
            [(64, 64, 64), (20, 20, 20), (64, 48, 32), (256, 256, 256), (36, 64, 20)], [1, 2],
            [1, 7, 143, 890, 4681, 49597, 378106, 1487165], [64 << 20, 512 << 20, 1 << 30, 3 << 30, 24 << 30, 150 << 30],
            [2048, 16384]):
        px, py, pz = R.pool_size(8 * bpv, mb, total, max3d, mem)
        if min(px, py, pz) == 0:
            continue          # less memory than one brick per axis: the reference builds an empty texture and fails later
        assert _capacity(mb, bpv, total, mem, max3d) == (px // mb[0], py // mb[1], pz // mb[2]), (mb, bpv, total, mem, max3d)
        cases += 1
    assert cases > 500
