"""CPU, world_size 2 (gloo): the sort-last exchange (stripe partition, depth order, all_to_all,
ordered blend, gather) against a direct ordered composite computed by the oracle."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imagevis3d_b200 import sortlast as SL
from oracle import oracle as O


def _oracle_composite(parts):
    parts = [np.ascontiguousarray(p, np.float32) for p in parts]
    pix = parts[0].shape[0] * parts[0].shape[1]
    out = np.zeros_like(parts[0])
    arr = (C.c_void_p * len(parts))(*[p.ctypes.data for p in parts])
    O.lib().orc_composite_ordered(arr, C.c_uint32(len(parts)), C.c_uint64(pix), out.ctypes.data_as(C.c_void_p))
    return out


def _partial(rank, h, w):
    rng = np.random.default_rng(100 + rank)
    a = rng.uniform(0, 1, (h, w, 1)).astype(np.float32)
    return np.concatenate([rng.uniform(0, 1, (h, w, 3)).astype(np.float32) * a, a], axis=2)


def _worker(rank, world, port, h, w, order, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = torch.from_numpy(_partial(rank, h, w))
    full = SL.exchange_and_composite(part, order, lambda ps: torch.from_numpy(_oracle_composite([p.numpy() for p in ps])),
                                     world, rank)
    np.save(os.path.join(out_dir, "full_%d.npy" % rank), full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("h,w,order", [(48, 64, [0, 1]), (37, 50, [1, 0])])
def test_direct_send_compositing_world2(tmp_path, h, w, order):
    world, port = 2, 29500 + (os.getpid() + h) % 2000
    mp.spawn(_worker, args=(world, port, h, w, order, str(tmp_path)), nprocs=world, join=True)
    want = _oracle_composite([_partial(r, h, w) for r in order])
    for r in range(world):
        got = np.load(str(tmp_path / ("full_%d.npy" % r)))
        assert got.shape == (h, w, 4)
        assert np.array_equal(got, want)          # same fp32 operations in the same order


def test_slab_partition_and_depth_order():
    parts = SL.slab_partition((512, 512, 2048), 4, halo=64)
    assert [(p["z0"], p["z1"]) for p in parts] == [(0, 512), (512, 1024), (1024, 1536), (1536, 2048)]
    assert parts[0]["zs"] == 0 and parts[0]["ze"] == 576 and parts[1]["zs"] == 448 and parts[3]["ze"] == 2048
    assert parts[1]["local_dims"] == (512, 512, 640)
    eye = np.eye(4)
    eye[3, :3] = [0.5, 0.5, -1.0]                 # eye in front of slab 0
    assert SL.depth_order(eye, parts) == [0, 1, 2, 3]
    eye[3, 2] = 2.0                               # behind the last slab
    assert SL.depth_order(eye, parts) == [3, 2, 1, 0]
    eye[3, 2] = 0.6                               # inside slab 2: it comes first, the rest by distance
    o = SL.depth_order(eye, parts)
    assert o[0] == 2 and o.index(1) < o.index(0) and set(o) == {0, 1, 2, 3}
    assert SL.stripe_rows(1080, 8) == 135 and SL.stripe_rows(37, 2) == 19
