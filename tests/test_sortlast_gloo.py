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


def test_block_partition_and_visibility_order():
    """regular-grid boxes: cover the volume exactly, cuts at multiples of 64, and depth_order is a correct
    front-to-back order -- checked by brute force: along random rays from the eye the boxes met must appear in
    increasing position of the order."""
    assert SL.block_grid((1024, 1024, 1024), 8) == (2, 2, 2) and SL.block_grid((1024, 512, 512), 2) == (2, 1, 1)
    assert SL.block_grid((1024, 1024, 512), 4) == (2, 2, 1) and SL.block_grid((64, 64, 192), 3) == (1, 1, 3)
    # ties: y (the up axis views orbit about) is cut first, then x, then z
    assert SL.block_grid((512, 512, 512), 2) == (1, 2, 1) and SL.block_grid((512, 512, 512), 4) == (2, 2, 1)
    rng = np.random.default_rng(7)
    for dims, world in [((1024, 1024, 1024), 8), ((1024, 1024, 512), 4), ((768, 512, 640), 6), ((512, 512, 512), 12)]:
        parts = SL.block_partition(dims, world, halo=64)
        assert len(parts) == world
        vol = sum(int(np.prod(np.array(p["hi"]) - np.array(p["lo"]))) for p in parts)
        assert vol == int(np.prod(dims))
        for p in parts:
            assert all(v % 64 == 0 or v == d for v, d in zip(p["lo"] + p["hi"], dims + dims))
            assert all(p["s"][a] <= p["lo"][a] and p["hi"][a] <= p["e"][a] for a in range(3))
        g = np.array(dims, np.float64)
        for trial in range(40):
            eye = rng.uniform(-1.5, 2.5, 3)
            m = np.eye(4)
            m[3, :3] = eye
            order = SL.depth_order(m, parts)
            pos = {r: i for i, r in enumerate(order)}
            for ray in range(25):
                target = rng.uniform(0, 1, 3)
                d = target - eye
                hits = []
                for r, p in enumerate(parts):                       # slab test against the box in normalised coords
                    lo, hi = np.array(p["lo"]) / g, np.array(p["hi"]) / g
                    with np.errstate(divide="ignore", invalid="ignore"):
                        t0, t1 = (lo - eye) / d, (hi - eye) / d
                    tn, tf = np.minimum(t0, t1).max(), np.maximum(t0, t1).min()
                    if tf > max(tn, 0.0) + 1e-9:
                        hits.append((max(tn, 0.0), r))
                hits.sort()
                seq = [pos[r] for _, r in hits]
                assert seq == sorted(seq), (dims, world, eye, hits, order)
