#!/usr/bin/env python
"""Multi-GPU check of the fused sort-last exchange (run under torchrun, one rank per GPU):
every rank renders its z-slab; the image assembled by the peer-memory path (K1 peer stores + arrival flags +
stripe compositor, csrc/tvk_exchange.cu) must equal BIT FOR BIT the image of the NCCL direct-send path
(all_to_all + tvk_composite_ordered + all_gather) for the same partial images, over several frames.
Rank 0 prints one JSON line.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/sortlast_check.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench as B
    from imagevis3d_b200 import hostmath as HM
    from imagevis3d_b200 import sortlast as SL
    from imagevis3d_b200.context import TvkContext
    from tests.util import resident_pool

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    tvk = TvkContext(local)
    work_stream = torch.cuda.Stream(device=dev)      # (the legacy default stream's handle 0 would mean "own stream")
    torch.cuda.set_stream(work_stream)
    tvk.set_stream(work_stream.cuda_stream)
    G = (256, 256, 128 * world) if os.environ.get("TVK_PARTITION", "blocks") == "slabs" else (256 * (2 if world >= 2 else 1), 256 * (2 if world >= 4 else 1), 256 * max(world // 4, 1))
    W, Hh = 1283, 717                              # neither divisible by the world size nor by the tile size
    parts = SL.slab_partition(G, world, halo=64) if os.environ.get("TVK_PARTITION", "blocks") == "slabs" else SL.block_partition(G, world, halo=64)
    part = parts[rank]
    raw = B.volume_torch(part["local_dims"], dev, origin=part["s"], total=G).cpu().numpy()
    tf, lo, hi = B.bench_tf()
    resident_pool(tvk, raw, 64, tf, lo, hi, sampler=2)
    fused = SL.FusedExchange(tvk, world, rank, W, Hh, dist)

    class _CAI:
        def __init__(self, ptr, shape):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (int(ptr), False), "version": 2}

    def gpu_composite(stripes):
        out = torch.empty_like(stripes[0])
        tvk.composite_ordered([t.data_ptr() for t in stripes], stripes[0].shape[0] * stripes[0].shape[1], out.data_ptr())
        return out

    mismatches, maxv = 0, 0.0
    for k in range(5):
        cam = B.camera_for_step(HM, k, 5)
        p = HM.raycast_params(cam, W, Hh, G, slab=part)
        pg = HM.raycast_params(cam, W, Hh, G)
        order = SL.depth_order(np.array(pg.eye_to_model[:]).reshape(4, 4), parts)
        ptr = fused.render(p, order)
        tvk.synchronize()
        got = torch.as_tensor(_CAI(ptr, (Hh, W, 4)), device=dev).clone()
        partial = torch.as_tensor(_CAI(tvk.image_device_ptr(), (Hh, W, 4)), device=dev)
        want = SL.exchange_and_composite(partial, order, gpu_composite, world, rank)
        torch.cuda.synchronize()
        if rank == 0:
            mismatches += int((got != want).any(dim=-1).sum().item())
            maxv = max(maxv, float(want.max().item()))
    timed_out = tvk.exchange_timed_out()
    t = torch.tensor([1.0 if timed_out else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"check": "fused sort-last exchange == NCCL direct send", "world": world, "frames": 5,
                          "resolution": [W, Hh], "mismatching_pixels": mismatches, "image_max": maxv,
                          "timed_out": bool(t.item()), "ok": mismatches == 0 and maxv > 0.05 and not t.item()}))
    dist.barrier()
    tvk.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
