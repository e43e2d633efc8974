#!/usr/bin/env python
"""Pager (K3) micro-benchmark: bricks/s and host->device GB/s of tvk_pool_upload_bricks streaming
64^3 bricks from pinned host DRAM into pool slots (DMA on the side stream + placement kernel),
in batches of 509 requests = one sub-frame of the reference (GLHashTable size, MasterController.h:61-75).
Reference counterpart: GLVolumePool::UploadBricks -> one synchronous glTexSubImage3D per brick from
pageable memory (GLVolumePool.cpp:673-717,1362-1393).  Prints one JSON line per configuration."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from imagevis3d_b200 import _lib as L
    from imagevis3d_b200.context import TvkContext
    tvk = TvkContext(0)
    for bpv, sampler, name in [(1, L.SAMPLER_EXACT, "u8/exact"), (2, L.SAMPLER_EXACT, "u16/exact"),
                               (2, L.SAMPLER_EXACT_Q4, "u16/q4"), (2, L.SAMPLER_TEX3D, "u16/tex3d")]:
        cap = (16, 16, 8)                                     # 2048 slots
        nslots = cap[0] * cap[1] * cap[2]
        tvk.pool_create((4096, 4096, 4096), (64, 64, 64), 2, bpv, cap, 8, sampler)
        nbricks = 4096
        brick_bytes = 64 ** 3 * bpv
        host = torch.empty(nbricks * brick_bytes, dtype=torch.uint8, pin_memory=True)
        host.random_(0, 255)
        rng = np.random.default_rng(1)
        batch = 509
        sizes = np.full((batch, 3), 64, np.uint32).reshape(-1)
        total_bricks, total_s = 0, 0.0
        for it in range(12):
            ids = rng.integers(0, nbricks, batch)             # scattered source bricks (request order)
            slots = rng.permutation(nslots)[:batch].astype(np.uint32)
            src = (C.c_void_p * batch)(*[host.data_ptr() + int(i) * brick_bytes for i in ids])
            tvk.synchronize()
            t0 = time.perf_counter()
            tvk._ck(tvk.lib.tvk_pool_upload_bricks(tvk.h, C.c_uint32(batch), slots.ctypes.data_as(C.c_void_p), src,
                                                   sizes.ctypes.data_as(C.c_void_p), C.c_int(0)), "upload")
            tvk._ck(tvk.lib.tvk_pool_upload_wait(tvk.h), "wait")
            dt = time.perf_counter() - t0
            if it >= 2:                                       # 2 warm-up batches
                total_bricks += batch
                total_s += dt
        print(json.dumps({"bench": "pager", "config": name, "brick": 64, "batch": batch,
                          "bricks_per_s": total_bricks / total_s,
                          "h2d_GBps": total_bricks * brick_bytes / total_s / 1e9,
                          "ms_per_batch": 1e3 * total_s / (total_bricks / batch)}))
    tvk.close()


if __name__ == "__main__":
    main()
