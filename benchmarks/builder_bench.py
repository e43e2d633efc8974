#!/usr/bin/env python
"""Builder (K2) benchmark: voxels/s and effective GB/s (algorithmic bytes 2.387*bpv per voxel)
of tvk_build_octree on volumes resident in HBM, with the reference's own converter timed beside it when
oracle/_ref/ref_driver exists (build container only) or the oracle C port otherwise (smaller volume)."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from imagevis3d_b200.context import TvkContext
    sys.path.insert(0, ROOT)
    import bench as B
    tvk = TvkContext(0)
    dev = torch.device("cuda", 0)
    for dims, bpv in [((512, 512, 512), 1), ((1024, 1024, 1024), 1), ((1024, 1024, 1024), 2)]:
        raw = B.volume_torch(dims, dev)
        if bpv == 2:
            raw = (raw.to(torch.int32) * 257).to(torch.uint16)
        torch.cuda.synchronize()
        best = 1e9
        for it in range(4):
            t0 = time.perf_counter()
            tvk.build_octree(raw_dptr=raw.data_ptr(), volume_size=dims, bpv=bpv)
            tvk.synchronize()
            best = min(best, time.perf_counter() - t0)
            tvk.build_release()
        n = dims[0] * dims[1] * dims[2]
        print(json.dumps({"bench": "builder", "volume": list(dims), "bpv": bpv, "ms": 1e3 * best,
                          "voxels_per_s": n / best, "algorithmic_GBps": 2.387 * bpv * n / best / 1e9}))
    # CPU side: the oracle port on 256^3 (single thread, like the reference converter)
    from oracle import oracle as O
    vol = B.volume_numpy((256, 256, 256))
    t0 = time.perf_counter()
    O.build_octree(vol, 64, 2)
    dt = time.perf_counter() - t0
    print(json.dumps({"bench": "builder_cpu_port", "volume": [256, 256, 256], "bpv": 1, "ms": 1e3 * dt,
                      "voxels_per_s": 256 ** 3 / dt, "cores": 1}))
    drv = O.ref_driver_path()
    if drv:
        with tempfile.TemporaryDirectory() as td:
            f = os.path.join(td, "v.raw")
            vol.tofile(f)
            t0 = time.perf_counter()
            subprocess.check_call([drv, "convert", f, "256", "256", "256", "u8", "64", "2", "0", "0", os.path.join(td, "o.bin"), "0"],
                                  stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
        print(json.dumps({"bench": "builder_reference", "volume": [256, 256, 256], "bpv": 1, "ms": 1e3 * dt,
                          "voxels_per_s": 256 ** 3 / dt, "cores": 1, "note": "unmodified ExtendedOctreeConverter, file-backed (tmpfs)"}))
    tvk.close()


if __name__ == "__main__":
    main()
