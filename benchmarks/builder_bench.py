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
    cases = [((512, 512, 512), 1), ((1024, 1024, 1024), 1), ((1024, 1024, 1024), 2)]
    if "--big" in sys.argv:
        cases.append(((2048, 2048, 2048), 2))
    only = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--case=")]
    if only:
        cases = [cases[i] for i in only]
    for dims, bpv in cases:
        raw = torch.empty((dims[2], dims[1], dims[0]), dtype=torch.uint8 if bpv == 1 else torch.uint16, device=dev)
        for z0 in range(0, dims[2], 64):      # slab-wise: the fp32 temporaries of a 2048^3 volume would not fit
            zn = min(64, dims[2] - z0)
            s = B.volume_torch(dims, dev, z0=z0, zn=zn, ztotal=dims[2])
            raw[z0:z0 + zn] = s if bpv == 1 else (s.to(torch.int32) * 257).to(torch.uint16)
            del s
        torch.cuda.synchronize()
        best = 1e9
        kbest = 1e9
        for it in range(1 if "--once" in sys.argv else 4):
            t0 = time.perf_counter()
            tvk.build_octree(raw_dptr=raw.data_ptr(), volume_size=dims, bpv=bpv)
            tvk.synchronize()
            best = min(best, time.perf_counter() - t0)
            kbest = min(kbest, tvk.build_last_ms() * 1e-3)
            if it == 0:
                h1 = tvk.build_hist1d()
                tvk.synchronize()
                t2 = time.perf_counter()
                tvk.build_hist2d(min(len(h1), 4096), float(tvk.build_minmax()[:, 1].max()))
                h2_ms = 1e3 * (time.perf_counter() - t2)
                h2_kernel_ms = tvk.build_hist2d_last_ms()
            tvk.build_release()
        n = dims[0] * dims[1] * dims[2]
        print(json.dumps({"bench": "builder", "volume": list(dims), "bpv": bpv, "wall_ms_incl_alloc": 1e3 * best,
                          "kernels_ms": 1e3 * kbest, "voxels_per_s": n / kbest,
                          "algorithmic_GBps": 2.387 * bpv * n / kbest / 1e9, "hist2d_wall_ms": h2_ms, "hist2d_kernels_ms": h2_kernel_ms,
                          "note": "kernels_ms = CUDA events around pyramid + brick cut + min/max + 1-D histogram"}), flush=True)
        del raw
        torch.cuda.empty_cache()
    if only:
        return
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
