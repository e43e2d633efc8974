#!/usr/bin/env python
"""BASELINE configs[2] measured (it is a parity-test case, not the bench line): synthetic 2048^3 uint16 volume
(17 GB raw, 49,597 pool bricks, fully resident), 1920x1080, orbiting camera, one converged frame per step:
1D-TF, 1D-TF lit, 2D-TF lit and isosurface, for the samplers that fit.  One JSON line per (mode, sampler) with
ms/frame, ray samples/s and the algorithmic HBM figure (1.2136*2/2 = 1.21 B/sample + 16 B/ray)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench as B
    from imagevis3d_b200 import _lib as L
    from imagevis3d_b200 import hostmath as HM
    from imagevis3d_b200.context import TvkContext
    from imagevis3d_b200.scene import load_resident
    from tests.test_gpu_fullsize import _volume_u16, N
    dev = torch.device("cuda", 0)
    raw = _volume_u16(dev)
    torch.cuda.synchronize()
    rmax = max(int(raw[z0:z0 + 64].view(torch.int16).to(torch.int32).bitwise_and(0xFFFF).max()) for z0 in range(0, N, 64))
    torch.cuda.empty_cache()                 # generator temporaries: the quad-packed pool (106 GB) needs the room
    peak, _ = B.peaks()
    steps = 8
    samplers = [("gather", L.SAMPLER_GATHER), ("exact", L.SAMPLER_EXACT), ("q4", L.SAMPLER_EXACT_Q4), ("tex3d", L.SAMPLER_TEX3D)]
    only = [a.split("=")[1] for a in sys.argv if a.startswith("--sampler=")]
    for sname, sampler in samplers:
        if only and sname not in only:
            continue
        tvk = TvkContext(0)
        sys.stderr.write("%s: free HBM before load %.1f GB\n" % (sname, torch.cuda.mem_get_info()[0] / 2 ** 30))
        try:
            sc = load_resident(tvk, raw.data_ptr(), (N, N, N), 2, sampler=sampler)
        except Exception as e:          # noqa: BLE001 -- e.g. the 4x quad-packed pool (104 GB) next to raw + store
            print(json.dumps({"bench": "config3", "sampler": sname, "skipped": str(e)[:120]}), flush=True)
            tvk.close()
            continue
        tvk.build_release()              # the pool is filled; the store (24 GB) is not needed for rendering
        tf1, lo, hi = B.bench_tf()
        tf1 = np.ascontiguousarray(np.repeat(tf1, 16, axis=0))          # 4096-entry table like a 16-bit dataset's
        v = np.linspace(0, 1, 4096, dtype=np.float32)
        g = np.linspace(0, 1, 256, dtype=np.float32)
        tf2 = np.zeros((256, 4096, 4), np.float32)
        tf2[..., 0] = v[None, :]; tf2[..., 1] = 0.3 + 0.7 * g[:, None]; tf2[..., 2] = 0.4     # noqa: E702
        tf2[..., 3] = np.clip((v[None, :] - 0.25) * 0.2, 0, 1) * (0.5 + 0.5 * g[:, None])
        tf2 = (np.clip(tf2, 0, 1) * np.float32(255)).astype(np.uint8)
        modes = [a.split("=")[1] for a in sys.argv if a.startswith("--mode=")] or ["1d", "1d-lit", "2d-lit", "iso"]
        for mode in modes:
            tvk.set_metadata(sc["slots"] + 3)
            if mode in ("1d", "1d-lit"):
                tvk.set_tf1d(tf1)
                tvk.recompute_visibility(0, lo * rmax / 255.0, hi * rmax / 255.0)
                kw = dict(render_mode=0, lighting=(mode == "1d-lit"), trans_scale=65535.0 / rmax)
            elif mode == "2d-lit":
                tvk.set_tf2d(tf2)
                tvk.recompute_visibility(1, 0.25 * rmax, float(rmax), -1e308, 1e308)
                kw = dict(render_mode=1, lighting=True, trans_scale=65535.0 / rmax, gradient_scale=40.0)
            else:
                tvk.recompute_visibility(2, rmax / 2.0)
                kw = dict(render_mode=2, isoval=(rmax / 2.0) / 65536.0)
            ms, samples, rays, nbricks = 0.0, 0, 0, 0
            for k in range(-2, steps):
                cam = B.camera_for_step(HM, max(k, 0), steps)
                p = HM.raycast_params(cam, 1920, 1080, (N, N, N), **kw)
                tvk.stats(reset=True)
                tvk.raycast(p)
                st = tvk.stats()
                if k >= 0:
                    ms += st.last_raycast_ms; samples += int(st.samples); rays += int(st.rays); nbricks += int(st.brick_steps)     # noqa: E702
            assert len(tvk.read_requests()) == 0
            alg = samples * 1.2136 + rays * 16.0
            print(json.dumps({"bench": "config3", "volume": [N, N, N], "dtype": "uint16", "resolution": [1920, 1080], "mode": mode,
                              "sampler": sname, "ms_per_frame": ms / steps, "samples_per_frame": samples / steps, "rays_per_frame": rays / steps,
                              "brick_visits_per_frame": nbricks / steps,
                              "samples_per_s": samples / (ms * 1e-3), "algorithmic_GBps": alg / (ms * 1e-3) / 1e9,
                              "hbm_roofline_frac": alg / (ms * 1e-3) / 1e9 / peak}), flush=True)
        tvk.close()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
