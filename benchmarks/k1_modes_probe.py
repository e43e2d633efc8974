#!/usr/bin/env python
"""K1 in its other render modes on the bench scene (512^3 uint8, 1920x1080, Q4 sampler): 1D-TF lit, 2D-TF lit,
isosurface.  Small enough to sit under `ncu --set full`; prints one JSON line per mode."""
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
    dev = torch.device("cuda", 0)
    N = 512
    raw = B.volume_torch((N, N, N), dev)
    rmax = int(raw.max())
    tvk = TvkContext(0)
    sc = load_resident(tvk, raw.data_ptr(), (N, N, N), 1, sampler=L.SAMPLER_EXACT_Q4)
    tf1, lo, hi = B.bench_tf()
    v = np.linspace(0, 1, 256, dtype=np.float32)
    g = np.linspace(0, 1, 256, dtype=np.float32)
    tf2 = np.zeros((256, 256, 4), np.float32)
    tf2[..., 0] = v[None, :]; tf2[..., 1] = 0.3 + 0.7 * g[:, None]; tf2[..., 2] = 0.4     # noqa: E702
    tf2[..., 3] = np.clip((v[None, :] - 0.25) * 0.2, 0, 1) * (0.5 + 0.5 * g[:, None])
    tf2 = (np.clip(tf2, 0, 1) * np.float32(255)).astype(np.uint8)
    steps = 4
    for mode in ("1d-lit", "2d-lit", "iso"):
        tvk.set_metadata(sc["slots"] + 3)
        if mode == "1d-lit":
            tvk.set_tf1d(tf1)
            tvk.recompute_visibility(0, lo, hi)
            kw = dict(render_mode=0, lighting=True, trans_scale=255.0 / rmax)
        elif mode == "2d-lit":
            tvk.set_tf2d(tf2)
            tvk.recompute_visibility(1, 0.25 * rmax, float(rmax), -1e308, 1e308)
            kw = dict(render_mode=1, lighting=True, trans_scale=255.0 / rmax, gradient_scale=40.0)
        else:
            tvk.recompute_visibility(2, rmax / 2.0)
            kw = dict(render_mode=2, isoval=(rmax / 2.0) / 256.0)
        ms, samples = 0.0, 0
        for k in range(-1, steps):
            p = HM.raycast_params(B.camera_for_step(HM, max(k, 0), steps), 1920, 1080, (N, N, N), **kw)
            tvk.stats(reset=True)
            tvk.raycast(p)
            st = tvk.stats()
            if k >= 0:
                ms += st.last_raycast_ms; samples += int(st.samples)     # noqa: E702
        print(json.dumps({"probe": "k1_modes", "volume": [N, N, N], "mode": mode, "ms_per_frame": ms / steps,
                          "samples_per_frame": samples / steps, "samples_per_s": samples / (ms * 1e-3)}), flush=True)
    tvk.close()


if __name__ == "__main__":
    main()
