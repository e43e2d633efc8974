#!/usr/bin/env python
"""Quick K1 timing on the bench's configs[2] scene (2048^3 uint16 sphere, 1920x1080): ms/frame of the four render
modes, 6 frames each after 2 warm-ups.  Used with TVK_CUDA_LIB=build_variants/libtvk_X.so to compare kernel variants
(benchmarks/k1_sweep.py build ...).  One JSON line per mode."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench as B
    from imagevis3d_b200 import _lib as L
    from imagevis3d_b200 import hostmath as HM
    from imagevis3d_b200.context import TvkContext
    dev = torch.device("cuda", 0)
    sampler = {"exact": L.SAMPLER_EXACT, "tex3d": L.SAMPLER_TEX3D, "q4": L.SAMPLER_EXACT_Q4, "gather": L.SAMPLER_GATHER}[os.environ.get("TVK_SAMPLER", "gather")]
    tvk = TvkContext(0)
    raw = B.sphere_torch(B.C3_VOL, (0, 0, 0), B.C3_VOL, 16, dev)
    scene = B.Scene(tvk, raw, B.C3_VOL, 16, want_host_copy=False)
    del raw
    torch.cuda.empty_cache()
    scene.fill_pool(sampler)
    tvk.build_release()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    steps = 6
    modes = [a for a in sys.argv[1:] if a in B.MODES] or ["2d_lit", "1d", "1d_lit", "iso"]
    out = {}
    for mode in modes:
        ms = B.mode_setup(mode, scene.rmax, 16, B.gradient_scale(B.C3_VOL))
        scene.set_mode(ms)
        params = [HM.raycast_params(B.camera_for_step(HM, k, steps), 1920, 1080, B.C3_VOL, **ms["kw"]) for k in range(steps)]
        rec = B.timed_frames(tvk, torch, flush, params, steps, 2, 0)
        out[mode] = rec["ms_per_frame"]
        print(json.dumps({"lib": os.path.basename(os.environ.get("TVK_CUDA_LIB", "libtvk_cuda.so")), "mode": mode, "ms_per_frame": rec["ms_per_frame"],
                          "gsamples_per_s": rec["samples_per_s"] / 1e9, "samples_per_frame": rec["samples"] / steps}), flush=True)
    tvk.close()


if __name__ == "__main__":
    main()
