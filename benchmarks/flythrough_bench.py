#!/usr/bin/env python
"""Out-of-core fly-through (BASELINE configs[4], scaled to one box's host RAM): the K2 builder bricks a
synthetic volume on the GPU, the brick store moves to PINNED host memory, the pool is capped well below the
dataset, and a camera flies a Bezier path through key points as LuaScripts/FlyThroughAnimator.lua does
(computeTangents / decasteljau, :198-330; uniform curve parameter instead of its arc-length table), each frame
painted until converged through the Lua-named renderer surface (setViewPos / setViewDir / setUpDir / paint /
converged).  The ray-caster's request table drives the pager (LRU, 509 bricks per sub-frame).
Prints one JSON line: cold pass (empty pool) and warm pass (same path again) ms/frame, sub-frames and paged
bricks per frame, host->device GB/s while paging.

  python benchmarks/flythrough_bench.py [--size 1024] [--frames 120] [--pool-mb 384] [--res 1920x1080]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def vlerp(a, b, u):
    return a + (b - a) * u


def decasteljau(p0, p1, p2, p3, u):
    p10, p11, p12 = vlerp(p0, p1, u), vlerp(p1, p2, u), vlerp(p2, p3, u)
    p20, p21 = vlerp(p10, p11, u), vlerp(p11, p12, u)
    return vlerp(p20, p21, u)


def tangents(p0, p1, p2, p3):
    alpha = np.linalg.norm(p2 - p1) / 2
    if alpha == 0.0:
        return p1, p2
    n = lambda v: v / max(np.linalg.norm(v), 1e-20)      # noqa: E731
    return p1 + n(p2 - p0) * alpha, p2 + n(p1 - p3) * alpha


def camera_path(keys, frames):
    """keys: list of dicts eye/ref/vup (numpy 3-vectors); the path runs from key 1 to key n-2 (the first and
    last key only shape the end tangents, as in the Lua script)."""
    segs = len(keys) - 3
    out = []
    for f in range(frames):
        u = 1 + segs * f / max(frames - 1, 1)
        i = min(int(np.floor(u)), len(keys) - 3)
        w = u - i
        cam = {}
        for k in ("eye", "ref", "vup"):
            p0, p1, p2, p3 = keys[i - 1][k], keys[i][k], keys[i + 1][k], keys[i + 2][k]
            q1, q2 = tangents(p0, p1, p2, p3)
            cam[k] = decasteljau(p1, q1, q2, p2, w)
        out.append(cam)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1024", help="N or XxYxZ")
    ap.add_argument("--bits", type=int, default=8)
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--pool-mb", type=int, default=384)
    ap.add_argument("--res", default="1920x1080")
    ap.add_argument("--streamed", action="store_true",
                    help="generate and brick the volume slab by slab straight into pinned host memory (volumes larger than HBM)")
    args = ap.parse_args()
    import torch
    import bench as B
    from imagevis3d_b200.renderer import Renderer, RT_HEADLESS
    w, h = [int(v) for v in args.res.split("x")]
    dims = [int(v) for v in args.size.split("x")]
    dims = dims * 3 if len(dims) == 1 else dims
    bits = args.bits
    dev = torch.device("cuda", 0)
    r = Renderer(0)
    r.setState(hash_table_size=4093, rehash_count=16)
    t0 = time.perf_counter()
    build_ms = None
    if args.streamed:
        build_ms = r.loadStreamed(dims, bits // 8, lambda z0, rows: B.sphere_torch((dims[0], dims[1], rows), (0, 0, z0), dims, bits, dev),
                                  gpu_mem_bytes=args.pool_mb << 20)
    else:
        raw = B.sphere_torch(dims, (0, 0, 0), dims, bits, dev)
        vol = raw.cpu().numpy() if bits == 8 else raw.view(torch.int16).cpu().numpy().view(np.uint16)
        del raw
        torch.cuda.empty_cache()
        r.loadRawVolume(vol, brick_size=64, overlap=2, out_of_core=True, gpu_mem_bytes=args.pool_mb << 20)
        del vol
    t_load = time.perf_counter() - t0
    torch.cuda.empty_cache()
    info = r.datasetInfo()
    tf, lo, hi = B.tf1d_std(info.tf_size)
    r.set1DTrans(tf)
    r.setLightingEnabled(False)
    r.resize((w, h))
    r.setRendererTarget(RT_HEADLESS)
    V = lambda *a: np.array(a, np.float64)               # noqa: E731
    keys = [dict(eye=V(0, 0, 2.2), ref=V(0, 0, -1), vup=V(0, 1, 0)),
            dict(eye=V(0, 0, 1.6), ref=V(0, 0, -1), vup=V(0, 1, 0)),
            dict(eye=V(0.25, 0.1, 0.75), ref=V(-0.3, -0.1, -1), vup=V(0, 1, 0)),
            dict(eye=V(0.45, 0.15, 0.1), ref=V(-1, -0.3, -0.4), vup=V(0, 1, 0)),
            dict(eye=V(0.1, 0.3, -0.45), ref=V(-0.2, -0.6, 1), vup=V(0, 1, 0.2)),
            dict(eye=V(-0.5, 0.2, -0.9), ref=V(0.5, -0.2, 1), vup=V(0, 1, 0)),
            dict(eye=V(-0.9, 0.1, -1.4), ref=V(0.6, -0.1, 1), vup=V(0, 1, 0))]
    path = camera_path(keys, args.frames)
    passes = {}
    clock = B.ClockSampler(0)
    clock.start()
    for name in ("cold", "warm"):
        if name == "cold":
            r.clearWorkingSet()
        p0 = r.perf()
        ms, subs = [], []
        t_pass = time.perf_counter()
        for cam in path:
            r.setViewPos(cam["eye"]); r.setViewDir(cam["ref"]); r.setUpDir(cam["vup"])      # noqa: E702
            tf0 = time.perf_counter()
            subs.append(r.paintUntilConverged(4000))
            ms.append(1e3 * (time.perf_counter() - tf0))
        dt = time.perf_counter() - t_pass
        p1 = r.perf()
        paged = p1["paged_bricks"] - p0["paged_bricks"]
        up = p1["uploaded_bytes"] - p0["uploaded_bytes"]
        passes[name] = {"ms_per_frame_mean": float(np.mean(ms)), "ms_per_frame_median": float(np.median(ms)),
                        "ms_per_frame_max": float(np.max(ms)), "subframes_per_frame": float(np.mean(subs)),
                        "paged_bricks_per_frame": paged / len(path), "uploaded_GB": up / 1e9,
                        "h2d_GBps_over_pass": up / dt / 1e9, "raycast_ms_per_frame": (p1["raycast_ms"] - p0["raycast_ms"]) / len(path),
                        "samples_per_s": (p1["samples"] - p0["samples"]) / dt}
    clocks = clock.stop()
    cap = list(info.pool_capacity)
    voxels = float(dims[0]) * dims[1] * dims[2]
    host = {}
    try:
        with open("/proc/meminfo") as f:
            host = {ln.split(":")[0]: int(ln.split()[1]) for ln in f if ln.split(":")[0] in ("MemTotal", "MemAvailable")}
    except OSError:
        pass
    out = {"bench": "out-of-core fly-through (BASELINE configs[4], FlyThroughAnimator.lua path)", "volume": dims, "bpv": bits // 8,
           "raw_GB": voxels * bits / 8 / 1e9, "resolution": [w, h], "frames": len(path), "total_bricks": int(info.total_bricks),
           "pool_slots": cap[0] * cap[1] * cap[2], "pool_capacity": cap, "pool_budget_MB": args.pool_mb,
           "load_s_incl_generation_build_and_host_store": t_load, "streamed_build": bool(args.streamed),
           "builder_kernel_ms": build_ms,
           "builder_algorithmic_GBps": (2.387 * bits / 8 * voxels / (build_ms * 1e-3) / 1e9) if build_ms else None,
           "host_mem_kB": host, "clocks": clocks, "cold": passes["cold"], "warm": passes["warm"]}
    print(json.dumps(out))
    r.cleanup()


if __name__ == "__main__":
    main()
