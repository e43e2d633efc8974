"""Per-sub-frame breakdown of bench.py's e2e leg (tvkr_* only): wall time of each paint(), march ms, bricks paged,
tuvok.perf counters.  Usage: python benchmarks/e2e_probe.py [--vol 2048] [--hash 4093] [--strategy 3] [--frames 4]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vol", type=int, default=2048)
    ap.add_argument("--hash", type=int, default=4093)
    ap.add_argument("--rehash", type=int, default=16)
    ap.add_argument("--strategy", type=int, default=3)
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--mode", default="2d_lit")
    args = ap.parse_args()
    import torch
    import bench
    from imagevis3d_b200 import hostmath as HM
    from imagevis3d_b200.renderer import Renderer
    dev = torch.device("cuda:0")
    vol = (args.vol,) * 3
    W, Hh = 1920, 1080
    raw = bench.sphere_torch(vol, (0, 0, 0), vol, 16, dev)
    host = np.empty((vol[2], vol[1], vol[0]), np.uint16)
    for z0 in range(0, vol[2], 64):
        host[z0:z0 + 64] = raw[z0:z0 + 64].view(torch.int16).cpu().numpy().view(np.uint16)
    del raw
    torch.cuda.empty_cache()
    r = Renderer(0)
    r.setState(brick_strategy=args.strategy, hash_table_size=args.hash, rehash_count=args.rehash, sampler=0xFFFFFFFF)
    r.loadRawVolume(host, out_of_core=True, hist2d=True)
    del host
    r.resize((W, Hh))
    tf, nz = bench.tf2d(4096)
    if args.mode == "2d_lit":
        r.set2DTrans(tf, nz)
        r.setRenderMode(1)
        r.setLightingEnabled(True)
    else:
        r.set1DTrans(bench.tf1d_std(65536) if hasattr(bench, "tf1d_std") else None)
        r.setRenderMode(0)
    names = {1: "render", 2: "raycast", 3: "read_htable", 6: "upload_bricks", 7: "pool_sort", 9: "pool_get_brick",
             19: "pool_upload_brick", 21: "pool_upload_meta"}
    out = []
    for k in range(args.frames):
        r.setRotation4x4((HM.rotation_y(bench.orbit_angle(k, 20)) @ HM.rotation_x(0.35)).astype(np.float32))
        t00 = time.perf_counter()
        r.clearWorkingSet()
        if args.mode == "2d_lit":
            r.set2DTrans(tf, nz)
        t_setup = time.perf_counter() - t00
        subs = []
        p_prev = r.perf()
        for c in names:
            r.perfQuery(c)
        while True:
            t0 = time.perf_counter()
            r.paint()
            dt = time.perf_counter() - t0
            p = r.perf()
            rec = {"wall_ms": round(1e3 * dt, 3), "march_ms": round(p["raycast_ms"] - p_prev["raycast_ms"], 3),
                   "paged": int(p["paged_bricks"] - p_prev["paged_bricks"]),
                   "Msamples": round((p["samples"] - p_prev["samples"]) / 1e6, 2)}
            for c, nm in names.items():
                rec[nm] = round(r.perfQuery(c), 3)
            p_prev = p
            subs.append(rec)
            if r.converged() or len(subs) > 100:
                break
        t0 = time.perf_counter()
        img = r.frameRGBA8()
        t_read = time.perf_counter() - t0
        total = time.perf_counter() - t00
        out.append({"frame": k, "total_ms": round(1e3 * total, 3), "setup_ms": round(1e3 * t_setup, 3),
                    "readback_ms": round(1e3 * t_read, 3), "subframes": subs})
        print(json.dumps(out[-1]), flush=True)
    r.cleanup()


if __name__ == "__main__":
    main()
