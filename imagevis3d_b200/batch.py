#!/usr/bin/env python
"""BatchRenderer for the B200 renderer: the flow of the reference's LuaScripts/BatchRenderer.lua / BatchRenderer
binary (BatchRenderer/main.cpp) with the same call sequence on the Lua-named surface of imagevis3d_b200.renderer --

    renderer = tuvok.renderer.new(type, ...)     ->  Renderer(device); setRendererType(...)
    renderer.loadDataset(dataset)                ->  loadDataset(file.uvf)   (UVF v5 TOC files; legacy single-LOD
                                                                              RasterDataBlock files are re-bricked)
    renderer.addShaderPath(dir)                  ->  no-op
    renderer.initialize(context); resize({w,h})  ->  resize((w, h))
    [get1DTrans / changed1DTrans]                ->  load1DTrans(file.1dt)   (the reference auto-loads <dataset>.1dt)
    setRendererTarget(RT_Headless); paint()      ->  repeat paint() until converged()   (ray-guided path is progressive)
    setRendererTarget(RT_Capture); captureSingleFrame(png, true)

Usage:  python -m imagevis3d_b200.batch DATASET.uvf OUT.png [--tf FILE.1dt] [--size WxH] [--renderer gridleaper|raycaster]
                                         [--lighting 0|1] [--rotate-y DEG] [--rotate-x DEG] [--device N]
                                         [--iso VALUE [--clearview CVISO]] [--stereo rb|scanline|sbs|af [--eye-dist D]]
                                         [--slice sagittal|axial|coronal:INDEX | --mip sagittal|axial|coronal[:ANGLE]]
--iso: isosurface mode (ClearView needs the raycaster); --stereo: both eyes, composed; --slice: a 2-D slice view of the finest
single-brick level of detail; --mip: the capture mode's full-resolution MIP, rotated by ANGLE degrees (orthographic).
Prints one JSON line with the timings."""
import argparse
import json
import math
import os
import sys
import time

import numpy as np


def rotation(axis, deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    m = np.eye(4, dtype=np.float32)
    if axis == "y":
        m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, -s, s, c
    else:
        m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, s, -s, c
    return m


def main(argv=None):
    ap = argparse.ArgumentParser(description="render one frame of a UVF dataset to a PNG")
    ap.add_argument("dataset")
    ap.add_argument("output")
    ap.add_argument("--tf", default=None, help="1-D transfer function (.1dt); default: <dataset without extension>.1dt if present")
    ap.add_argument("--size", default="1024x1024")
    ap.add_argument("--renderer", default="gridleaper", choices=["gridleaper", "raycaster"])
    ap.add_argument("--lighting", type=int, default=1)
    ap.add_argument("--rotate-y", type=float, default=0.0)
    ap.add_argument("--rotate-x", type=float, default=0.0)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--iso", type=float, default=None, help="isosurface mode with this isovalue (data units)")
    ap.add_argument("--clearview", type=float, default=None, help="ClearView lens showing the surface at this isovalue (raycaster, with --iso)")
    ap.add_argument("--stereo", default=None, choices=["rb", "scanline", "sbs", "af"])
    ap.add_argument("--eye-dist", type=float, default=0.02)
    ap.add_argument("--slice", default=None, help="sagittal|axial|coronal:INDEX -- a 2-D slice view instead of the 3-D one")
    ap.add_argument("--mip", default=None, help="sagittal|axial|coronal[:ANGLE] -- full-resolution MIP instead of the 3-D view")
    args = ap.parse_args(argv)
    from .renderer import Renderer, RT_CAPTURE, RT_HEADLESS
    w, h = [int(v) for v in args.size.lower().split("x")]
    t0 = time.perf_counter()
    renderer = Renderer(args.device)
    renderer.loadDataset(args.dataset)
    renderer.addShaderPath("")
    renderer.setRendererType(args.renderer == "raycaster")
    tf = args.tf or (os.path.splitext(args.dataset)[0] + ".1dt")
    if os.path.exists(tf):
        renderer.load1DTrans(tf)                      # GLRenderer::Initialize auto-load (GLRenderer.cpp:174-197)
    renderer.setLightingEnabled(bool(args.lighting))
    renderer.resize((w, h))
    if args.rotate_y or args.rotate_x:
        renderer.setRotation4x4(rotation("y", args.rotate_y) @ rotation("x", args.rotate_x))
    modes = {"sagittal": 0, "axial": 1, "coronal": 2}
    if args.iso is not None:
        from .renderer import RM_ISOSURFACE
        renderer.setRenderMode(RM_ISOSURFACE)
        renderer.setIsoValue(args.iso)
        if args.clearview is not None:
            renderer.setClearViewEnabled(True)
            renderer.setCVIsoValue(args.clearview)
    if args.stereo:
        renderer.setStereoEnabled(True)
        renderer.setStereoMode(args.stereo)
        renderer.setStereoEyeDist(args.eye_dist)
    t1 = time.perf_counter()
    subframes = 0
    if args.slice or args.mip:
        name, _, val = (args.slice or args.mip).partition(":")
        if name not in modes:
            raise SystemExit("--slice / --mip: sagittal, axial or coronal")
        if args.slice:
            renderer.render2DView(modes[name], int(val or 0))
        else:
            renderer.setRendererTarget(RT_CAPTURE)
            renderer.renderHQMIP(modes[name], float(val or 0.0))
        subframes = 1
    else:
        renderer.setRendererTarget(RT_HEADLESS)
        while True:
            renderer.paint()
            subframes += 1
            if renderer.converged() or subframes >= 10000:
                break
    t2 = time.perf_counter()
    renderer.setRendererTarget(RT_CAPTURE)
    renderer.captureSingleFrame(args.output, not (args.slice or args.mip or args.stereo))
    perf = renderer.perf()
    info = renderer.datasetInfo()
    print(json.dumps({"dataset": args.dataset, "domain": list(info.domain), "renderer": args.renderer, "size": [w, h],
                      "subframes": subframes, "load_s": t1 - t0, "render_ms": 1e3 * (t2 - t1), "raycast_kernel_ms": perf["raycast_ms"],
                      "samples": perf["samples"], "paged_bricks": perf["paged_bricks"], "output": args.output}))
    renderer.cleanup()
    return 0


if __name__ == "__main__":
    sys.exit(main())
