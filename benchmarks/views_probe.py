#!/usr/bin/env python
"""The f.3 views on a 1024^3 uint16 sphere volume (reference CT_SPHERE generator), 1920x1080, through the renderer API: 2-D slice
view, MIP preview, the capture mode's full-resolution MIP (orthographic), a ClearView frame on the classic renderer, and a
stereo frame on the ray-guided renderer (resident working set).  Wall time around the call + frame read-back sync; one JSON
line per view.  Usage: python benchmarks/views_probe.py [size]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, sync, n=5):
    fn(); sync()
    best = 1e9
    for _ in range(n):
        t0 = time.perf_counter()
        fn(); sync()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def main():
    import torch
    import bench
    from imagevis3d_b200 import hostmath as HM
    from imagevis3d_b200.renderer import Renderer, RT_CAPTURE, RT_HEADLESS, RM_ISOSURFACE
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    dev = torch.device("cuda:0")
    raw = bench.sphere_torch((N, N, N), (0, 0, 0), (N, N, N), 16, dev)
    host = raw.view(torch.int16).cpu().numpy().view(np.uint16)
    del raw
    torch.cuda.empty_cache()
    r = Renderer(0)
    r.loadRawVolume(host, brick_size=64, overlap=2)
    del host
    W, Hh = 1920, 1080
    r.resize((W, Hh))
    tf = np.zeros((4096, 4), np.uint8)
    tf[:, 0] = np.arange(4096) // 16
    tf[:, 1] = 255 - np.arange(4096) // 16
    tf[:, 2] = 80
    tf[1024:, 3] = 40
    r.set1DTrans(tf)
    sync = lambda: torch.cuda.synchronize()        # noqa: E731
    out = []
    out.append({"view": "slice (axial, finest single-brick LOD)", "ms": timed(lambda: r.render2DView(1, N // 2), sync)})
    out.append({"view": "MIP preview (axial)", "ms": timed(lambda: r.render2DView(1, N // 2, mip=True), sync)})
    r.setRendererTarget(RT_CAPTURE)
    p, bricks, lod = r.planHQMIP(2, 30.0)
    out.append({"view": "full-resolution MIP, orthographic, 30 degrees", "lod": lod, "bricks": len(bricks),
                "ms": timed(lambda: r.renderHQMIP(2, 30.0), sync, 3)})
    r.setRendererTarget(RT_HEADLESS)
    r.setRotation4x4((HM.rotation_y(0.6) @ HM.rotation_x(0.35)).astype(np.float32))
    # ClearView on the classic renderer (capture target: finest LOD, one paint = the frame)
    r.setRendererType(True)
    r.setRendererTarget(RT_CAPTURE)
    r.setRenderMode(RM_ISOSURFACE)
    r.setIsoValue(0.3 * 65535)

    def classic_frame():
        r.setViewPos((0.0, 0.0, 1.6))              # schedules a complete redraw
        r.paintUntilConverged(50)
    out.append({"view": "classic isosurface (LOD 0, every visible brick)", "ms": timed(classic_frame, sync, 3)})
    r.setClearViewEnabled(True)
    r.setCVIsoValue(0.6 * 65535)
    out.append({"view": "classic isosurface + ClearView", "ms": timed(classic_frame, sync, 3)})
    r.setClearViewEnabled(False)
    # stereo on the ray-guided renderer, working set resident
    r.setRendererType(False)
    r.setRendererTarget(RT_HEADLESS)
    r.setRenderMode(0)
    r.setLightingEnabled(True)

    def guided_frame():
        r.setViewPos((0.0, 0.0, 1.6))
        return r.paintUntilConverged(200)
    guided_frame(); guided_frame()
    out.append({"view": "ray-guided 1D-TF lit, mono, resident working set", "ms": timed(guided_frame, sync, 3)})
    r.setStereoEnabled(True)
    guided_frame(); guided_frame()
    out.append({"view": "ray-guided 1D-TF lit, stereo (anaglyph), resident working set", "ms": timed(guided_frame, sync, 3)})
    for o in out:
        o.update(volume=[N, N, N], voxel_type="uint16", resolution=[W, Hh])
        print(json.dumps(o))
    r.cleanup()


if __name__ == "__main__":
    main()
