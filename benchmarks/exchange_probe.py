"""Times the stripe compositor (exchange_composite_kernel) alone: `world` ranks emulated on ONE GPU at 3840x2160
(no NVLink involved), CUDA events around the launch.  Usage: python benchmarks/exchange_probe.py [world]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                   # noqa: E402
from imagevis3d_b200 import hostmath as HM                     # noqa: E402
from imagevis3d_b200 import sortlast as SL                     # noqa: E402
from imagevis3d_b200.context import TvkContext                 # noqa: E402
from oracle import host as H                                   # noqa: E402
from tests.util import resident_pool                           # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
G = (128, 128, 64 * world)
vol = H.synth_blobs(G, np.uint8, seed=23)
parts = SL.slab_partition(G, world, halo=32)
tf = H.tf1d_std(256, 0.3, 0.3)
tf[:, 3] *= 0.15
lo, hi = H.tf_nonzero_limits(tf)
W, Hh = 3840, 2160
ctxs = []
for r, part in enumerate(parts):
    t = TvkContext(0)
    resident_pool(t, np.ascontiguousarray(vol[part["zs"]:part["ze"]]), 64, H.tf_bytes(tf), float(lo), float(hi))
    t.exchange_create(world, r, W, Hh)
    ctxs.append(t)
bases = [t.exchange_local_base() for t in ctxs]
for t in ctxs:
    t.exchange_set_peers(bases)
out = []
for frame in range(1, 7):
    cam = HM.Camera()
    cam.rotation = HM.rotation_y(0.3 + 0.2 * frame) @ HM.rotation_x(0.25)
    pg = HM.raycast_params(cam, W, Hh, G)
    order = SL.depth_order(np.array(pg.eye_to_model[:]).reshape(4, 4), parts)
    for r, t in enumerate(ctxs):
        t.exchange_begin_frame(frame)
        t.raycast(HM.raycast_params(cam, W, Hh, G, slab=parts[r]))
        t.synchronize()
    torch.cuda.synchronize()
    ms = []
    for t in ctxs:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t.set_stream(torch.cuda.current_stream().cuda_stream)
        a.record()
        t.exchange_composite(order, frame)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    ctxs[0].exchange_final(frame)
    torch.cuda.synchronize()
    out.append(ms)
print(json.dumps({"world": world, "resolution": [W, Hh], "composite_ms_per_rank_per_frame": out}))
