#!/usr/bin/env python
"""2-D histogram (K2b) timing, CUDA events around both passes (tvk_build_hist2d_last_ms), on the bench volume (the
reference's CT_SPHERE) and on a noisy 12-bit volume (smooth blobs + noise: nearly every voxel has its own bin).
usage: hist2d_probe.py [bpv] [n] ; TVK_HIST2D_V4=1 / TVK_HIST2D_SCALAR=1 select the older kernels"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench as B
from imagevis3d_b200.context import TvkContext
tvk = TvkContext(0)
dev = torch.device("cuda", 0)
bpv = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dims = (n, n, n)
bits = 8 * bpv


def noisy(n, bits):
    g = torch.Generator(device=dev).manual_seed(5)
    out = torch.empty((n, n, n), dtype=torch.uint8 if bits == 8 else torch.uint16, device=dev)
    ax = torch.arange(n, device=dev, dtype=torch.float32) / n
    top = 255.0 if bits == 8 else 4095.0
    for z0 in range(0, n, 32):
        z = ax[z0:z0 + 32][:, None, None]
        v = 0.5 + 0.25 * torch.sin(9 * ax)[None, None, :] * torch.cos(7 * ax)[None, :, None] + 0.2 * torch.sin(11 * z)
        v = v * top * 0.9 + torch.randn((32, n, n), device=dev, generator=g) * (top * 0.01)
        out[z0:z0 + 32] = v.clamp(0, top).to(torch.int32).to(out.dtype)
    return out


for name, gen in (("sphere", lambda: B.sphere_torch(dims, (0, 0, 0), dims, bits, dev)), ("noisy", lambda: noisy(n, bits))):
    raw = gen()
    torch.cuda.synchronize()
    tvk.build_octree(raw_dptr=raw.data_ptr(), volume_size=dims, bpv=bpv)
    mx = float(tvk.build_minmax()[:, 1].max())
    best = 1e9
    for it in range(3):
        h, mg = tvk.build_hist2d(256 if bpv == 1 else 4096, mx)
        best = min(best, tvk.build_hist2d_last_ms())
    print(json.dumps({"volume": name, "bpv": bpv, "n": n, "v4": os.environ.get("TVK_HIST2D_V4", "0"), "lib": os.environ.get("TVK_CUDA_LIB", ""),
                      "hist2d_kernels_ms": best, "votes": int(h.sum()), "nonzero_bins": int((h != 0).sum()), "max_grad": float(mg),
                      "checksum": int((h.reshape(-1) * (np.arange(h.size, dtype=np.uint64) % np.uint64(65521) + np.uint64(1))).sum() % np.uint64(1 << 61))}), flush=True)
    tvk.build_release()
    del raw
    torch.cuda.empty_cache()
