#!/usr/bin/env python
"""per-kernel times of the 2-D histogram on the bench volume (ncu-free: CUDA events around both passes)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from imagevis3d_b200.context import TvkContext
tvk = TvkContext(0)
dev = torch.device("cuda", 0)
bpv = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dims = (n, n, n)
raw = torch.empty((n, n, n), dtype=torch.uint8 if bpv == 1 else torch.uint16, device=dev)
for z0 in range(0, n, 64):
    s = B.volume_torch(dims, dev, z0=z0, zn=64, ztotal=n)
    raw[z0:z0 + 64] = s if bpv == 1 else (s.to(torch.int32) * 257).to(torch.uint16)
torch.cuda.synchronize()
tvk.build_octree(raw_dptr=raw.data_ptr(), volume_size=dims, bpv=bpv)
mx = float(tvk.build_minmax()[:, 1].max())
for it in range(2):
    tvk.build_hist2d(256 if bpv == 1 else 4096, mx)
print(json.dumps({"bpv": bpv, "n": n, "dbg": os.environ.get("TVK_HIST2D_DBG", "0"), "scalar": os.environ.get("TVK_HIST2D_SCALAR", "0"), "hist2d_kernels_ms": tvk.build_hist2d_last_ms()}))
