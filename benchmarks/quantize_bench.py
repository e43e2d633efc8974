#!/usr/bin/env python
"""B5 quantisation throughput: tvk_quantize on volumes resident in HBM (device_ms = CUDA events around all passes).
Algorithmic bytes: one read of the input per pass that is needed + one write of the output."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from imagevis3d_b200.context import TvkContext
    tvk = TvkContext(0)
    n = 1024 ** 3
    g = torch.Generator(device="cuda").manual_seed(1)
    cases = [("int16 CT-like (bias to 0)", np.int16, lambda: (torch.randn(n, device="cuda", generator=g) * 400 + 200).clamp(-1024, 3071).to(torch.int16)),
             ("uint16 full range (histogram only)", np.uint16, lambda: torch.randint(0, 65536, (n,), device="cuda", generator=g, dtype=torch.int32).to(torch.uint16)),
             ("float32 smooth (min/max + scale)", np.float32, lambda: torch.rand(n, device="cuda", generator=g) * 3.5 - 1.25),
             ("int8 (signed -> unsigned)", np.int8, lambda: torch.randint(-128, 128, (n,), device="cuda", generator=g, dtype=torch.int32).to(torch.int8))]
    dst = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    for name, dt, make in cases:
        src = make()
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            r = tvk.quantize(src_dptr=src.data_ptr(), dtype=dt, count=n, dst_dptr=dst.data_ptr())
            best = r if best is None or r["device_ms"] < best["device_ms"] else best
        w = np.dtype(dt).itemsize
        print(json.dumps({"bench": "quantize", "case": name, "elements": n, "device_ms": best["device_ms"], "generated": best["generated"],
                          "out_bits": best["out_bits"], "elements_per_s": n / (best["device_ms"] * 1e-3),
                          "input_GBps_per_pass_equivalent": n * w / (best["device_ms"] * 1e-3) / 1e9}), flush=True)
        del src
    tvk.close()


if __name__ == "__main__":
    main()
