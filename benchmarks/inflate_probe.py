#!/usr/bin/env python
"""Device-side brick inflate (tvk_zlib_inflate / tvk_lz4_inflate): N bricks of 64^3 uint16 (sphere data and blob data), compressed
on the host, inflated by one warp per brick; wall time of the call (it includes the copies of the compressed bytes up and of
the inflated bytes down) and the inflated GB/s.  Usage: python benchmarks/inflate_probe.py [bricks]"""
import ctypes as C
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from imagevis3d_b200.context import TvkContext
    from oracle import host as H
    from tests.util import lz4_compress_block
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    tvk = TvkContext(0)
    base = [H.synth_blobs((64, 64, 64), np.uint16, seed=s).tobytes() for s in range(4)]
    base.append(H.gen_sphere((256, 256, 256), np.uint16)[96:160, 96:160, 96:160].copy().tobytes())
    for codec, comp, fn in (("zlib", lambda d: zlib.compress(d, 6), tvk.lib.tvk_zlib_inflate), ("lz4", lz4_compress_block, tvk.lib.tvk_lz4_inflate)):
        cb = [comp(d) for d in base]
        blocks = [cb[i % len(cb)] for i in range(n)]
        sizes = [len(base[i % len(base)]) for i in range(n)]
        bufs = [np.frombuffer(b, np.uint8).copy() for b in blocks]
        outs = [np.zeros(s, np.uint8) for s in sizes]
        src = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        dst = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        cl, dl = (C.c_uint32 * n)(*[len(b) for b in blocks]), (C.c_uint32 * n)(*sizes)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            rc = fn(tvk.h, C.c_uint32(n), src, cl, dst, dl)
            best = min(best, time.perf_counter() - t0)
            assert rc == 0
        assert all(o.tobytes() == base[i % len(base)] for i, o in enumerate(outs))
        print(json.dumps({"codec": codec, "bricks": n, "compressed_MB": sum(len(b) for b in blocks) / 1e6, "inflated_MB": sum(sizes) / 1e6,
                          "ratio": sum(sizes) / sum(len(b) for b in blocks), "call_ms": best * 1e3, "inflated_GBps_incl_copies": sum(sizes) / best / 1e9}))
    tvk.close()


if __name__ == "__main__":
    main()
