"""Generates the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and oracle/_ref/ref_driver,
built by `make -C oracle/ref_build`).  The fixtures travel with the repo, so the
tests on the GPU box never touch /root/reference.

  python tests/golden/make_golden.py

Outputs
  builder_cases.npz  raw inputs + the reference's bricks / min-max / histograms for
                     a handful of small volumes (odd sizes, tiny last bricks, median)
  c60.npz            DemoData/c60.uvf decoded by the reference's UVF reader: the
                     128^3 uint8 brick, its 1-D histogram, global min/max, and
                     c60.1dt as the RGBA8 table TransferFunction1D::GetByteArray yields
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

CASES = [  # (sx,sy,sz), dtype, brick, overlap, median
    ((8, 8, 1), "u16", 16, 2, 1),      # Tuvok/IO/test/rebricking.h:11-34
    ((20, 12, 9), "u16", 12, 2, 0),
    ((33, 17, 9), "u8", 12, 2, 0),     # 33 % 8 == 1: last brick thinner than the overlap
    ((41, 25, 13), "u8", 12, 2, 1),
    ((25, 25, 25), "u16", 12, 2, 0),
    ((70, 65, 61), "u8", 20, 2, 0),
    ((130, 70, 61), "u8", 64, 2, 0),
]


def main():
    drv = O.ref_driver_path()
    if drv is None or not os.path.isdir(REF):
        raise SystemExit("needs oracle/_ref/ref_driver and /root/reference")
    rng = np.random.default_rng(20261017)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for i, (vol, ty, brick, ov, med) in enumerate(CASES):
            dt = np.uint8 if ty == "u8" else np.uint16
            if vol == (8, 8, 1):
                raw = np.arange(64, dtype=np.uint16).reshape(1, 8, 8)
            else:
                # smooth-ish random field so that the 2-D histogram is not degenerate
                raw = rng.integers(0, 256 if ty == "u8" else 4096, size=vol[::-1]).astype(dt)
                if raw.size > 100000:  # keep the fixture small: low-entropy palette
                    raw = (rng.integers(0, 16, size=vol[::-1]) * 17).astype(dt)
            rawf, binf = os.path.join(td, "in.raw"), os.path.join(td, "out.bin")
            raw.tofile(rawf)
            subprocess.check_call([drv, "convert", rawf, *map(str, vol), ty, str(brick), str(ov),
                                   str(med), "0", binf, "1"], stderr=subprocess.DEVNULL)
            ref = O.parse_ref_convert(binf, raw.dtype.itemsize)
            k = "c%d_" % i
            out[k + "raw"] = raw
            out[k + "params"] = np.array([brick, ov, med], np.int64)
            out[k + "lod_size"] = ref["lod_size"]
            out[k + "lod_bricks"] = ref["lod_bricks"]
            out[k + "lod_aspect"] = ref["lod_aspect"]
            out[k + "sizes"] = ref["sizes"]
            out[k + "minmax"] = ref["minmax"]
            out[k + "store"] = ref["store"]
            out[k + "hist1d"] = ref["hist1d"]
            out[k + "hist2d"] = ref["hist2d"]
            out[k + "max_grad"] = np.float32(ref["max_grad"])
        out["n_cases"] = np.int64(len(CASES))
        np.savez_compressed(os.path.join(HERE, "builder_cases.npz"), **out)

        uvff, tff = os.path.join(td, "c60.bin"), os.path.join(td, "c60tf.bin")
        subprocess.check_call([drv, "uvfinfo", os.path.join(REF, "DemoData/c60.uvf"), uvff])
        subprocess.check_call([drv, "tf1d", os.path.join(REF, "DemoData/c60.1dt"), tff])
        b = open(uvff, "rb").read()
        assert b[:8] == b"TVKUVF1\0"
        o = 8
        dom = np.frombuffer(b, np.uint64, 3, o); o += 24
        bs = np.frombuffer(b, np.uint64, 3, o); o += 24
        bits = int(np.frombuffer(b, np.uint64, 1, o)[0]); o += 8
        n = int(np.frombuffer(b, np.uint64, 1, o)[0]); o += 8
        data = np.frombuffer(b, np.uint8, n, o); o += n
        n1 = int(np.frombuffer(b, np.uint64, 1, o)[0]); o += 8
        h1 = np.frombuffer(b, np.uint64, n1, o); o += 8 * n1
        mn, mx = np.frombuffer(b, np.float64, 2, o); o += 16
        mg = np.frombuffer(b, np.float32, 1, o)[0]
        t = open(tff, "rb").read()
        tn = int(np.frombuffer(t, np.uint64, 1, 0)[0])
        tf = np.frombuffer(t, np.uint8, tn * 4, 8).reshape(tn, 4)
        lim = np.frombuffer(t, np.uint64, 2, 8 + tn * 4)
        np.savez_compressed(os.path.join(HERE, "c60.npz"), domain=dom, brick_size=bs, bits=np.int64(bits),
                            brick=data.reshape(int(bs[2]), int(bs[1]), int(bs[0])), hist1d=h1,
                            minmax=np.array([mn, mx]), max_grad=np.float32(mg), tf1d_rgba8=tf,
                            tf1d_nonzero=lim)
    for f in ("builder_cases.npz", "c60.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
