#!/usr/bin/env python
"""Golden vectors for the re-bricking path, produced by the REFERENCE ITSELF: oracle/_ref/ref_driver
`writeuvf` (the reference's UVF writer) makes a source file, `rebrick` opens it with the reference's UVFDataset, wraps it
in the reference's DynamicBrickingDS (IO/DynamicBrickingDS.cpp) and dumps every target brick and MaxMinForKey() in each
of the three min/max modes.  Build container only; the result is tests/golden/rebrick_cases.npz."""
import gzip
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
# name: (volume x,y,z, dtype, source brick, overlap, [target brick sizes INCLUDING ghost])
CASES = {
    "u8_41x37x29_src20": ((41, 37, 29), np.uint8, 20, 2, [(12, 12, 12), (20, 12, 8)]),
    "u16_33x18x25_src36": ((33, 18, 25), np.uint16, 36, 2, [(12, 20, 8), (36, 36, 36)]),
}


def case_volume(name):
    vol, dt = CASES[name][0], CASES[name][1]
    rng = np.random.default_rng(sum(vol))
    z, y, x = np.meshgrid(np.arange(vol[2]), np.arange(vol[1]), np.arange(vol[0]), indexing="ij")
    top = 255 if dt == np.uint8 else 4095
    v = (np.sin(x * 0.31) * np.cos(y * 0.23) + np.sin(z * 0.4)) * 0.25 + 0.5
    return np.clip(v * top + rng.integers(-3, 4, v.shape), 0, top).astype(dt)


def parse(blob, bpv):
    assert blob[:8] == b"TVKRBK1\0"
    lods = struct.unpack_from("<I", blob, 8)[0]
    o = 12
    layout, domain = [], []
    for _ in range(lods):
        t = struct.unpack_from("<3I3Q", blob, o); o += 36
        layout.append(t[:3]); domain.append(t[3:])
    total = struct.unpack_from("<Q", blob, o)[0]; o += 8
    sizes = np.zeros((total, 3), np.uint32); mm = np.zeros((total, 2), np.float64)
    for i in range(total):
        t = struct.unpack_from("<3I2d", blob, o); o += 28
        sizes[i] = t[:3]; mm[i] = t[3:]
    n = int((sizes.astype(np.uint64).prod(axis=1)).sum()) * bpv
    payload = np.frombuffer(blob, np.uint8, n, o)
    assert o + n == len(blob)
    return np.array(layout, np.uint32), np.array(domain, np.uint64), sizes, mm, payload


def main():
    out = {"names": np.array(list(CASES))}
    for name, (vol, dt, brick, overlap, targets) in CASES.items():
        raw = case_volume(name)
        bpv = raw.dtype.itemsize
        with tempfile.TemporaryDirectory() as td:
            f = os.path.join(td, "v.raw")
            raw.tofile(f)
            uvf = os.path.join(td, "o.uvf")
            subprocess.check_call([O.ref_driver_path(), "writeuvf", f, str(vol[0]), str(vol[1]), str(vol[2]),
                                   "u16" if bpv == 2 else "u8", str(brick), str(overlap), "0", "0", "1", uvf])
            out[name + "/uvf_gz"] = np.frombuffer(gzip.compress(open(uvf, "rb").read(), mtime=0), np.uint8)
            out[name + "/targets"] = np.array(targets, np.uint32)
            for ti, t in enumerate(targets):
                for mode in (0, 1, 2):
                    dump = os.path.join(td, "rb.bin")
                    subprocess.check_call([O.ref_driver_path(), "rebrick", uvf, str(t[0]), str(t[1]), str(t[2]), str(mode), dump])
                    layout, domain, sizes, mm, payload = parse(open(dump, "rb").read(), bpv)
                    k = "%s/t%d" % (name, ti)
                    if mode == 0:
                        out[k + "/layout"], out[k + "/domain"], out[k + "/sizes"], out[k + "/payload"] = layout, domain, sizes, payload
                    else:
                        assert (out[k + "/payload"] == payload).all() and (out[k + "/sizes"] == sizes).all()
                    out[k + "/minmax%d" % mode] = mm
            # what the reference refuses: a target core that does not divide the source core
            rc = subprocess.call([O.ref_driver_path(), "rebrick", uvf, "11", "12", "12", "0", os.path.join(td, "x.bin")],
                                 stderr=subprocess.DEVNULL)
            assert rc == 3
    np.savez_compressed(os.path.join(GOLD, "rebrick_cases.npz"), **out)
    print(os.path.getsize(os.path.join(GOLD, "rebrick_cases.npz")))


if __name__ == "__main__":
    main()
