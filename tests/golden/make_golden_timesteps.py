#!/usr/bin/env python
"""Golden vectors for time-dependent UVF files (setTimestep), produced by the REFERENCE: ref_driver
`writeuvfts` assembles a two-step file from the reference's block classes in the order of RAWConverter's time-step loop,
`dsdump` reads each step back through the reference's UVFDataset.  Build container only -> tests/golden/timesteps_case.npz."""
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

VOL, BRICK, OVERLAP = (29, 22, 17), 16, 2


def volumes():
    z, y, x = np.meshgrid(np.arange(VOL[2]), np.arange(VOL[1]), np.arange(VOL[0]), indexing="ij")
    a = (120 + 100 * np.sin(x * 0.4) * np.cos(y * 0.3 + z * 0.2)).astype(np.uint8)
    b = (np.clip(250 - 9 * np.sqrt((x - 14.0) ** 2 + (y - 11.0) ** 2 + (z - 8.0) ** 2), 0, 255)).astype(np.uint8)
    return [a, b]


def parse(blob):
    assert blob[:8] == b"TVKDSD1\0"
    steps, lods, total = struct.unpack_from("<QIQ", blob, 8)
    o = 28
    sizes = np.zeros((total, 3), np.uint32); mm = np.zeros((total, 2))
    for i in range(total):
        t = struct.unpack_from("<3I2d", blob, o); o += 28
        sizes[i] = t[:3]; mm[i] = t[3:]
    n = int(sizes.astype(np.uint64).prod(axis=1).sum())
    assert o + n == len(blob)
    return steps, sizes, mm, np.frombuffer(blob, np.uint8, n, o)


def main():
    out = {}
    with tempfile.TemporaryDirectory() as td:
        raws = []
        for t, v in enumerate(volumes()):
            f = os.path.join(td, "t%d.raw" % t)
            v.tofile(f)
            raws.append(f)
            out["raw%d" % t] = v
        uvf = os.path.join(td, "ts.uvf")
        subprocess.check_call([O.ref_driver_path(), "writeuvfts", uvf, str(VOL[0]), str(VOL[1]), str(VOL[2]), "u8", str(BRICK), str(OVERLAP)] + raws)
        out["uvf_gz"] = np.frombuffer(gzip.compress(open(uvf, "rb").read(), mtime=0), np.uint8)
        for t in range(len(raws)):
            dump = os.path.join(td, "d.bin")
            subprocess.check_call([O.ref_driver_path(), "dsdump", uvf, str(t), dump])
            steps, sizes, mm, payload = parse(open(dump, "rb").read())
            assert steps == len(raws)
            out["t%d/sizes" % t], out["t%d/minmax" % t], out["t%d/payload" % t] = sizes, mm, payload
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "timesteps_case.npz"), **out)
    print(os.path.getsize(os.path.join(ROOT, "tests", "golden", "timesteps_case.npz")))


if __name__ == "__main__":
    main()
