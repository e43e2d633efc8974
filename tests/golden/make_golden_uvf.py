#!/usr/bin/env python
"""Writes tests/golden/ref_*.uvf with the REFERENCE (oracle/_ref/ref_driver writeuvf: TOCBlock::
FlatDataToBrickedLOD, Histogram1D/2DDataBlock::Compute, MaxMinDataBlock, KeyValuePairDataBlock, UVF::Create /
Close with MD5 -- the assembly of RAWConverter::ConvertRAWDataset).  Build container only."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O                       # noqa: E402
from tests.test_uvf import CASES, GOLD, case_volume  # noqa: E402

for name, (vol, dt, brick, overlap, comp, level) in CASES.items():
    raw = case_volume(name)
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "v.raw")
        raw.tofile(f)
        subprocess.check_call([O.ref_driver_path(), "writeuvf", f, str(vol[0]), str(vol[1]), str(vol[2]),
                               "u16" if raw.dtype.itemsize == 2 else "u8", str(brick), str(overlap), "0", str(comp), str(level),
                               os.path.join(td, "o.uvf")])
        import gzip
        with gzip.GzipFile(os.path.join(GOLD, name + ".uvf.gz"), "wb", mtime=0) as g:
            g.write(open(os.path.join(td, "o.uvf"), "rb").read())
    print(name, os.path.getsize(os.path.join(GOLD, name + ".uvf.gz")))
