#!/usr/bin/env python
"""Generates tests/golden/quantize_cases.npz from the REFERENCE's own templates (oracle/_ref/ref_driver
quantize -> Quantize<T,U>, BinningQuantize<T,U>, Process8Bits, QuantizeTo8Bit); run in the build
container, where /root/reference exists.  The inputs are regenerated from a fixed seed by
tests/test_oracle_quantize.quant_cases(), so only the outputs are stored."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O                      # noqa: E402
from tests.test_oracle_quantize import quant_cases  # noqa: E402

out = {}
for name, (data, to8) in quant_cases().items():
    r = O.ref_quantize(data, to8)
    out[name + "/bits"] = np.uint32(r["out_bits"])
    out[name + "/generated"] = np.uint32(r["generated"])
    out[name + "/data"] = r["data"] if r["generated"] else np.zeros(0, np.uint8)
    out[name + "/hist"] = r["hist"] if r["hist"] is not None else np.zeros(0, np.uint64)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "quantize_cases.npz"), **out)
print("wrote", len(quant_cases()), "cases")
