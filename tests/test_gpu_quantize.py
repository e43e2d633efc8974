"""GPU: tvk_quantize (B5) against the golden vectors generated from the reference's own templates and
against the CPU oracle, bit-exact, over every input type and every branch of the dispatcher."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.test_oracle_quantize import GOLD, quant_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(quant_cases().keys()))
def test_quantize_matches_reference_golden(tvk, name):
    data, to8 = quant_cases()[name]
    g = np.load(GOLD)
    r = tvk.quantize(data, to8)
    assert r["out_bits"] == int(g[name + "/bits"]) and r["generated"] == bool(g[name + "/generated"])
    if r["generated"]:
        assert r["data"].dtype == g[name + "/data"].dtype and (r["data"] == g[name + "/data"]).all()
    hist = g[name + "/hist"]
    assert (r["hist"] is None) == (hist.size == 0)
    if r["hist"] is not None:
        assert r["hist"].size == hist.size and (r["hist"] == hist).all()


@pytest.mark.parametrize("dt,lo,hi,n", [(np.int16, -2000, 30000, 1 << 20), (np.float32, None, None, (1 << 20) + 13),
                                        (np.uint16, 0, 65536, (1 << 21) + 5), (np.uint32, 10, 4000000000, 300007),
                                        (np.uint16, 100, 330, 1 << 19), (np.int8, -128, 128, 100003)])
def test_quantize_matches_oracle_large(tvk, dt, lo, hi, n):
    rng = np.random.default_rng(n)
    data = (rng.standard_normal(n) * 40 + 3).astype(np.float32) if lo is None else rng.integers(lo, hi, n).astype(dt)
    for to8 in (False, True):
        a, b = tvk.quantize(data, to8), O.quantize(data, to8)
        assert a["out_bits"] == b["out_bits"] and a["generated"] == b["generated"]
        if a["generated"]:
            assert (a["data"] == b["data"]).all()
        assert (a["hist"] is None) == (b["hist"] is None)
        if a["hist"] is not None:
            assert (a["hist"] == b["hist"]).all()


def test_quantize_then_build(tvk):
    """The B5 -> B1..B4 hand-over on the device: signed 16-bit CT-like data, biased to 0, then bricked."""
    import torch
    vol = (70, 66, 40)
    rng = np.random.default_rng(3)
    raw = rng.integers(-1024, 3000, size=vol[::-1]).astype(np.int16)
    src = torch.from_numpy(raw).cuda()
    dst = torch.empty(raw.size * 2, dtype=torch.uint8, device="cuda")
    r = tvk.quantize(src_dptr=src.data_ptr(), dtype=np.int16, count=raw.size, dst_dptr=dst.data_ptr())
    assert r["generated"] and r["out_bits"] == 16
    q = O.quantize(raw)
    tvk.build_octree(raw_dptr=dst.data_ptr(), volume_size=vol, bpv=2, brick=32, overlap=2)
    oc = O.build_octree(q["data"].reshape(raw.shape), 32, 2)
    from tests.util import cuda_store_tight
    store, _ = cuda_store_tight(tvk)
    assert (store == oc["store"]).all() and (tvk.build_minmax() == oc["minmax"]).all()
    tvk.build_release()
