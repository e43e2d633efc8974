"""CPU: the quantisation restatement (oracle/tvk_oracle_quantize.c) against
(a) the known answers of the reference's own tests (Tuvok/IO/test/quantize.h:104-262),
(b) the committed golden vectors generated from the reference templates (tests/golden/make_golden_quantize.py),
(c) the reference itself where oracle/_ref/ref_driver exists (build container only)."""
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "quantize_cases.npz")


def quant_cases():
    """name -> (array, to8bit); deterministic, shared with the golden generator and the GPU tests."""
    rng = np.random.default_rng(11)
    c = {}
    c["u8_noise"] = (rng.integers(0, 256, 5000).astype(np.uint8), False)
    c["i8_noise"] = (rng.integers(-128, 128, 5000).astype(np.int8), False)
    c["u16_12bit"] = (rng.integers(0, 4096, 20000).astype(np.uint16), False)
    c["u16_full"] = (rng.integers(0, 65536, 20000).astype(np.uint16), False)
    c["u16_biased"] = (rng.integers(30000, 36000, 20000).astype(np.uint16), False)
    c["u16_fewvalues"] = ((rng.integers(0, 100, 20000) * 37).astype(np.uint16), False)
    c["u16_range200"] = (rng.integers(50000, 50200, 20000).astype(np.uint16), False)
    c["i16_ct"] = (rng.integers(-1024, 3071, 20000).astype(np.int16), False)
    c["i16_full"] = (rng.integers(-32768, 32768, 20000).astype(np.int16), False)
    c["i32_wide"] = (rng.integers(-10 ** 9, 10 ** 9, 20000).astype(np.int32), False)
    c["i32_narrow"] = (rng.integers(-500, 700, 20000).astype(np.int32), False)
    c["u32_wide"] = (rng.integers(0, 2 ** 32, 20000).astype(np.uint32), False)
    c["u32_few"] = ((rng.integers(0, 200, 20000) * 1000003).astype(np.uint32), False)
    c["f32_smooth"] = ((rng.random(20000) * 3.5 - 1.25).astype(np.float32), False)
    c["f32_binned"] = ((rng.integers(0, 3000, 20000) * 0.125 - 7).astype(np.float32), False)
    c["f32_binned_small"] = ((rng.integers(0, 200, 20000) * 0.5).astype(np.float32), False)
    c["f64_smooth"] = (rng.random(20000) * 1e4 - 17.0, False)
    c["i64_wide"] = (rng.integers(-2 ** 40, 2 ** 40, 20000).astype(np.int64), False)
    c["u64_wide"] = (rng.integers(0, 2 ** 50, 20000).astype(np.uint64), False)
    for k in ("u16_full", "i16_ct", "i32_wide", "u32_wide", "f32_smooth", "f64_smooth", "u16_12bit", "i8_noise", "u8_noise"):
        c[k + "_to8"] = (c[k][0], True)
    return c


def _same(a, b):
    assert a["out_bits"] == b["out_bits"] and a["generated"] == b["generated"]
    assert a["data"].dtype == b["data"].dtype and (a["data"] == b["data"]).all()
    assert (a["hist"] is None) == (b["hist"] is None)
    if a["hist"] is not None:
        assert a["hist"].size == b["hist"].size and (a["hist"] == b["hist"]).all()


@pytest.mark.parametrize("dt", [np.int8, np.int16, np.float32, np.float64, np.uint8, np.uint16, np.int32, np.uint32])
def test_reference_kat_16bit(dt):
    """verify_type<T>: 100 consecutive values from -64 (signed) or 0 (unsigned)."""
    sgn = np.issubdtype(dt, np.signedinteger) or np.issubdtype(dt, np.floating)
    start = -64 if sgn else 0
    data = np.arange(start, start + 100).astype(dt)
    r = O.quantize(data)
    out = r["data"]
    if dt in (np.int8, np.uint8):
        want = (np.arange(start, start + 100) + (128 if sgn else 0)).astype(np.uint8)
        assert (out == want).all()
        low = start + 128 if sgn else 0
        h = np.zeros(256, np.uint64)
        h[low:low + 100] = 1
        assert (r["hist"] == h).all()
    elif not np.issubdtype(dt, np.floating):
        assert (out.astype(np.int64) == np.arange(100)).all()          # bias to 0, factor min(65535/100, 1) = 1
        h = r["hist"] if r["hist"] is not None else np.bincount(out, minlength=4096)
        assert (h[:100] == 1).all() and (h[100:] == 0).all()


@pytest.mark.parametrize("dt", [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32])
def test_reference_kat_8bit(dt):
    """verify_8b_type<T>."""
    sgn = np.issubdtype(dt, np.signedinteger)
    start = -64 if sgn else 0
    data = np.arange(start, start + 100).astype(dt)
    r = O.quantize(data, to8bit=True)
    out = r["data"]
    assert r["out_bits"] == 8
    if dt == np.int8:
        assert (out == (np.arange(start, start + 100) + 128).astype(np.uint8)).all()
    else:
        assert (out.astype(np.int64) == np.arange(100)).all()
    if dt in (np.int8, np.uint8):
        low = start + 128 if sgn else 0
        assert (r["hist"][low:low + 100] == 1).all() and r["hist"].sum() == 100
    else:
        assert (r["hist"][:100] == 1).all() and (r["hist"][100:] == 0).all()


def test_against_golden():
    g = np.load(GOLD)
    for name, (data, to8) in quant_cases().items():
        r = O.quantize(data, to8)
        assert r["out_bits"] == int(g[name + "/bits"]) and r["generated"] == bool(g[name + "/generated"]), name
        if r["generated"]:
            assert (r["data"] == g[name + "/data"]).all(), name
        hist = g[name + "/hist"]
        assert (r["hist"] is None) == (hist.size == 0), name
        if r["hist"] is not None:
            assert r["hist"].size == hist.size and (r["hist"] == hist).all(), name


@pytest.mark.skipif(O.ref_driver_path() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_against_reference_build():
    for name, (data, to8) in quant_cases().items():
        _same(O.quantize(data, to8), O.ref_quantize(data, to8))
