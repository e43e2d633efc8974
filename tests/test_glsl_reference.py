"""CPU: the restated image oracle (oracle/tvk_oracle_render.c) against THE REFERENCE'S OWN SHADER TEXT executed on the CPU
(oracle/glslref.py: Tuvok/Shaders/*.glsl and the GLSL GLVolumePool / GLHashTable generate, compiled by g++ against a GLSL
value-type shim).  This is what pins the image oracle to something the builder did not write.  Needs /root/reference
(skipped on the GPU box, where tests/test_gpu_glsl_golden.py checks the CUDA path against the committed frames instead)."""
import ctypes as C

import numpy as np
import pytest

from oracle import glslref as G
from oracle import host as H
from tests import glsl_scenes as S

pytestmark = pytest.mark.skipif(not G.available(), reason="/root/reference (the shader sources) is not mounted")


@pytest.mark.parametrize("name", sorted(S.SCENES))
def test_oracle_equals_reference_shaders(name):
    # one OpenMP thread: the oracle then visits the pixels in the harness's order (row major), so that colliding
    # requests probe the table in the same order and the TABLES can be compared entry by entry, not only as sets
    gomp = C.CDLL("libgomp.so.1")
    threads = gomp.omp_get_max_threads()
    gomp.omp_set_num_threads(1)
    try:
        a = S.render_progressive(H.raycast, S.build(name))
    finally:
        gomp.omp_set_num_threads(threads)
    b = S.render_progressive(G.raycast, S.build(name))
    # integer side: every sub-frame's request TABLE is the same array (same bricks in the same table entries)
    assert a["subframes"] == b["subframes"], (a["subframes"], b["subframes"])
    for k, (ha, hb) in enumerate(zip(a["hashes"], b["hashes"])):
        assert (ha == hb).all(), "request table of sub-frame %d differs" % k
        assert {tuple(r) for r in a["requests"][k]} == {tuple(r) for r in b["requests"][k]}
    assert a["subframes"] > 1 and sum(len(r) for r in a["requests"]) > 0, "no paging: the scene would be vacuous"
    assert b["image"].max() > 0.05
    # float side: the oracle fuses multiply-adds where a GPU compiler would (explicit fmaf), the shader text compiled
    # here does not contract anything, so the two differ by rounding only
    d = float(np.abs(a["image"] - b["image"]).max())
    psnr = H.psnr(a["image"], b["image"])
    print("%s: %d sub-frames, max |diff| %.3g, PSNR %.1f dB" % (name, a["subframes"], d, psnr))
    if S.SCENES[name]["mode"] == 2:
        # a hit / miss flip on a silhouette pixel is a full-scale difference: bound their number
        dd = np.abs(a["image"] - b["image"]).max(axis=2)
        assert (dd > 1e-3).sum() <= 3, int((dd > 1e-3).sum())
    else:
        # ... except on the few rays where that rounding moves a step count (int(ceil(.)) of a brick's length) by one sample
        dd = np.abs(a["image"] - b["image"]).max(axis=2)
        print("    pixels > 1e-4: %d of %d" % (int((dd > 1e-4).sum()), dd.size))
        assert d <= 2.0 / 255.0 and psnr >= 80.0, (d, psnr)
        assert (dd > 1e-4).sum() <= max(2, dd.size // 500), int((dd > 1e-4).sum())
    # resume state: every ray finished (w == 1000; iso hits keep w = 500 + x), at the same place
    wa, wb = a["pos"][..., 3], b["pos"][..., 3]
    assert (np.floor(wa) == np.floor(wb)).all() and ((wa == 1000.0) | (np.floor(wa) == 501.0)).all()


def test_generated_pool_fragment_is_the_reference_text():
    """The GLSL that runs is what the reference's GetShaderFragment prints for this pool -- spot-check the baked constants."""
    sc = S.build("u8_1d")
    pool_src, hash_src = G.shader_text(sc["pool"], sc["params"])
    assert "bool GetBrick(in vec3 normEntryCoords, inout uint iLOD" in pool_src
    assert "#define poolCapacity ivec3(8, 8, 8)" in pool_src and "#define volumeSize vec3(150, 131, 97)" in pool_src
    assert "if(brickInfo == BI_MISSING && iStartLOD+2 == iLOD)" in pool_src          # SkipTwoLevels, the reference default
    assert "return uint(serializedValue % 509);" in hash_src and "imageAtomicCompSwap(hashTable" in hash_src
    pool_src2, hash_src2 = G.shader_text(S.build("u8_1d_reqall")["pool"], S.build("u8_1d_reqall")["params"])
    assert "uimage2D hashTable" in hash_src2 and "if(brickInfo == BI_MISSING) ReportMissingBrick(brickCoords);" in pool_src2
