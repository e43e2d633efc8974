"""GPU: the CUDA ray-caster against frames rendered by THE REFERENCE'S OWN SHADERS (tests/golden/glsl_frames.npz, written in
the build container by tests/golden/make_golden_glsl.py through oracle/glslref.py).  Nothing here reads /root/reference.

Per scene: the view is painted to convergence from a cold pool through the C ABI; after every sub-frame the brick requests
read back from the device must be the SET the reference's shaders produced in that sub-frame (bit-exact side), the pool is
paged with exactly those requests, and the converged image must be within the north-star tolerance of the golden frame
(PSNR >= 45 dB, max channel error <= 2/255)."""
import os

import numpy as np
import pytest

from oracle import host as H
from tests import glsl_scenes as S
from tests.util import GOLDEN, sync_pool, to_cuda_params

pytestmark = pytest.mark.gpu

PSNR_MIN = 45.0
MAX_ERR = 2.0 / 255.0


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "glsl_frames.npz"))


@pytest.mark.parametrize("sampler", [0, 3], ids=["exact", "gather"])
@pytest.mark.parametrize("name", sorted(S.SCENES))
def test_cuda_matches_reference_shader_frames(tvk, golden, name, sampler):
    sc = S.build(name)
    p, pool, tf = sc["params"], sc["pool"], sc["tf"]
    want = golden[name + "/image"]
    nsub = int(golden[name + "/subframes"])
    mirror = set()
    tvk.pool_create(pool.vol, pool.max_brick, pool.overlap, pool.bpv, pool.capacity, pool.lod_count, sampler)
    sync_pool(tvk, pool, mirror)
    if tf is None:
        tf = H.tf_bytes(S._tf(256))                # the isosurface kernels read no table; the context wants one bound
    (tvk.set_tf1d if tf.ndim == 2 else tvk.set_tf2d)(tf)
    p.new_view = 1
    for sub in range(nsub):
        tvk.raycast(to_cuda_params(p))
        p.new_view = 0
        got = {tuple(int(v) for v in r) for r in tvk.read_requests()}
        ref = golden["%s/req%d" % (name, sub)].astype(np.uint32)
        assert got == {tuple(int(v) for v in r) for r in ref}, "sub-frame %d: request set differs from the reference shaders'" % sub
        if len(ref):
            pool.upload_requests(ref)
            sync_pool(tvk, pool, mirror)
    assert len(ref) == 0, "golden sequence must end converged"
    img = tvk.read_image()
    d = np.abs(img.astype(np.float64) - want.astype(np.float64))
    psnr = H.psnr(img, want)
    print("%s: PSNR %.1f dB, max err %.3g (%.3f/255), identical pixels %.4f" % (name, psnr, d.max(), d.max() * 255, float((img == want).all(axis=2).mean())))
    assert want.max() > 0.05
    if S.SCENES[name]["mode"] == 2:
        # a hit / miss flip on a silhouette pixel is a full-scale difference: bound their number instead
        bad = int((d.max(axis=2) > MAX_ERR).sum())
        assert bad <= max(4, int(0.002 * d.shape[0] * d.shape[1])), bad
    else:
        assert psnr >= PSNR_MIN and d.max() <= MAX_ERR, (psnr, d.max())
