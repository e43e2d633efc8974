"""GPU: sort-last z-slab rendering (clip box + slab parameter mapping + ordered compositor) on ONE
GPU, the ranks emulated one after the other.  (a) CUDA slabs + CUDA compositor == oracle slabs +
oracle compositor within the image tolerance; (b) the composited slabs reproduce the single-volume
render (different sample phase per slab, hence a looser bound)."""
import ctypes as C

import numpy as np
import pytest

from imagevis3d_b200 import sortlast as SL
from oracle import host as H
from oracle import oracle as O
from tests.util import image_errors, lockstep_render

pytestmark = pytest.mark.gpu


def _tf():
    t = H.tf1d_std(256, 0.3, 0.3)
    t[:, 3] *= 0.12
    t[:, 1] *= 0.7
    return t


@pytest.mark.parametrize("angle", [0.3, 1.9, 3.5])
def test_slabs_composite_like_one_volume(tvk, angle):
    import torch
    G = (64, 64, 192)
    vol = H.synth_blobs(G, np.uint8, seed=21)
    parts = SL.slab_partition(G, 3, halo=32)
    tf = _tf()
    lo, hi = H.tf_nonzero_limits(tf)
    cam = H.Camera()
    cam.rotation = H.rotation_y(angle) @ H.rotation_x(0.25)
    W, Hh = 160, 120
    # single volume reference (oracle)
    oc = O.build_octree(vol, 20, 2)
    pool = H.OraclePool(oc, (6, 6, 14))
    pool.upload_first_brick()
    pool.recompute_visibility(0, float(lo), float(hi))
    p_global = H.raycast_params(cam, W, Hh, G)
    whole, _, _, _ = H.render_converged(p_global, pool, H.tf_bytes(tf))
    order = SL.depth_order(np.array(p_global.eye_to_model[:]).reshape(4, 4), parts)
    cuda_parts, orc_parts = {}, {}
    for r, part in enumerate(parts):
        local = np.ascontiguousarray(vol[part["zs"]:part["ze"]])
        ocl = O.build_octree(local, 20, 2)
        pl = H.OraclePool(ocl, (6, 6, 10))
        pl.upload_first_brick()
        pl.recompute_visibility(0, float(lo), float(hi))
        p = H.raycast_params(cam, W, Hh, G, slab=part)
        res = lockstep_render(tvk, p, pl, H.tf_bytes(tf))
        psnr, err = image_errors(res["cuda"], res["oracle"])
        assert psnr >= 45.0 and err <= 2.0 / 255.0, (r, psnr, err)
        cuda_parts[r], orc_parts[r] = res["cuda"], res["oracle"]
    # oracle composite
    pix = W * Hh
    plist = [np.ascontiguousarray(orc_parts[r]) for r in order]
    want = np.zeros((Hh, W, 4), np.float32)
    arr = (C.c_void_p * len(plist))(*[a.ctypes.data for a in plist])
    O.lib().orc_composite_ordered(arr, C.c_uint32(len(plist)), C.c_uint64(pix), want.ctypes.data_as(C.c_void_p))
    # CUDA composite
    dev = [torch.from_numpy(np.ascontiguousarray(cuda_parts[r])).cuda() for r in order]
    out = torch.zeros((Hh, W, 4), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    tvk.composite_ordered([d.data_ptr() for d in dev], pix, out.data_ptr())
    tvk.synchronize()
    got = out.cpu().numpy()
    psnr, err = image_errors(got, want)
    assert want.max() > 0.05 and psnr >= 45.0 and err <= 2.0 / 255.0, (psnr, err)
    # against the un-partitioned render: same integral, different sample phase inside each slab
    psnr, err = image_errors(got, whole)
    assert psnr >= 38.0 and err <= 12.0 / 255.0, (psnr, err)


def test_clip_box_validation(tvk):
    from imagevis3d_b200._lib import TvkError
    from tests.util import to_cuda_params
    p = to_cuda_params(H.raycast_params(H.Camera(), 32, 32, (64, 64, 64)))
    p.box_min[2], p.box_max[2] = 0.6, 0.4
    with pytest.raises(TvkError):
        tvk.raycast(p)


@pytest.mark.parametrize("world,angle", [(4, 0.4), (8, 2.3)])
def test_block_partition_composites_like_one_volume(tvk, world, angle):
    """regular-grid boxes (sortlast.block_partition) instead of z-slabs: per-box renders composited in
    depth_order reproduce the single-volume render (oracle), as the slab test above does."""
    import torch
    G = (128, 128, 128)
    vol = H.synth_blobs(G, np.uint8, seed=29)
    parts = SL.block_partition(G, world, halo=32, align=32)
    tf = _tf()
    lo, hi = H.tf_nonzero_limits(tf)
    cam = H.Camera()
    cam.rotation = H.rotation_y(angle) @ H.rotation_x(0.35)
    W, Hh = 160, 120
    oc = O.build_octree(vol, 20, 2)
    pool = H.OraclePool(oc, (9, 9, 9))
    pool.upload_first_brick()
    pool.recompute_visibility(0, float(lo), float(hi))
    p_global = H.raycast_params(cam, W, Hh, G)
    whole, _, _, _ = H.render_converged(p_global, pool, H.tf_bytes(tf))
    order = SL.depth_order(np.array(p_global.eye_to_model[:]).reshape(4, 4), parts)
    partials = {}
    for r, part in enumerate(parts):
        local = np.ascontiguousarray(vol[part["s"][2]:part["e"][2], part["s"][1]:part["e"][1], part["s"][0]:part["e"][0]])
        ocl = O.build_octree(local, 20, 2)
        pl = H.OraclePool(ocl, (7, 7, 7))
        pl.upload_first_brick()
        pl.recompute_visibility(0, float(lo), float(hi))
        res = lockstep_render(tvk, H.raycast_params(cam, W, Hh, G, slab=part), pl, H.tf_bytes(tf))
        psnr, err = image_errors(res["cuda"], res["oracle"])
        assert psnr >= 45.0 and err <= 2.0 / 255.0, (r, psnr, err)
        partials[r] = res["cuda"]
    dev = [torch.from_numpy(np.ascontiguousarray(partials[r])).cuda() for r in order]
    out = torch.zeros((Hh, W, 4), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    tvk.composite_ordered([d.data_ptr() for d in dev], W * Hh, out.data_ptr())
    tvk.synchronize()
    psnr, err = image_errors(out.cpu().numpy(), whole)
    assert whole.max() > 0.05 and psnr >= 38.0 and err <= 12.0 / 255.0, (psnr, err)


@pytest.mark.parametrize("world,angle", [(3, 0.3), (3, 3.5), (4, 1.9), (16, 0.8)])      # 16 = the most members the flag tables hold
def test_fused_exchange_equals_separate_compositor(world, angle):
    """The fused direct-send path (K1 peer stores + arrival flags + stripe compositor, csrc/tvk_exchange.cu)
    with the ranks emulated inside one process on one GPU: its image must equal, bit for bit, the ordered
    compositor applied to the ranks' partial images -- over several frames (both buffer parities)."""
    import torch
    from imagevis3d_b200.context import TvkContext
    from tests.util import resident_pool, to_cuda_params
    G = (64, 64, 64 * world) if world <= 4 else (48, 48, 40 * world)
    vol = H.synth_blobs(G, np.uint8, seed=23)
    parts = SL.slab_partition(G, world, halo=32)
    tf = _tf()
    lo, hi = H.tf_nonzero_limits(tf)
    W, Hh = 161, 123                                   # height not divisible by the world size
    ctxs = []
    for r, part in enumerate(parts):
        t = TvkContext(0)
        resident_pool(t, np.ascontiguousarray(vol[part["zs"]:part["ze"]]), 20, H.tf_bytes(tf), float(lo), float(hi))
        t.exchange_create(world, r, W, Hh)
        ctxs.append(t)
    bases = [t.exchange_local_base() for t in ctxs]
    for t in ctxs:
        t.exchange_set_peers(bases)
    for frame in range(1, 4):
        cam = H.Camera()
        cam.rotation = H.rotation_y(angle + 0.2 * frame) @ H.rotation_x(0.25)
        pg = H.raycast_params(cam, W, Hh, G)
        order = SL.depth_order(np.array(pg.eye_to_model[:]).reshape(4, 4), parts)
        partial = []
        for r, t in enumerate(ctxs):
            t.exchange_begin_frame(frame)
            t.raycast(to_cuda_params(H.raycast_params(cam, W, Hh, G, slab=parts[r])))
            t.synchronize()
            partial.append(torch.from_numpy(t.read_image()).cuda())
        for t in ctxs:
            t.exchange_composite(order, frame)
        final = ctxs[0].exchange_final(frame)
        for t in ctxs:
            assert not t.exchange_timed_out()
        class _CAI:
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (int(ptr), False), "version": 2}
        got = torch.as_tensor(_CAI(final, (Hh, W, 4)), device="cuda").cpu().numpy()
        want = torch.zeros((Hh, W, 4), dtype=torch.float32, device="cuda")
        ctxs[0].composite_ordered([partial[r].data_ptr() for r in order], W * Hh, want.data_ptr())
        ctxs[0].synchronize()
        want = want.cpu().numpy()
        assert want.max() > 0.05 and np.array_equal(got, want), frame
    for t in ctxs:
        t.close()


@pytest.mark.parametrize("members,mode", [(2, "1d"), (4, "1d_lit"), (8, "1d")])
def test_sort_last_through_the_renderer_api(members, mode):
    """tvkr_sortlast_*: every box is a Renderer (its part of the volume loaded as its own dataset, out of core, paged on
    request); the group's frame -- members emulated on one GPU -- equals the frame ONE renderer paints of the whole volume
    within the image tolerance (boxes resolve their bricks against their own pyramids: the cut faces see the halo)."""
    from imagevis3d_b200.renderer import Renderer, SortLastGroup, RT_HEADLESS
    G = (128, 128, 128)
    vol = H.synth_blobs(G, np.uint8, seed=23)
    tf = _tf()
    cam = H.rotation_y(0.7) @ H.rotation_x(0.3)
    W, Hh = 200, 150

    def setup(r, tf_size):
        t = H.tf1d_std(tf_size, 0.3, 0.3)
        t[:, 3] *= 0.15
        r.set1DTransFloat(t)
        r.setLightingEnabled(mode == "1d_lit")
        r.resize((W, Hh))
        r.setRendererTarget(RT_HEADLESS)
        r.setRotation4x4(cam)

    whole = Renderer(0)
    whole.loadRawVolume(vol, brick_size=32, out_of_core=True)
    setup(whole, whole.datasetInfo().tf_size)
    whole.paintUntilConverged()
    want = whole.frameRGBA8().astype(np.int32)
    vmax = float(whole.datasetInfo().range[1])
    whole.cleanup()

    parts = SL.block_partition(G, members, halo=32, align=32)
    rs = []
    for part in parts:
        r = Renderer(0)
        s, e = part["s"], part["e"]
        r.loadRawVolume(np.ascontiguousarray(vol[s[2]:e[2], s[1]:e[1], s[0]:e[0]]), brick_size=32, out_of_core=True)
        setup(r, 256)
        rs.append(r)
    grp = SortLastGroup(rs, list(range(members)), members, G, parts, vmax)
    assert sorted(grp.depth_order()) == list(range(members))
    for frame in range(2):                                     # both buffer parities
        subs = grp.render()
        assert subs >= members
        got = grp.frame_rgba8().astype(np.int32)
        err = np.abs(got - want)
        psnr = 10 * np.log10(255.0 ** 2 / max(float((err.astype(np.float64) ** 2).mean()), 1e-12))
        print("members %d %s frame %d: PSNR %.1f dB, max err %d/255, lit pixels %d" % (members, mode, frame, psnr, err.max(), int((want[..., 3] > 0).sum())))
        assert want[..., 3].max() > 30 and psnr >= 48.0 and err.max() <= 6, (psnr, err.max())
    for r in rs:
        r.cleanup()
