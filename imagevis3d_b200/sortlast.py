"""Sort-last rendering across GPUs (new in this build; the reference is single-GPU).

The brick grid is split into convex z-slabs, one per rank.  Every rank renders its slab over
the whole image (rays clipped to the slab by tvk_raycast_params.box_min/box_max), which gives a
premultiplied RGBA32F partial image; the partial images are then combined with ordered "under"
compositing (Tuvok/Shaders/Compositing.glsl:33-38 for premultiplied input) by DIRECT SEND:
rank g owns scan-line stripe g, one all_to_all moves every foreign stripe to its owner (one hop
over NVSwitch), the owner blends its N stripes front to back with the compositor kernel and the
finished stripes are gathered.  Bytes on the wire per rank: (N-1)/N * W*H*16 B each way.

exchange_and_composite() is backend agnostic (nccl on GPUs, gloo in the CPU tests): the blend itself is
passed in as a callable.  FusedExchange is the B200 path: the same direct send, but the ray-march kernel
itself stores each finished pixel into the owner's buffer over NVLink peer memory and the compositor waits
on arrival flags -- no collective call on the data path (csrc/tvk_exchange.cu).
"""
import math

import numpy as np


def slab_partition(global_dims, world, halo=64):
    """Splits [0, Gz) into `world` z-slabs.  Each rank also keeps `halo` voxels of its neighbours so
    that trilinear/gradient taps and the coarser LODs (up to log2(halo)-1) see real data at the cut.
    Returns one dict per rank: z0,z1 = the slab proper, zs,ze = slab + halo, local_dims."""
    gx, gy, gz = [int(v) for v in global_dims]
    out = []
    for r in range(world):
        z0, z1 = (gz * r) // world, (gz * (r + 1)) // world
        zs, ze = max(0, z0 - halo), min(gz, z1 + halo)
        out.append(dict(z0=z0, z1=z1, zs=zs, ze=ze, local_dims=(gx, gy, ze - zs), global_dims=(gx, gy, gz)))
    return out


def depth_order(eye_to_model_global, parts):
    """Front-to-back order of the slabs for the eye of this view.  The slabs are stacked along one
    axis, so sorting them by the distance of their centre plane from the eye's z is exact: a ray
    travels monotonically in z and meets the slabs of its side in exactly this order."""
    m = np.asarray(eye_to_model_global, np.float64).reshape(4, 4)
    eye_z = float((np.array([0, 0, 0, 1.0]) @ m)[2])            # eye in global normalised coords
    gz = float(parts[0]["global_dims"][2])
    centres = [0.5 * (p["z0"] + p["z1"]) / gz for p in parts]
    return sorted(range(len(parts)), key=lambda r: (abs(centres[r] - eye_z), r))


def stripe_rows(height, world):
    return int(math.ceil(height / float(world)))


def exchange_and_composite(partial, order, composite, world, rank):
    """partial: torch tensor [H, W, 4] float32 (premultiplied) on this rank's device.
    composite: callable(list of [rows, W, 4] tensors, front to back) -> [rows, W, 4] tensor.
    Returns the [H, W, 4] image on every rank (stripes are all-gathered)."""
    import torch
    import torch.distributed as dist
    h, w, _ = partial.shape
    rows = stripe_rows(h, world)
    hp = rows * world
    if hp != h:
        padded = torch.zeros((hp, w, 4), dtype=partial.dtype, device=partial.device)
        padded[:h] = partial
    else:
        padded = partial.contiguous()
    if world == 1:
        return composite([padded])[:h]
    send = padded.view(world, rows, w, 4)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)                         # stripe `rank` of every rank's partial
    mine = composite([recv[k] for k in order]).contiguous()
    full = torch.empty((world, rows, w, 4), dtype=partial.dtype, device=partial.device)
    dist.all_gather(list(full.unbind(0)), mine)               # (gloo has no all_gather_into_tensor)
    return full.view(hp, w, 4)[:h]


class FusedExchange:
    """Direct-send compositing fused into the kernels over CUDA-IPC peer memory (one process per GPU).
    `dist` is torch.distributed (only used once, to swap the 64-byte IPC handles)."""

    def __init__(self, tvk, world, rank, width, height, dist=None):
        self.tvk, self.world, self.rank, self.frame = tvk, world, rank, 0
        handle = tvk.exchange_create(world, rank, width, height)
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, handle)
            tvk.exchange_open_peers(handles)
            dist.barrier()                       # every rank has mapped every region before the first peer store

    def render(self, params, order):
        """march + send, then blend this rank's stripe; returns the device pointer of the finished H x W
        image (complete on rank 0 once the stream has passed this call; other ranks hold their own stripe)."""
        self.frame += 1
        self.tvk.exchange_begin_frame(self.frame)
        self.tvk.raycast(params)
        self.tvk.exchange_composite(order, self.frame)
        return self.tvk.exchange_final(self.frame)
