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


def _part(lo, hi, gdims, halo):
    s = [max(0, lo[a] - halo) for a in range(3)]
    e = [min(gdims[a], hi[a] + halo) for a in range(3)]
    return dict(lo=tuple(lo), hi=tuple(hi), s=tuple(s), e=tuple(e), local_dims=tuple(e[a] - s[a] for a in range(3)),
                global_dims=tuple(gdims), z0=lo[2], z1=hi[2], zs=s[2], ze=e[2])


def slab_partition(global_dims, world, halo=64):
    """Splits [0, Gz) into `world` z-slabs.  Each rank also keeps `halo` voxels of its neighbours so
    that trilinear/gradient taps and the coarser LODs (up to log2(halo)-1) see real data at the cut.
    Returns one dict per rank: lo/hi = the slab proper, s/e = slab + halo, local_dims (z0,z1,zs,ze: the z range)."""
    gx, gy, gz = [int(v) for v in global_dims]
    return [_part((0, 0, (gz * r) // world), (gx, gy, (gz * (r + 1)) // world), (gx, gy, gz), halo) for r in range(world)]


def block_grid(global_dims, world, prefer=(1, 0, 2)):
    """(px, py, pz) with px*py*pz == world: the longest block edge is halved until the factors are used up, so the
    blocks stay as cubic as the volume allows (8 -> 2x2x2 for a cube).  Ties go to the axes in `prefer` order: y first,
    the renderer's up axis -- interactive views orbit about it, and the two halves of a cut across y stay equidistant from
    the eye at every azimuth, so their cost (LOD, screen footprint) stays balanced; a cut across x or z puts one half in
    front of the other for most of the orbit (measured on 2 x B200, 4096^3: the front half cost up to 2.7 x the back one)."""
    grid, n, dims = [1, 1, 1], int(world), [float(v) for v in global_dims]
    f = 2
    factors = []
    while n > 1:
        while n % f == 0:
            factors.append(f)
            n //= f
        f += 1
    for f in sorted(factors, reverse=True):
        a = max(range(3), key=lambda k: (dims[k] / grid[k], -prefer.index(k)))
        grid[a] *= f
    return tuple(grid)


def block_partition(global_dims, world, halo=64, align=64):
    """Convex boxes on a regular px x py x pz grid, one per rank (rank = ix + px*(iy + py*iz)), cut at multiples of
    `align` voxels.  Compared with z-slabs a ray meets about world^(1/3) boxes instead of `world`, so the per-ray
    set-up work a rank repeats for rays that only graze its part shrinks accordingly.  Same dict as slab_partition."""
    g = [int(v) for v in global_dims]
    px, py, pz = block_grid(g, world)
    cuts = []
    for a, p in enumerate((px, py, pz)):
        c = [int(round(g[a] * i / p / align)) * align for i in range(p + 1)]
        c[0], c[-1] = 0, g[a]
        cuts.append(c)
    parts = []
    for iz in range(pz):
        for iy in range(py):
            for ix in range(px):
                parts.append(_part((cuts[0][ix], cuts[1][iy], cuts[2][iz]), (cuts[0][ix + 1], cuts[1][iy + 1], cuts[2][iz + 1]), g, halo))
    return parts


def depth_order(eye_to_model_global, parts):
    """Front-to-back order of the parts (slabs or grid boxes) for the eye of this view.  Along every axis a ray
    moves monotonically, so it meets the parts in non-decreasing distance of their [lo, hi) interval from the eye's
    coordinate on EACH axis; sorting by the tuple of the three per-axis distances is a linear extension of that
    product order, hence a correct visibility order for every ray at once (ties are parts no ray can cross both)."""
    m = np.asarray(eye_to_model_global, np.float64).reshape(4, 4)
    eye = (np.array([0, 0, 0, 1.0]) @ m)[:3]                    # eye in global normalised coords
    g = [float(v) for v in parts[0]["global_dims"]]

    def key(r):
        p = parts[r]
        k = []
        for a in range(3):
            lo, hi = p["lo"][a] / g[a], p["hi"][a] / g[a]
            k.append(max(lo - eye[a], 0.0, eye[a] - hi))
        return (k[0], k[1], k[2], r)
    return sorted(range(len(parts)), key=key)


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


def antipodal_boxes(global_dims, ranks, boxes_per_rank=2, halo=64, align=64):
    """Over-decomposition for view-independent load balance: the volume is cut into ranks * boxes_per_rank convex boxes
    (block_partition) and every rank gets boxes in ANTIPODAL PAIRS -- a box and its reflection through the volume's
    vertical centre line (x and z mirrored).  Whatever the azimuth, one box of a pair lies on the near side and the other on
    the far side, so the pair's cost (finer LODs and a larger footprint near the eye) stays about constant, where a single
    box per rank swings by 2-3 x over an orbit.  Along y (the up axis, cut first and most often) the partner is the mirrored
    layer when there are two layers and the layer HALF THE STACK AWAY when there are four or more: an outer layer is
    paired with an inner one, because scanned objects sit in the middle of their volume and the outer layers are mostly
    empty (4096^3 sphere on 8 GPUs, 2x4x2 boxes: mirrored pairs gave 2.2 ms on four GPUs and 4.6 ms on the others).
    -> (parts, owner): parts as block_partition gives them, owner[i] = the rank that renders box i."""
    v = int(ranks) * int(boxes_per_rank)
    assert boxes_per_rank % 2 == 0, "boxes come in antipodal pairs"
    px, py, pz = block_grid(global_dims, v)
    parts = block_partition(global_dims, v, halo=halo, align=align)

    def partner_y(iy):
        return (iy + py // 2) % py if (py >= 4 and py % 2 == 0) else py - 1 - iy
    owner, nxt = [-1] * v, 0
    for iz in range(pz):
        for iy in range(py):
            for ix in range(px):
                i = ix + px * (iy + py * iz)
                if owner[i] >= 0:
                    continue
                j = (px - 1 - ix) + px * (partner_y(iy) + py * (pz - 1 - iz))
                owner[i] = owner[j] = nxt % ranks
                nxt += 1
    return parts, owner


class MultiBoxExchange:
    """FusedExchange for ranks that render several boxes: every box is an exchange MEMBER (its own context, pool, partial
    image and scan-line stripe); members of one process address each other's regions directly, remote members through
    CUDA IPC mappings that the process opens once and shares among its members."""

    def __init__(self, members, member_ids, total_members, width, height, dist=None):
        """members: this process's TvkContexts, member_ids: their exchange ranks (ascending), total_members: all of them"""
        self.members, self.ids, self.total, self.frame = list(members), list(member_ids), int(total_members), 0
        handles = {m: t.exchange_create(self.total, m, width, height) for t, m in zip(self.members, self.ids)}
        table = [None] * self.total
        if dist is not None and dist.get_world_size() > 1:
            gathered = [None] * dist.get_world_size()
            dist.all_gather_object(gathered, handles)
            for g in gathered:
                for m, h in g.items():
                    table[m] = h
        else:
            for m, h in handles.items():
                table[m] = h
        local = {m: t.exchange_local_base() for t, m in zip(self.members, self.ids)}
        first = self.members[0]
        first.exchange_connect(handles=[None if m in local else table[m] for m in range(self.total)],
                               bases=[local.get(m) for m in range(self.total)])
        bases = [first.exchange_peer_base(m) for m in range(self.total)]
        for t in self.members[1:]:
            t.exchange_connect(bases=bases)
        if dist is not None and dist.get_world_size() > 1:
            dist.barrier()                       # every process has mapped every region before the first peer store

    def render(self, params_per_member, order):
        """march + send every local box, then blend every local stripe; -> device pointer of member 0's final image (or None
        where member 0 lives in another process)"""
        self.frame += 1
        for t, p in zip(self.members, params_per_member):
            t.exchange_begin_frame(self.frame)
            t.raycast(p)
        for t in self.members:
            t.exchange_composite(order, self.frame)
        return self.members[0].exchange_final(self.frame) if self.ids[0] == 0 else None
