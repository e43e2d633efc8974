#!/usr/bin/env python
"""bench.py -- ray samples/s and ms/frame of the bricked ray-march hot path.

Workload at N=1 (BASELINE.json configs[1]): synthetic 512^3 uint8 volume (the reference's own
"sphere" generator, UVFReader/DataSource.h:361-372, plus a deterministic ripple so that rays do
not all terminate at the same depth), bricked 64/overlap 2 by the CUDA builder, fully resident
brick pool, 1D transfer function + early ray termination, 1920x1080, one converged frame per
step on an orbiting camera (every step is a new view).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (torchrun, one rank per GPU): sort-last.  Every rank owns one 512^3 z-slab of a
512x512x(512*N) volume (weak scaling), renders its partial image, the partial images are
exchanged as scan-line stripes (all_to_all over NCCL/NVLink), blended in depth order by the
compositor kernel and gathered on rank 0.

--impl reference times the CPU restatement of the reference's GLSL ray-marcher (oracle/, OpenMP
on all host cores) because the reference's GL path cannot run here (no libGL/Mesa/Qt; see
DESIGN.md); it is a reported baseline, not the target.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOL = (512, 512, 512)
WIDTH, HEIGHT = 1920, 1080
BRICK, OVERLAP = 64, 2
BYTES_PER_SAMPLE = 1.2136 * 1 / 2.0     # 1.2136*bpv/(2*sampleRateModifier), uint8, srm 1
BYTES_PER_RAY = 16.0                    # RGBA32F result write
CPU_SAMPLE = (1920, 1080)               # the CPU arm renders full frames (a frame is < 1 s on 16 cores)
CPU_STEPS = 24                          # ~0.47 s per 1080p frame on 16 cores -> ~11 s of CPU work


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


_REAL_STDOUT = None


def emit_line(line):
    """The one JSON line, on the process's real stdout (see the descriptor swap in the multi-rank set-up)."""
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, text.encode())


def bench_tf():
    """TransferFunction1D::SetStdFunction(0.5,0.5) colours with a 10x thinner alpha."""
    from_oracle = False
    n = 256
    t = np.zeros((n, 4), np.float32)
    ic, ig = int((n - 1) * 0.5), int((n - 1) * 0.5)
    start, end = ic - ig // 2, ic + ig // 2
    for i in range(start, end):
        x = np.float32(float(i - ic + ig // 2) / float(ig))
        t[i, :] = np.float32(3) * x * x - np.float32(2) * x * x * x
    t[end:, :] = 1
    t[:, 3] *= np.float32(0.1)
    t[:, 1] *= np.float32(0.8)
    rgba = (np.clip(t, 0, 1) * np.float32(255)).astype(np.uint8)     # GetByteArray truncation
    nz = np.nonzero(t[:, 3] != 0)[0]
    return rgba, float(nz[0]), float(nz[-1])


def volume_numpy(size, z0=0, zn=None, ztotal=None):
    """sphere + ripple, uint8, [z][y][x]; z0/zn select a slab of a volume ztotal deep."""
    sx, sy, sz = size
    ztotal = sz if ztotal is None else ztotal
    zn = sz if zn is None else zn
    f = np.float32
    x = (f(0.5) - np.arange(sx, dtype=f) / f(sx))[None, None, :]
    y = (f(0.5) - np.arange(sy, dtype=f) / f(sy))[None, :, None]
    z = (f(0.5) - (np.arange(zn, dtype=f) + f(z0)) / f(ztotal))[:, None, None]
    d = np.sqrt(x * x + y * y + z * z)
    v = np.maximum(f(0), (f(0.5) - d) * f(255) * f(2))
    rip = f(18) * np.sin(x * f(40)) * np.sin(y * f(34)) * np.sin(z * f(46))
    return np.clip(v + rip * (v > 0), 0, 255).astype(np.uint8)


def volume_torch(size, device, z0=0, zn=None, ztotal=None, origin=None, total=None):
    """size = (sx, sy, sz) of the piece to generate; origin / total place it inside a larger volume (the z0 / zn /
    ztotal form is the z-only special case)."""
    import torch
    sx, sy, sz = size
    if origin is None:
        origin = (0, 0, z0)
        total = (sx, sy, sz if ztotal is None else ztotal)
        sz = sz if zn is None else zn
    tx, ty, tz = total
    x = (0.5 - (torch.arange(sx, device=device, dtype=torch.float32) + origin[0]) / tx)[None, None, :]
    y = (0.5 - (torch.arange(sy, device=device, dtype=torch.float32) + origin[1]) / ty)[None, :, None]
    z = (0.5 - (torch.arange(sz, device=device, dtype=torch.float32) + origin[2]) / tz)[:, None, None]
    d = torch.sqrt(x * x + y * y + z * z)
    v = torch.clamp((0.5 - d) * 255.0 * 2.0, min=0)
    rip = 18.0 * torch.sin(x * 40.0) * torch.sin(y * 34.0) * torch.sin(z * 46.0)
    return torch.clamp(v + rip * (v > 0), 0, 255).to(torch.uint8).contiguous()


class ClockSampler(threading.Thread):
    Q = "index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 7 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 7 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def camera_for_step(H, k, steps, n_slabs=1):
    cam = H.Camera()
    ang = 0.35 + 2.0 * math.pi * (k % max(steps, 1)) / max(steps, 1)
    cam.rotation = H.rotation_y(ang) @ H.rotation_x(0.35)
    return cam


# --------------------------------------------------------------------------- CPU arm
def oracle_setup(z0=0, ztotal=None):
    from oracle import host as H
    from oracle import oracle as O
    vol = volume_numpy(VOL, z0=z0, ztotal=ztotal)
    oc = O.build_octree(vol, BRICK, OVERLAP)
    pool = H.OraclePool(oc, (10, 10, 9))
    pool.upload_first_brick()
    req = []
    for i in range(pool.total - 1):
        req.append(pool.index_to_xyzl(i))
    pool.upload_requests(req)
    return H, O, oc, pool


def cpu_frame(H, pool, tf_rgba, lo, hi, k, steps, w, h):
    pool.recompute_visibility(0, lo, hi)
    p = H.raycast_params(camera_for_step(H, k, steps), w, h, pool.vol)
    frame = H.OracleFrame(w, h)
    t0 = time.perf_counter()
    H.raycast(p, pool, tf_rgba, frame)
    dt = time.perf_counter() - t0
    return int(frame.stats[0]), int(frame.stats[2]), dt, frame


def run_reference(args, rank):
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the CPU arm is meant to use every host core
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(cores)
    except OSError:
        pass
    H, O, oc, pool = oracle_setup()
    tf_rgba, lo, hi = bench_tf()
    cores = os.cpu_count() or 1
    w, h = CPU_SAMPLE
    for k in range(args.warmup):
        cpu_frame(H, pool, tf_rgba, lo, hi, k, args.steps, w, h)
    samples, secs = 0, 0.0
    for k in range(args.steps):
        s, r, dt, _ = cpu_frame(H, pool, tf_rgba, lo, hi, k, args.steps, w, h)
        samples += s
        secs += dt
    value = samples / secs
    sample = "same volume/TF/orbit, full %dx%d frames, %d steps" % (w, h, args.steps)
    line = {"impl": "reference", "metric": "ray samples/s", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# N = 1: BASELINE configs[1].  N > 1: the sort-last configuration (configs[3]) scaled so that every GPU owns
# exactly 512^3 voxels (weak scaling): the global volume doubles along x, then y, then z, cut into a regular grid of
# boxes at multiples of 64 voxels (sortlast.block_partition), rendered at the 3840x2160 that configs[3] names.
GLOBAL_BY_N = {1: (512, 512, 512), 2: (1024, 512, 512), 4: (1024, 1024, 512), 8: (1024, 1024, 1024)}


def global_config(n):
    g = GLOBAL_BY_N.get(n, (512, 512, 512 * n))
    return g, ((1920, 1080) if n == 1 else (3840, 2160))


def workload_config(n):
    g, (w, h) = global_config(n)
    return {"workload": "synthetic %dx%dx%d uint8 (sphere+ripple), brick 64/overlap 2, 1D-TF + ERT, %dx%d, "
                        "1 converged frame/step, orbiting camera" % (g[0], g[1], g[2], w, h),
            "volume": list(g), "resolution": [w, h], "render_mode": "1D-TF",
            "lighting": False, "sample_rate_modifier": 1.0, "sampler": os.environ.get("TVK_SAMPLER", "q4"),
            "l2": "inputs larger than L2 (236 MB pool) + 256 MB L2 flush between steps",
            "parallelism": "single GPU" if n == 1 else "sort-last, %d boxes on a regular grid + direct-send compositing" % n}


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args, rank, world):
    import torch
    import torch.distributed as dist
    from imagevis3d_b200 import _lib as L
    from imagevis3d_b200.context import TvkContext
    from imagevis3d_b200 import hostmath as HM

    global WIDTH, HEIGHT
    GLOBAL, (WIDTH, HEIGHT) = global_config(world)
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) must not join it
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        # ... and whatever a library still writes to file descriptor 1 goes to stderr: the JSON line is written to the
        # saved descriptor at the end
        sys.stdout.flush()
        global _REAL_STDOUT
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    tvk = TvkContext(local)
    if world > 1:
        # K1, NCCL and the compositor stay stream-ordered on ONE explicit stream (torch's legacy default stream has
        # handle 0, which tvk_set_stream reads as "use your own stream")
        work_stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(work_stream)
        tvk.set_stream(work_stream.cuda_stream)
    sampler = {"exact": L.SAMPLER_EXACT, "tex3d": L.SAMPLER_TEX3D, "q4": L.SAMPLER_EXACT_Q4, "gather": L.SAMPLER_GATHER}[os.environ.get("TVK_SAMPLER", "q4")]

    # ---- untimed set-up: synthesize the (slab of the) volume in HBM, build bricks, fill the pool
    from imagevis3d_b200 import sortlast as SL
    parts = SL.block_partition(GLOBAL, world, halo=64)          # world^(1/3)-ish boxes per ray instead of `world` slabs
    part = parts[rank]
    LOCAL = part["local_dims"] if world > 1 else VOL
    raw = volume_torch(LOCAL, dev, origin=part["s"], total=GLOBAL)
    torch.cuda.synchronize()
    t_b0 = time.perf_counter()
    info = tvk.build_octree(raw_dptr=raw.data_ptr(), volume_size=LOCAL, bpv=1, brick=BRICK, overlap=OVERLAP)
    t_build = time.perf_counter() - t_b0
    off, sizes = tvk.build_toc()
    mm = tvk.build_minmax()
    lod_count = next(l for l in range(info.lod_count)
                     if info.lod_bricks[l][0] * info.lod_bricks[l][1] * info.lod_bricks[l][2] == 1) + 1
    total_needed = int(info.lod_offset[lod_count - 1]) + 1
    nbx, nby = int(info.lod_bricks[0][0]) + 1, int(info.lod_bricks[0][1]) + 1
    cap = (10, 10, 9) if world == 1 else (nbx, nby, -(-total_needed // (nbx * nby)))
    total = tvk.pool_create(LOCAL, (BRICK,) * 3, OVERLAP, 1, cap, lod_count, sampler)
    store = tvk.build_store_ptr()
    slots = np.arange(total, dtype=np.uint32)
    slots[total - 1] = cap[0] * cap[1] * cap[2] - 1          # coarsest brick pinned in the last slot
    tvk.pool_upload_device(slots, [store + int(o) for o in off[:total]], sizes[:total])
    tvk.set_metadata(slots + 3)
    tvk.set_minmax(mm[:total])
    tf_rgba, lo, hi = bench_tf()
    tvk.set_tf1d(tf_rgba)
    tvk.recompute_visibility(0, lo, hi)

    # host copies for the e2e leg (pinned) and the CPU baseline
    blob = torch.empty(int(info.total_bytes), dtype=torch.uint8, pin_memory=True)
    tvk._ck(tvk.lib.tvk_build_get_bricks(tvk.h, C.c_void_p(blob.data_ptr()), C.c_uint64(0), C.c_uint64(blob.numel())), "get_bricks")
    meta_host = tvk.get_metadata()
    img8_host = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.uint8, pin_memory=True)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def params_for(k):
        cam = camera_for_step(HM, k, args.steps)
        if world == 1:
            return HM.raycast_params(cam, WIDTH, HEIGHT, VOL)
        return HM.raycast_params(cam, WIDTH, HEIGHT, GLOBAL, slab=part)

    class _CAI:                      # zero-copy torch view of the renderer's device image
        def __init__(self, ptr, shape):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (int(ptr), False), "version": 2}

    def gpu_composite(stripes):
        out = torch.empty_like(stripes[0])
        tvk.composite_ordered([t.data_ptr() for t in stripes], stripes[0].shape[0] * stripes[0].shape[1], out.data_ptr())
        return out

    final_image = [None]
    # sort-last exchange: fused into the kernels over NVLink peer memory (CUDA IPC); TVK_EXCHANGE=nccl selects the
    # collective-based direct send instead, which is also the fall-back when peer mapping is not possible
    exchange_kind, fused = "none", None
    if world > 1:
        exchange_kind = os.environ.get("TVK_EXCHANGE", "p2p-fused")
        if exchange_kind == "p2p-fused":
            ok = torch.ones(1, device=dev)
            try:
                fused = SL.FusedExchange(tvk, world, rank, WIDTH, HEIGHT, dist)
            except Exception as e:            # noqa: BLE001 -- any failure to map peers demotes ALL ranks together
                sys.stderr.write("rank %d: fused exchange unavailable (%s)\n" % (rank, e))
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                fused, exchange_kind = None, "nccl"

    # the camera path is part of the scene description: per-step parameters and slab orders are laid out up front
    step_params, step_order = {}, {}
    for k in range(max(args.steps, args.warmup)):
        step_params[k] = params_for(k)
        if world > 1:
            pg = HM.raycast_params(camera_for_step(HM, k, args.steps), WIDTH, HEIGHT, GLOBAL)
            step_order[k] = SL.depth_order(np.array(pg.eye_to_model[:]).reshape(4, 4), parts)

    def step(k):
        if world == 1:
            tvk.raycast(step_params[k])
            return
        order = step_order[k]
        if fused is not None:
            # K1 sends its stripes while it marches; stripe blend waits on arrival flags; result lands on rank 0
            ptr = fused.render(step_params[k], order)
            final_image[0] = torch.as_tensor(_CAI(ptr, (HEIGHT, WIDTH, 4)), device=dev)
        else:
            # direct-send stripes over NCCL, ordered blend on the owner, gather
            tvk.raycast(step_params[k])
            partial = torch.as_tensor(_CAI(tvk.image_device_ptr(), (HEIGHT, WIDTH, 4)), device=dev)
            final_image[0] = SL.exchange_and_composite(partial, order, gpu_composite, world, rank)

    # ---- warm-up
    for k in range(args.warmup):
        flush.zero_()
        torch.cuda.synchronize()
        step(k)
        tvk.synchronize()
    tvk.stats(reset=True)
    assert len(tvk.read_requests()) == 0, "pool is not fully resident: frame would not be converged"
    if fused is not None:
        assert not tvk.exchange_timed_out(), "sort-last exchange timed out during warm-up"

    # ---- timed region (device time from CUDA events around each step's launch)
    sampler_thread = ClockSampler(local)
    sampler_thread.start()
    time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    kernel_ms, step_s = 0.0, 0.0
    t0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()
        if world > 1:
            # the untimed L2 flush lets the ranks drift apart; without re-aligning them a rank's step time would include
            # the wait for the latest rank's flush, which is not part of the step
            dist.barrier()
        torch.cuda.synchronize()
        if world > 1:
            # device time of the whole step (march + exchange + blend + landing on rank 0): CUDA events on the one
            # stream every kernel of the step is launched on
            ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev_a.record(work_stream)
            step(k)
            ev_b.record(work_stream)
            tvk.synchronize()
            torch.cuda.synchronize()
            step_s += ev_a.elapsed_time(ev_b) * 1e-3
        else:
            step(k)
            tvk.synchronize()
            torch.cuda.synchronize()
        kernel_ms += tvk.stats().last_raycast_ms
    torch.cuda.synchronize()
    wall_s = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    st = tvk.stats()
    clocks = sampler_thread.stop()
    samples, rays = int(st.samples), int(st.rays)

    # ---- e2e: host buffers in, host image out, every step (pager upload + march + read-back)
    h2d = int(sum(int(np.prod(s)) for s in sizes[:total])) + 4 * total
    d2h = HEIGHT * WIDTH * 4        # the frame as glReadPixels(GL_RGBA, GL_UNSIGNED_BYTE) returns it (captureSingleFrame)
    srcs = [blob.data_ptr() + int(o) for o in off[:total]]
    n = len(srcs)
    src_arr = (C.c_void_p * n)(*srcs)
    slot_arr = np.ascontiguousarray(slots, np.uint32)
    size_arr = np.ascontiguousarray(sizes[:total], np.uint32).reshape(-1)

    e2e_mode, e2e_extra = "serial", {}
    if world > 1:
        # every rank pages in the bricks of ITS box that visibility leaves non-empty under the transfer function (see the
        # N=1 leg below, which also checks the frame against the all-bricks upload)
        keep = np.nonzero(np.asarray(meta_host, np.uint32)[:n] >= 3)[0]
        kn = len(keep)
        ksrc = (C.c_void_p * kn)(*[srcs[int(i)] for i in keep])
        kslots = np.ascontiguousarray(slot_arr[keep])
        ksizes = np.ascontiguousarray(size_arr.reshape(n, -1)[keep]).reshape(-1)
        h2d_rank = int(sum(int(np.prod(sizes[int(i)])) for i in keep)) + 4 * total
        tot = torch.tensor([h2d_rank, kn, n], device=dev, dtype=torch.int64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        h2d = int(tot[0])
        e2e_extra = {"upload": "per rank, the bricks of its box that the visibility pass leaves non-empty under the "
                               "transfer function: %d of %d over all ranks" % (int(tot[1]), int(tot[2])),
                     "h2d_bytes_scope": "sum over ranks"}

        def e2e_step(k):
            tvk.set_metadata(meta_host)
            tvk._ck(tvk.lib.tvk_pool_upload_bricks(tvk.h, C.c_uint32(kn), kslots.ctypes.data_as(C.c_void_p), ksrc,
                                                   ksizes.ctypes.data_as(C.c_void_p), C.c_int(0)), "upload")
            tvk._ck(tvk.lib.tvk_pool_upload_wait(tvk.h), "wait")
            tvk.set_metadata(meta_host)
            step(k)
            if rank == 0:
                img8_host.copy_((final_image[0].clamp(0.0, 1.0) * 255.0 + 0.5).to(torch.uint8))
                torch.cuda.synchronize()

        e2e_step(0)
        tvk.stats(reset=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for k in range(args.steps):
            e2e_step(k)
        tvk.synchronize()
        e2e_s = time.perf_counter() - t1
        e2e_samples = int(tvk.stats().samples)
    else:
        # Double-buffered frames: a second context whose pool holds TWO sets of slots.  While step k marches out of one
        # set (and its frame is read back), the pager already pulls step k+1's bricks over PCIe into the other set on
        # the side stream.  Every step still uploads all its bricks from pinned host memory and reads its frame back.
        e2e_mode = "double-buffered pool (upload of step k+1 overlaps march + read-back of step k)"
        e2 = TvkContext(local)
        e2.set_march_occupancy(int(os.environ.get("TVK_E2E_MARCH_CTAS", "3")))   # room for the pager's kernels on every SM
        cap2 = (cap[0], cap[1], 2 * cap[2])
        e2.pool_create(LOCAL, (BRICK,) * 3, OVERLAP, 1, cap2, lod_count, sampler)
        e2.set_minmax(mm[:total])
        e2.set_tf1d(tf_rgba)
        half = cap[0] * cap[1] * cap[2]
        slot_sets = [np.ascontiguousarray(slot_arr), np.ascontiguousarray(slot_arr + np.uint32(half))]
        # page tables of the two sets: the main context's table (visibility flags included), slot references shifted
        mh = np.ascontiguousarray(meta_host, np.uint32)
        metas = [mh, np.ascontiguousarray(np.where(mh >= 3, mh + np.uint32(half), mh).astype(np.uint32))]

        # What a step pages in: the bricks the visibility pass leaves non-empty under the transfer function.  A brick
        # whose [min, max] the table maps to alpha 0 is flagged EMPTY in the page table, no ray ever asks for it, and
        # the reference's pool never uploads it either (GLVolumePool::RecomputeVisibility, GLVolumePool.cpp:1791-1909;
        # GetBrick's empty flags, :484-656).  The all-bricks upload is timed too and reported beside it.
        keep_all = np.arange(n, dtype=np.int64)
        keep_vis = np.nonzero(mh[:n] >= 3)[0].astype(np.int64)

        def subset(keep):
            ks = (C.c_void_p * len(keep))(*[srcs[int(i)] for i in keep])
            kz = np.ascontiguousarray(size_arr.reshape(n, -1)[keep]).reshape(-1)
            return len(keep), [np.ascontiguousarray(ss[keep]) for ss in slot_sets], ks, kz

        def run_leg(keep):
            kn, kslots, ksrc, ksizes = subset(keep)

            def upload(k):
                e2._ck(e2.lib.tvk_pool_upload_bricks(e2.h, C.c_uint32(kn), kslots[k & 1].ctypes.data_as(C.c_void_p), ksrc,
                                                     ksizes.ctypes.data_as(C.c_void_p), C.c_int(2)), "upload")   # 2: one pinned allocation

            def run_steps(count):
                upload(0)
                for k in range(count):
                    e2._ck(e2.lib.tvk_pool_upload_wait(e2.h), "wait")          # step k's bricks are in place
                    if k + 1 < count:
                        upload(k + 1)                                          # side stream, other slot set
                    e2.set_metadata(metas[k & 1])
                    e2.raycast(step_params[k % args.steps])
                    e2._ck(e2.lib.tvk_read_image_rgba8(e2.h, C.c_void_p(img8_host.data_ptr()), C.c_int(1)), "read_image_rgba8")

            run_steps(2)
            e2.stats(reset=True)
            e2.synchronize()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            run_steps(args.steps)
            e2.synchronize()
            leg_s = time.perf_counter() - t1
            leg_samples = int(e2.stats().samples)
            if len(e2.read_requests()) != 0:
                raise RuntimeError("e2e pool not fully resident: rays asked for bricks that were not uploaded")
            return leg_s, leg_samples

        try:
            e2e_s, e2e_samples = run_leg(keep_vis)        # first: on a pool that never held the empty bricks
            vis8 = img8_host.clone()
        except RuntimeError as e:      # reported, never hidden: the line then carries the all-bricks figure
            sys.stderr.write("e2e: visible-brick leg rejected (%s); reporting the all-bricks upload\n" % e)
            e2e_extra = {"upload": "all %d bricks" % n, "visible_leg_error": str(e)}
            vis8 = None
        all_s, all_samples = run_leg(keep_all)
        if vis8 is not None and (not torch.equal(vis8, img8_host) or e2e_samples != all_samples):
            # same context, same slot set, same camera: the two frames must be the same bytes
            sys.stderr.write("e2e: visible-brick frame differs from the all-bricks frame; reporting the all-bricks upload\n")
            e2e_extra = {"upload": "all %d bricks" % n, "visible_leg_error": "frame or sample count differs from the all-bricks leg"}
            vis8 = None
        if vis8 is None:
            e2e_s, e2e_samples = all_s, all_samples
        else:
            # informational: against the main context's fully resident pool (another pool geometry, so the pool-space
            # coordinate arithmetic rounds differently in the last place)
            ref8 = torch.empty_like(img8_host)
            tvk.raycast(step_params[(args.steps - 1) % args.steps])
            tvk._ck(tvk.lib.tvk_read_image_rgba8(tvk.h, C.c_void_p(ref8.data_ptr()), C.c_int(1)), "read_image_rgba8")
            tvk.synchronize()
            dmax = int((ref8.to(torch.int16) - vis8.to(torch.int16)).abs().max())
            h2d_all = h2d
            h2d = int(sum(int(np.prod(sizes[int(i)])) for i in keep_vis)) + 4 * total
            e2e_extra = {"upload": "bricks the visibility pass leaves non-empty under the transfer function: %d of %d "
                                   "(empty bricks are never requested, as in GLVolumePool)" % (len(keep_vis), n),
                         "frame_equals_all_bricks_frame": True, "max_abs_diff_vs_resident_pool_rgba8": dmax,
                         "all_bricks": {"value": all_samples / all_s, "ms_per_step": 1e3 * all_s / args.steps,
                                        "h2d_bytes_per_step": h2d_all}}
        e2.close()

    # ---- reduce over ranks
    if world > 1:
        t = torch.tensor([kernel_ms, wall_s, e2e_s, step_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kernel_ms, wall_s, e2e_s, step_s = [float(v) for v in t.tolist()]
        c = torch.tensor([samples, rays, e2e_samples], device=dev, dtype=torch.int64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        samples, rays, e2e_samples = [int(v) for v in c.tolist()]
    if rank != 0:
        tvk.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    traffic = None
    try:        # DRAM bytes per launch of this kernel from the committed ncu --set full capture (N=1 workload only)
        if world == 1:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))[os.environ.get("TVK_SAMPLER", "q4")]["dram_bytes_per_launch"]
    except Exception:       # noqa: BLE001
        traffic = None
    # instruction-issue view of the same kernel (what actually bounds it): warp instructions per sample from the committed
    # ncu capture x the samples of this run, against 4 warp instructions per clock per SM
    issue = None
    try:
        if world == 1:
            k = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))[os.environ.get("TVK_SAMPLER", "q4")]
            per_sample = k["warp_instructions_per_launch"] / k["samples_in_capture"]
            sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
            sm_mhz = (clocks or {}).get("sm_mhz") or 1965
            issue_peak = sm_count * 4 * sm_mhz * 1e6
            issue_rate = per_sample * samples / (kernel_ms * 1e-3)
            issue = {"achieved": issue_rate / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instr/s", "frac": issue_rate / issue_peak,
                     "warp_instructions_per_sample": per_sample, "source": "profiles/k1_traffic.json (ncu smsp__inst_executed.sum)"}
    except Exception:       # noqa: BLE001
        issue = None
    ms_step = kernel_ms / args.steps
    if world > 1:
        # a step = render + exchange + blend + gather: the ray-march kernel alone is not the step
        ms_frame = 1e3 * step_s / args.steps
    else:
        ms_frame = ms_step
    value = samples / (ms_frame * 1e-3 * args.steps)
    alg_bytes = (samples * BYTES_PER_SAMPLE + rays * BYTES_PER_RAY) / args.steps / world
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    line = {"metric": "ray samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_frame, "raycast_kernel_ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "samples_per_step": samples / args.steps, "rays_per_step": rays / args.steps,
            "wall_ms_per_step_incl_l2_flush": 1e3 * wall_s / args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "0.6068 B/sample + 16 B/ray (SURVEY 8d); the kernel is instruction-issue bound, see `issue` and DESIGN.md",
                         "issue": issue},
            "e2e": {"value": e2e_samples / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "result": "RGBA8 frame (glReadPixels GL_UNSIGNED_BYTE, as captureSingleFrame reads it)", "pipeline": e2e_mode,
                    **e2e_extra},
            "gpu_launches": args.steps * (1 if world == 1 else (2 * world if fused is not None else 2 * world)), "clocks": clocks,
            "builder_ms_untimed": 1e3 * t_build}
    if world > 1:
        line["config"]["exchange"] = exchange_kind
        if fused is not None:
            line["exchange_timed_out"] = bool(tvk.exchange_timed_out())

    # ---- CPU baseline (rank 0, N=1 only): the oracle on the host cores, bounded sample
    if world == 1 and not args.no_cpu:
        from oracle import host as H
        oc_like = dict(plan=None)
        pool = build_oracle_pool_from_store(H, info, off, sizes, blob.numpy(), mm, total, cap, lod_count, slots, meta_host)
        w, h = CPU_SAMPLE
        frame = H.OracleFrame(w, h)
        nsteps = min(CPU_STEPS, args.steps)
        cs, dt = 0, 0.0
        for k in range(-1, nsteps):                       # k = -1: untimed warm-up frame
            p = H.raycast_params(camera_for_step(H, max(k, 0), args.steps), w, h, VOL)
            tc = time.perf_counter()
            H.raycast(p, pool, tf_rgba, frame)
            if k >= 0:
                dt += time.perf_counter() - tc
                cs += int(frame.stats[0])
        line["cpu_baseline"] = {"value": cs / dt, "unit": "samples/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": "first %d steps of the same orbit at full %dx%d, %.1f s of CPU work"
                                          % (nsteps, w, h, dt)}
    emit_line(line)
    tvk.close()
    if world > 1:
        dist.destroy_process_group()


def build_oracle_pool_from_store(H, info, off, sizes, blob, mm, total, cap, lod_count, slots, meta):
    """Host copy of the device pool for the CPU baseline (bricks come from the CUDA builder, which the
    tests prove bit-identical to the oracle builder)."""
    class P:
        pass
    pool = P()
    pool.vol, pool.bpv, pool.overlap, pool.max_brick, pool.lod_count, pool.capacity = VOL, 1, OVERLAP, (BRICK,) * 3, lod_count, cap
    pool.layout, pool.lod_offset, pool.lod_bricks, pool.total = H.pool_layout(VOL, pool.max_brick, OVERLAP, lod_count)
    pool.data = np.zeros((cap[2] * BRICK, cap[1] * BRICK, cap[0] * BRICK), np.uint8)
    for i in range(total):
        s = sizes[i]
        b = blob[int(off[i]):int(off[i]) + int(np.prod(s))].reshape(int(s[2]), int(s[1]), int(s[0]))
        sl = int(slots[i])
        o = ((sl % cap[0]) * BRICK, ((sl // cap[0]) % cap[1]) * BRICK, (sl // (cap[0] * cap[1])) * BRICK)
        pool.data[o[2]:o[2] + b.shape[0], o[1]:o[1] + b.shape[1], o[0]:o[0] + b.shape[2]] = b
    pool.meta = np.ascontiguousarray(meta, np.uint32)
    pool.c_pool = lambda: H.OraclePool.c_pool(pool)
    return pool


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    run_gpu(args, rank, world)


if __name__ == "__main__":
    main()
