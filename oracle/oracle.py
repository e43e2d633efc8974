"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (imagevis3d_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
ORC_MAX_LOD = 32


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("tvk_oracle_builder.c", "tvk_oracle_render.c", "tvk_oracle_quantize.c", "tvk_oracle_classic.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


class Plan(C.Structure):
    _fields_ = [("lod_count", C.c_uint32),
                ("lod_size", (C.c_uint64 * 3) * ORC_MAX_LOD),
                ("lod_bricks", (C.c_uint64 * 3) * ORC_MAX_LOD),
                ("lod_aspect", (C.c_double * 3) * ORC_MAX_LOD),
                ("lod_offset", C.c_uint64 * ORC_MAX_LOD),
                ("total_bricks", C.c_uint64),
                ("brick_size", C.c_uint64 * 3),
                ("overlap", C.c_uint32)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_toc.restype = C.c_uint64
        _LIB.orc_hist1d.restype = C.c_uint32
        _LIB.orc_hist2d.restype = C.c_float
    return _LIB


def _vol3(vol):
    return (C.c_uint64 * 3)(*[int(v) for v in vol])


def plan(vol, brick=64, overlap=2):
    """vol = (sx, sy, sz).  Returns the Plan struct (ExtendedOctree::ComputeMetadata)."""
    p = Plan()
    rc = lib().orc_plan(_vol3(vol), C.c_uint64(brick), C.c_uint32(overlap), C.byref(p))
    if rc:
        raise ValueError("orc_plan failed (%d)" % rc)
    return p


def toc(p, bpv):
    n = int(p.total_bricks)
    offsets = np.zeros(n, np.uint64)
    sizes = np.zeros((n, 3), np.uint32)
    total = lib().orc_toc(C.byref(p), C.c_uint32(bpv), offsets.ctypes.data_as(C.c_void_p),
                          sizes.ctypes.data_as(C.c_void_p))
    return int(total), offsets, sizes


def build_octree(raw, brick=64, overlap=2, median=False, clamp=False):
    """raw: numpy array [z][y][x] of uint8/uint16.  Returns dict with plan, toc and data."""
    raw = np.ascontiguousarray(raw)
    bpv = raw.dtype.itemsize
    vol = raw.shape[::-1]
    p = plan(vol, brick, overlap)
    total, offsets, sizes = toc(p, bpv)
    store = np.zeros(total, np.uint8)
    minmax = np.zeros((int(p.total_bricks), 2), np.float64)
    rc = lib().orc_build_ex(raw.ctypes.data_as(C.c_void_p), _vol3(vol), C.c_uint32(bpv),
                            C.c_uint64(brick), C.c_uint32(overlap), C.c_int(1 if median else 0), C.c_int(1 if clamp else 0),
                            store.ctypes.data_as(C.c_void_p), minmax.ctypes.data_as(C.c_void_p))
    if rc:
        raise ValueError("orc_build failed (%d)" % rc)
    return dict(plan=p, offsets=offsets, sizes=sizes, store=store, minmax=minmax, bpv=bpv,
                vol=vol, brick=brick, overlap=overlap)


def hist1d(oct):
    bins = np.zeros(1 << (8 * oct["bpv"]), np.uint64)
    n = lib().orc_hist1d(oct["store"].ctypes.data_as(C.c_void_p), _vol3(oct["vol"]),
                         C.c_uint32(oct["bpv"]), C.c_uint64(oct["brick"]), C.c_uint32(oct["overlap"]),
                         bins.ctypes.data_as(C.c_void_p))
    return bins[:n].copy()


def hist2d(oct, bins_x, max_nonzero):
    bins = np.zeros((bins_x, 256), np.uint64)
    mg = lib().orc_hist2d(oct["store"].ctypes.data_as(C.c_void_p), _vol3(oct["vol"]),
                          C.c_uint32(oct["bpv"]), C.c_uint64(oct["brick"]), C.c_uint32(oct["overlap"]),
                          C.c_uint32(bins_x), C.c_double(max_nonzero), bins.ctypes.data_as(C.c_void_p))
    return bins, float(mg)


def brick(oct, i):
    s = oct["sizes"][i]
    n = int(s[0]) * int(s[1]) * int(s[2]) * oct["bpv"]
    o = int(oct["offsets"][i])
    dt = np.uint8 if oct["bpv"] == 1 else np.uint16
    return oct["store"][o:o + n].view(dt).reshape(int(s[2]), int(s[1]), int(s[0]))


# ---------------------------------------------------------------- quantisation (B5)
QUANT_TYPES = {"i8": (0, np.int8), "u8": (1, np.uint8), "i16": (2, np.int16), "u16": (3, np.uint16), "i32": (4, np.int32),
               "u32": (5, np.uint32), "f32": (6, np.float32), "i64": (7, np.int64), "u64": (8, np.uint64), "f64": (9, np.float64)}


class QuantResult(C.Structure):
    _fields_ = [("out_bits", C.c_uint32), ("generated", C.c_uint32), ("hist_size", C.c_uint32), ("bin_count", C.c_uint32)]


def quantize(data, to8bit=False):
    """RAWConverter.cpp quantize() on a 1-D numpy array.  Returns dict(out_bits, generated, data, hist):
    `data` is the quantised stream (uint8/uint16) if generated, else the input itself; `hist` is None
    when the reference leaves the histogram to Histogram1DDataBlock::Compute."""
    data = np.ascontiguousarray(data).reshape(-1)
    code = [c for c, (k, dt) in QUANT_TYPES.items() if np.dtype(dt) == data.dtype][0]
    out = np.zeros(2 * data.size, np.uint8)
    hist = np.zeros(4096, np.uint64)
    r = QuantResult()
    rc = lib().orc_quantize(data.ctypes.data_as(C.c_void_p), C.c_uint32(QUANT_TYPES[code][0]), C.c_uint64(data.size),
                            C.c_int(1 if to8bit else 0), out.ctypes.data_as(C.c_void_p), hist.ctypes.data_as(C.c_void_p), C.byref(r))
    if rc:
        raise ValueError("orc_quantize failed (%d)" % rc)
    if r.generated:
        q = out[:data.size].copy() if r.out_bits == 8 else out.view(np.uint16)[:data.size].copy()
    else:
        q = data
    return dict(out_bits=int(r.out_bits), generated=bool(r.generated), data=q,
                hist=hist[:r.hist_size].copy() if r.hist_size else None, bin_count=int(r.bin_count))


def ref_quantize(data, to8bit=False, tmpdir=None):
    """The same through the reference's own templates (oracle/_ref/ref_driver quantize)."""
    import tempfile
    data = np.ascontiguousarray(data).reshape(-1)
    code = [c for c, (k, dt) in QUANT_TYPES.items() if np.dtype(dt) == data.dtype][0]
    with tempfile.TemporaryDirectory(dir=tmpdir) as td:
        fi, fo = os.path.join(td, "in.raw"), os.path.join(td, "out.bin")
        data.tofile(fi)
        subprocess.check_call([ref_driver_path(), "quantize", fi, code, str(data.size), "1" if to8bit else "0", fo],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        buf = open(fo, "rb").read()
    bits, gen = [int(v) for v in np.frombuffer(buf, np.uint32, 2, 0)]
    nb = int(np.frombuffer(buf, np.uint64, 1, 8)[0])
    raw = np.frombuffer(buf, np.uint8, nb, 16)
    o = 16 + nb
    nh = int(np.frombuffer(buf, np.uint64, 1, o)[0])
    hist = np.frombuffer(buf, np.uint64, nh, o + 8).copy() if nh else None
    # (Quantize<T,U> pre-sizes its output file to bytes*sizeof(U); only the first `size` elements are data)
    q = (raw.copy() if bits == 8 else raw.view(np.uint16).copy())[:data.size] if gen else data
    return dict(out_bits=bits, generated=bool(gen), data=q, hist=hist)


# ---------------------------------------------------------------- reference dumps
def parse_ref_convert(path, bpv):
    """Parses the dump written by oracle/_ref/ref_driver convert (layout: see ref_driver.cpp)."""
    buf = open(path, "rb").read()
    assert buf[:8] == b"TVKREF1\0"
    o = 8
    lods = int(np.frombuffer(buf, np.uint32, 1, o)[0]); o += 4
    lod_size, lod_bricks, lod_aspect = [], [], []
    for _ in range(lods):
        lod_size.append(np.frombuffer(buf, np.uint64, 3, o)); o += 24
        lod_bricks.append(np.frombuffer(buf, np.uint64, 3, o)); o += 24
        lod_aspect.append(np.frombuffer(buf, np.float64, 3, o)); o += 24
    total = int(np.frombuffer(buf, np.uint64, 1, o)[0]); o += 8
    sizes = np.zeros((total, 3), np.uint64)
    minmax = np.zeros((total, 2), np.float64)
    for i in range(total):
        sizes[i] = np.frombuffer(buf, np.uint64, 3, o); o += 24
        minmax[i] = np.frombuffer(buf, np.float64, 2, o); o += 16
    nbytes = int((sizes.prod(axis=1) * bpv).sum())
    store = np.frombuffer(buf, np.uint8, nbytes, o); o += nbytes
    n1 = int(np.frombuffer(buf, np.uint64, 1, o)[0]); o += 8
    h1 = np.frombuffer(buf, np.uint64, n1, o); o += 8 * n1
    mg = float(np.frombuffer(buf, np.float32, 1, o)[0]); o += 4
    nx, ny = [int(v) for v in np.frombuffer(buf, np.uint64, 2, o)]; o += 16
    h2 = np.frombuffer(buf, np.uint64, nx * ny, o).reshape(nx, ny) if nx else None
    return dict(lod_size=np.array(lod_size), lod_bricks=np.array(lod_bricks),
                lod_aspect=np.array(lod_aspect), sizes=sizes, minmax=minmax, store=store,
                hist1d=h1, max_grad=mg, hist2d=h2)


def ref_driver_path():
    p = os.path.join(_HERE, "_ref", "ref_driver")
    return p if os.path.exists(p) else None
