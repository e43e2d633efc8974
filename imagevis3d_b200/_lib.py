"""ctypes binding of include/tvk_cuda.h (libtvk_cuda.so).

There is NO fallback: if the CUDA library has not been built (or cannot be loaded)
every entry point of this package raises.  Build it with
`python -m imagevis3d_b200.build` or `__graft_entry__.build()`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TVK_CUDA_LIB") or os.path.join(_HERE, "libtvk_cuda.so")   # override: kernel-variant sweeps
TVK_MAX_LOD = 16

TVK_OK, TVK_ERR_INVALID, TVK_ERR_CUDA, TVK_ERR_NOMEM, TVK_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
RM_1DTRANS, RM_2DTRANS, RM_ISOSURFACE = 0, 1, 2
MBS_ONLY_NEEDED, MBS_REQUEST_ALL, MBS_SKIP_ONE_LEVEL, MBS_SKIP_TWO_LEVELS = 0, 1, 2, 3
SAMPLER_EXACT, SAMPLER_TEX3D, SAMPLER_EXACT_Q4, SAMPLER_GATHER = 0, 1, 2, 3
BI_MISSING, BI_CHILD_EMPTY, BI_EMPTY, BI_FLAG_COUNT = 0, 1, 2, 3


class PoolDesc(C.Structure):
    _fields_ = [("volume_size", C.c_uint32 * 3), ("max_brick_size", C.c_uint32 * 3), ("overlap", C.c_uint32),
                ("bytes_per_voxel", C.c_uint32), ("pool_capacity", C.c_uint32 * 3), ("lod_count", C.c_uint32),
                ("sampler", C.c_uint32)]


class Visibility(C.Structure):
    _fields_ = [("render_mode", C.c_uint32), ("smin", C.c_double), ("smax", C.c_double),
                ("gmin", C.c_double), ("gmax", C.c_double)]


class RaycastParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32),
                ("eye_to_model", C.c_float * 16), ("model_to_eye", C.c_float * 16),
                ("model_view_it", C.c_float * 16), ("inv_projection", C.c_float * 16),
                ("lod_factor", C.c_float), ("level_zero_wse", C.c_float), ("sample_rate_modifier", C.c_float),
                ("trans_scale", C.c_float), ("gradient_scale", C.c_float), ("isoval", C.c_float),
                ("render_mode", C.c_uint32), ("lighting", C.c_uint32),
                ("light_ambient", C.c_float * 3), ("light_diffuse", C.c_float * 3), ("light_specular", C.c_float * 3),
                ("model_light_dir", C.c_float * 3), ("model_eye_pos", C.c_float * 3), ("domain_scale", C.c_float * 3),
                ("iso_color", C.c_float * 3), ("compose_light_dir", C.c_float * 3),
                ("new_view", C.c_uint32), ("missing_strategy", C.c_uint32), ("rehash_count", C.c_uint32),
                ("hash_table_size", C.c_uint32), ("clip_enabled", C.c_uint32), ("clip_plane", C.c_float * 4),
                ("box_min", C.c_float * 3), ("box_max", C.c_float * 3), ("count_samples", C.c_uint32),
                ("model_pos_scale_m1", C.c_float * 3), ("model_pos_offset", C.c_float * 3)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64), ("brick_steps", C.c_uint64),
                ("last_raycast_ms", C.c_float), ("launches", C.c_uint32)]


class BuildDesc(C.Structure):
    _fields_ = [("volume_size", C.c_uint32 * 3), ("brick_size", C.c_uint32), ("overlap", C.c_uint32),
                ("bytes_per_voxel", C.c_uint32), ("use_median", C.c_uint32), ("clamp_to_edge", C.c_uint32)]


class BuildInfo(C.Structure):
    _fields_ = [("lod_count", C.c_uint32), ("lod_size", (C.c_uint32 * 3) * TVK_MAX_LOD),
                ("lod_bricks", (C.c_uint32 * 3) * TVK_MAX_LOD), ("lod_offset", C.c_uint64 * TVK_MAX_LOD),
                ("total_bricks", C.c_uint64), ("total_bytes", C.c_uint64)]


class RebrickDesc(C.Structure):
    _fields_ = [("lod_count", C.c_uint32), ("lod_size", (C.c_uint32 * 3) * TVK_MAX_LOD), ("src_brick", C.c_uint32 * 3),
                ("tgt_brick", C.c_uint32 * 3), ("overlap", C.c_uint32), ("bytes_per_voxel", C.c_uint32)]


class RebrickInfo(C.Structure):
    _fields_ = [("tgt_brick", C.c_uint32 * 3), ("lod_bricks", (C.c_uint32 * 3) * TVK_MAX_LOD), ("lod_offset", C.c_uint64 * TVK_MAX_LOD),
                ("src_lod_bricks", (C.c_uint32 * 3) * TVK_MAX_LOD), ("src_lod_offset", C.c_uint64 * TVK_MAX_LOD),
                ("total_bricks", C.c_uint64), ("src_total_bricks", C.c_uint64), ("total_bytes", C.c_uint64)]


class SliceParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("window_mode", C.c_uint32), ("use_mip", C.c_uint32),
                ("slice", C.c_float), ("voxels", C.c_uint32 * 3), ("bytes_per_voxel", C.c_uint32), ("min_coords", C.c_float * 3),
                ("max_coords", C.c_float * 3), ("quad", C.c_float * 2), ("render_mode", C.c_uint32), ("trans_scale", C.c_float),
                ("gradient_scale", C.c_float)]


class ClearViewParams(C.Structure):
    _fields_ = [("cv_isoval", C.c_float), ("cv_color", C.c_float * 3), ("size", C.c_float), ("context_scale", C.c_float),
                ("border_scale", C.c_float), ("pick_pos", C.c_float * 3)]


class MipParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("inv_model_view", C.c_float * 16), ("inv_projection", C.c_float * 16),
                ("bytes_per_voxel", C.c_uint32), ("trans_scale", C.c_float)]


class ClassicBrick(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("extension", C.c_float * 3), ("eye_to_tex", C.c_float * 16),
                ("ray_step", C.c_float), ("voxels", C.c_uint32 * 3), ("store_offset", C.c_uint64)]


class ClassicParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("inv_model_view", C.c_float * 16),
                ("normal_matrix", C.c_float * 9), ("inv_projection", C.c_float * 16), ("step_scale", C.c_float),
                ("trans_scale", C.c_float), ("lighting", C.c_uint32), ("light_ambient", C.c_float * 3),
                ("light_diffuse", C.c_float * 3), ("light_specular", C.c_float * 3), ("light_dir", C.c_float * 3),
                ("domain_scale", C.c_float * 3), ("bytes_per_voxel", C.c_uint32), ("render_mode", C.c_uint32),
                ("gradient_scale", C.c_float), ("isoval", C.c_float), ("proj_param", C.c_float * 2),
                ("iso_color", C.c_float * 3), ("clip_enabled", C.c_uint32), ("clip_plane", C.c_float * 4)]


class QuantizeResult(C.Structure):
    _fields_ = [("out_bits", C.c_uint32), ("generated", C.c_uint32), ("hist_size", C.c_uint32), ("bin_count", C.c_uint32),
                ("device_ms", C.c_float)]


# TVK_DT_* codes by numpy dtype name
DTYPE_CODES = {"int8": 0, "uint8": 1, "int16": 2, "uint16": 3, "int32": 4, "uint32": 5, "float32": 6, "int64": 7,
               "uint64": 8, "float64": 9}


# every symbol include/tvk_cuda.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "tvk_create", "tvk_destroy", "tvk_last_error", "tvk_device_synchronize", "tvk_set_stream", "tvk_set_march_occupancy",
    "tvk_pool_create", "tvk_pool_destroy", "tvk_pool_total_bricks", "tvk_pool_upload_bricks",
    "tvk_pool_upload_wait", "tvk_pool_upload_bricks_lz4", "tvk_lz4_inflate", "tvk_zlib_inflate", "tvk_pool_upload_bricks_compressed", "tvk_pool_upload_fence", "tvk_pool_set_metadata", "tvk_pool_clear_metadata", "tvk_pool_patch_metadata", "tvk_pool_get_metadata",
    "tvk_visibility_set_minmax", "tvk_visibility_recompute", "tvk_set_tf1d", "tvk_set_tf2d", "tvk_tf2d_staging",
    "tvk_raycast_subframe", "tvk_read_requests", "tvk_read_image", "tvk_read_image_rgba8", "tvk_read_resume", "tvk_image_device_ptr",
    "tvk_get_stats", "tvk_probe_lod", "tvk_probe_brick_index",
    "tvk_build_plan", "tvk_build_octree", "tvk_build_get_toc", "tvk_build_get_minmax", "tvk_build_get_bricks",
    "tvk_build_store_device_ptr", "tvk_build_hist1d", "tvk_build_hist2d", "tvk_build_release", "tvk_build_last_ms", "tvk_build_hist2d_last_ms", "tvk_quantize",
    "tvk_build_forget_raw", "tvk_build_stream_begin", "tvk_build_stream_rows", "tvk_build_stream_slab", "tvk_build_stream_end",
    "tvk_rebrick_plan", "tvk_rebrick_toc", "tvk_rebrick",
    "tvk_composite_ordered", "tvk_raycast_classic", "tvk_raycast_classic_cv", "tvk_read_iso_hit", "tvk_raycast_mip", "tvk_render_slice", "tvk_select_eye", "tvk_read_eye_image", "tvk_compose_stereo",
    "tvk_exchange_create", "tvk_exchange_open_peers", "tvk_exchange_connect", "tvk_exchange_peer_base", "tvk_exchange_set_peers", "tvk_exchange_local_base",
    "tvk_exchange_begin_frame", "tvk_exchange_composite", "tvk_exchange_final", "tvk_exchange_read_final_rgba8", "tvk_exchange_status", "tvk_exchange_destroy",
]

_lib = None


class TvkError(RuntimeError):
    pass


def load():
    """Loads libtvk_cuda.so; raises if it is missing (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TvkError("libtvk_cuda.so is not built (%s); run `python -m imagevis3d_b200.build`. "
                       "This package has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.tvk_last_error.restype = C.c_char_p
    lib.tvk_last_error.argtypes = [C.c_void_p]
    lib.tvk_destroy.restype = None
    lib.tvk_destroy.argtypes = [C.c_void_p]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("tvk_last_error", "tvk_destroy"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(ctx, rc, what):
    if rc != TVK_OK:
        msg = load().tvk_last_error(ctx).decode() if ctx else ""
        raise TvkError("%s failed with status %d: %s" % (what, rc, msg))
