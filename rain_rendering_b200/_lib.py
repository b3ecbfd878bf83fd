"""ctypes binding of librain_b200.so (include/rain_b200.h).

This is the stub a maintainer of the reference would add: the reference is
pure Python, so its FFI is ctypes.  There is NO fallback: if the shared library is missing it
is built with nvcc (``build.py``); if that fails, importing raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RR_LIB_OVERRIDE") or os.path.join(_HERE, "librain_b200.so")     # override: tuning experiments only


class Camera(C.Structure):
    """rr_camera"""
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("focal_m", C.c_double), ("f_number", C.c_double),
                ("exposure_ms", C.c_double), ("gain", C.c_double), ("focus_plane_m", C.c_double),
                ("pix_size_m", C.c_double), ("radius", C.c_double), ("fov_deg", C.c_double),
                ("opacity_att", C.c_double), ("fallrate_mmh", C.c_double), ("render_scale", C.c_int32), ("reserved", C.c_int32)]


class FrameIO(C.Structure):
    """rr_frame_io"""
    _fields_ = [("bgr", C.c_void_p), ("depth", C.c_void_p), ("depth_format", C.c_int32), ("reserved", C.c_int32),
                ("streaks", C.c_void_p), ("streak_offsets", C.c_void_p), ("out_bgr", C.c_void_p), ("out_mask", C.c_void_p),
                ("out_bgr_u8", C.c_void_p), ("out_mask_idx8", C.c_void_p), ("out_mask_u16", C.c_void_p), ("out_mask_range", C.c_void_p),
                ("out_png_image", C.c_void_p), ("out_png_mask", C.c_void_p), ("out_png_image_sizes", C.c_void_p),
                ("out_png_mask_sizes", C.c_void_p), ("png_stride", C.c_size_t)]


DEPTH_F32_M, DEPTH_U16_256 = 0, 1


class SimParams(C.Structure):
    """rr_sim_params"""
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("focal_m", C.c_double), ("pix_size_m", C.c_double), ("exposure_ms", C.c_double),
                ("fallrate_mmh", C.c_double), ("sim_hz", C.c_double), ("cam_speed_kmh", C.c_double), ("z_near", C.c_double),
                ("z_far", C.c_double), ("d_min_mm", C.c_double), ("d_max_mm", C.c_double), ("min_width_px", C.c_double),
                ("seed", C.c_uint64)]


# rr_plan (csrc/rr_types.h) as a numpy dtype, for the stage parity tests
PLAN_DTYPE = np.dtype([
    ("valid", "<i4"), ("type", "<i4"), ("pw", "<i4"), ("ph", "<i4"), ("minx", "<i4"), ("miny", "<i4"),
    ("tex_off", "<i4"), ("tex_h", "<i4"), ("M", "<f8", 9), ("bw0", "<i4"), ("nW", "<i4"), ("nH", "<i4"),
    ("flip", "<i4"), ("resize_mode", "<i4"), ("_pad0", "<i4"), ("scale_x", "<f8"), ("scale_y", "<f8"),
    ("sig_y", "<f8"), ("sig_x", "<f8"), ("shift", "<i4"), ("ry", "<i4"), ("rx", "<i4"),
    ("bx0", "<i4"), ("by0", "<i4"), ("cropx", "<i4"), ("cropy", "<i4"), ("bw", "<i4"), ("bh", "<i4"), ("_pad1", "<i4"),
    ("g_off", "<i8"), ("a_off", "<i8"), ("kb", "<f8"), ("kg", "<f8"), ("kr", "<f8"), ("a_scale", "<f8"),
    ("c_scale", "<f8"), ("fov_x", "<f8"), ("fov_y", "<f8"), ("drop_Y", "<f8")])
assert PLAN_DTYPE.itemsize == 280

# rr_xml_frame
XML_FRAME_DTYPE = np.dtype([("id", "<i4"), ("t", "<i4"), ("d", "<i4"), ("rs", "<i4"), ("first", "<i8"), ("count", "<i8")])
assert XML_FRAME_DTYPE.itemsize == 32

T_NAMES = ("h2d", "fog", "env", "setup", "raster", "blur", "composite", "epilogue", "d2h", "total")
DBG = dict(fog=0, env=1, omega=2, plans=3, rainy=4, env_src=5, arena=6, fext=7)

_lib = None


def shutil_which_nvcc():
    try:
        return _build.find_nvcc()
    except Exception:
        return None


def load() -> C.CDLL:
    """Load (building first if needed) the CUDA library.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("RR_LIB_OVERRIDE") and (not os.path.exists(LIB_PATH) or _build.stale()):
        try:
            _build.build()
        except Exception as e:  # no silent fallback: neither to a CPU path nor to a library older than its sources
            if not os.path.exists(LIB_PATH):
                raise RuntimeError("librain_b200.so is missing and could not be built (%s); there is no CPU fallback" % e)
            if shutil_which_nvcc():
                raise RuntimeError("librain_b200.so is older than its sources and rebuilding it failed: %s" % e)
            import warnings
            warnings.warn("librain_b200.so is older than its sources and there is no nvcc to rebuild it; loading the existing library")
    lib = C.CDLL(LIB_PATH)
    vp, i32p, u8p, f32p, f64p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    lib.rr_version.restype = C.c_int
    lib.rr_last_error.restype = C.c_char_p
    protos = {
        "rr_create": [C.c_int, C.POINTER(vp)],
        "rr_destroy": [vp],
        "rr_set_streak_db": [vp, C.c_int, i32p, C.c_int, u8p],
        "rr_alloc_streak_db": [vp, C.c_int, i32p, C.c_int],
        "rr_streak_db_device_ptr": [vp, C.POINTER(vp), C.POINTER(C.c_size_t)],
        "rr_set_camera": [vp, C.POINTER(Camera), C.c_int],
        "rr_env_size": [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "rr_render_frames": [vp, C.c_int, u8p, f32p, vp, i32p, f32p, f32p, u8p],
        "rr_submit_frames": [vp, C.c_int, u8p, f32p, vp, i32p, f32p, f32p, u8p],
        "rr_wait_frames": [vp],
        "rr_render_frames_io": [vp, C.c_int, C.POINTER(FrameIO)],
        "rr_submit_frames_io": [vp, C.c_int, C.POINTER(FrameIO)],
        "rr_render_frames_device_io": [vp, C.c_int, C.POINTER(FrameIO), C.c_int],
        "rr_render_frames_device": [vp, C.c_int, u8p, f32p, vp, i32p, f32p, f32p, u8p, C.c_int],
        "rr_fog_only": [vp, C.c_int, u8p, f32p, f64p],
        "rr_envmap_only": [vp, C.c_int, f64p, u8p],
        "rr_streak_photometry_only": [vp, u8p, C.c_int, vp, f64p],
        "rr_debug_read": [vp, C.c_int, C.c_int, vp, C.c_size_t],
        "rr_solid_angles": [vp, C.c_int, C.c_int, f64p],
        "rr_simulate_particles": [vp, C.POINTER(SimParams), C.c_int64, C.c_int, C.c_int, vp, i32p, C.POINTER(C.c_double)],
        "rr_simulate_records_device": [vp, C.POINTER(SimParams), C.c_int64, C.c_int, C.c_int, f64p, C.c_int, C.c_double, C.c_double,
                                       C.POINTER(vp), i32p, C.POINTER(C.c_double)],
        "rr_set_option": [vp, C.c_char_p, C.c_int],
        "rr_timings": [vp, f32p],
        "rr_kernel_launches": [vp, C.POINTER(C.c_longlong)],
        "rr_stream": [vp, C.POINTER(vp)],
        "rr_synchronize": [vp],
        "rr_host_alloc": [C.POINTER(vp), C.c_size_t],
        "rr_host_free": [vp],
        "rr_host_alloc_flags": [C.POINTER(vp), C.c_size_t, C.c_int],
        "rr_host_link_probe": [C.c_int, C.c_size_t, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)],
        "rr_host_draw_randoms": [C.c_uint32, C.c_int, u8p, i32p, C.c_double, C.c_double, u8p, f64p],
        "rr_host_assemble_batch": [C.c_int, vp, i32p, vp, C.c_int, C.c_int, f64p, C.c_int, C.c_double, C.c_double, vp, C.c_int64, i32p, i32p],
        "rr_host_fov_polygon": [vp, C.c_double, C.c_double, C.c_int, C.c_int, f64p, C.POINTER(C.c_int32)],
        "rr_host_load_particles_xml": [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(vp)],
        "rr_host_particles_info": [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int64)],
        "rr_host_particles_copy": [vp, vp, vp],
        "rr_host_png_info": [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)],
        "rr_host_png_read_batch": [C.c_int, vp, vp, u8p, C.c_int, C.c_int, f32p, C.c_int, C.c_int, C.c_int, i32p],
        "rr_host_png_write_batch": [C.c_int, vp, u8p, vp, f32p, C.c_int, C.c_int, C.c_int, C.c_int],
        "rr_host_png_read_batch_u16": [C.c_int, vp, vp, u8p, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, i32p],
        "rr_host_png_write_batch_rgba": [C.c_int, vp, u8p, vp, u8p, C.c_int, C.c_int, C.c_int, C.c_int],
        "rr_host_png_write_batch_u16": [C.c_int, vp, u8p, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int],
        "rr_host_png_write_streams": [C.c_int, vp, u8p, C.c_size_t, vp, C.c_int, C.c_int, C.c_int],
        "rr_host_zlib_decompress_fast": [u8p, C.c_size_t, u8p, C.c_size_t],
        "rr_host_zlib_compress_fast": [u8p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_size_t)],
    }
    for name, args in protos.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.rr_png_stream_bound.argtypes = [C.c_int, C.c_int]
    lib.rr_png_stream_bound.restype = C.c_size_t
    lib.rr_host_free_particles.argtypes = [vp]
    lib.rr_host_free_particles.restype = None
    lib.rr_host_norm2.argtypes = [C.c_int, f64p, f64p, f64p]
    lib.rr_host_norm2.restype = None
    lib.rr_host_tables.argtypes = [f64p, f32p, i32p]
    lib.rr_host_tables.restype = None
    lib.rr_host_sim_physics.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.rr_host_sim_physics.restype = None
    _lib = lib
    return lib


class RainError(RuntimeError):
    pass


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().rr_last_error().decode(errors="replace")
        raise RainError("%s failed (%d): %s" % (what or "librain_b200", status, msg))


def ptr(a):
    """ctypes pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(C.c_void_p)
