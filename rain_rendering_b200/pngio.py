"""Python face of the library's native PNG codec (csrc/rr_host_png.cpp): whole batches are decoded into / encoded
from caller-provided arrays (the page-locked buffers of the frame pipeline) on native threads, one ctypes call
per batch (the GIL is released for its duration)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib

OK, UNSUPPORTED, SIZE = 0, -5, -6


def _paths(paths):
    if paths is None:
        return None, None
    arr = (C.c_char_p * len(paths))(*[None if p is None else os.fsencode(p) for p in paths])
    return arr, C.cast(arr, C.c_void_p)


def info(path):
    """-> (width, height, channels, bit depth) from the IHDR chunk."""
    lib = _lib.load()
    w, h, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    _lib.check(lib.rr_host_png_info(os.fsencode(path), C.byref(w), C.byref(h), C.byref(c), C.byref(d)), "rr_host_png_info")
    return w.value, h.value, c.value, d.value


def read_batch(image_paths, depth_paths, bgr_out, depth_out, n_threads=8):
    """image_paths / depth_paths: lists of n paths (either may be None).  bgr_out (>= n, Hi, Wi, 3) uint8 receives
    cv2.imread(path); depth_out (>= n, Hd, Wd) float32 receives cv2.imread(path, IMREAD_UNCHANGED) / 256.
    -> int32 status per frame: 0, or the reason the frame must take the caller's fallback decoder."""
    lib = _lib.load()
    n = len(image_paths if image_paths is not None else depth_paths)
    keep_i, pi = _paths(image_paths)
    keep_d, pd = _paths(depth_paths)
    Hi, Wi = (bgr_out.shape[1], bgr_out.shape[2]) if bgr_out is not None else (0, 0)
    Hd, Wd = (depth_out.shape[1], depth_out.shape[2]) if depth_out is not None else (0, 0)
    if bgr_out is not None:
        assert bgr_out.dtype == np.uint8 and bgr_out.flags["C_CONTIGUOUS"] and bgr_out.shape[0] >= n and bgr_out.shape[3] == 3
    if depth_out is not None:
        assert depth_out.dtype == np.float32 and depth_out.flags["C_CONTIGUOUS"] and depth_out.shape[0] >= n
    status = np.zeros(max(n, 1), np.int32)
    r = lib.rr_host_png_read_batch(n, pi, pd, _lib.ptr(bgr_out), Wi, Hi, _lib.ptr(depth_out), Wd, Hd, int(n_threads), _lib.ptr(status))
    if r < 0:
        _lib.check(r, "rr_host_png_read_batch")
    return status[:n]


def write_batch(image_paths, bgr, mask_paths, mask, level=1, n_threads=8):
    """Writes bgr[i] (H, W, 3) uint8 as an RGB PNG (what cv2.imwrite(path, bgr[i]) stores) and mask[i] (H, W) float32
    min/max-normalised to 16-bit gray.  -> number of files that could not be written."""
    lib = _lib.load()
    n = len(image_paths if image_paths is not None else mask_paths)
    if n == 0:
        return 0
    keep_i, pi = _paths(image_paths)
    keep_m, pm = _paths(mask_paths)
    ref = bgr if bgr is not None else mask
    H, W = ref.shape[1], ref.shape[2]
    if bgr is not None:
        assert bgr.dtype == np.uint8 and bgr.flags["C_CONTIGUOUS"] and bgr.shape[0] >= n and bgr.shape[1:] == (H, W, 3)
    if mask is not None:
        assert mask.dtype == np.float32 and mask.flags["C_CONTIGUOUS"] and mask.shape[0] >= n and mask.shape[1:] == (H, W)
    r = lib.rr_host_png_write_batch(n, pi, _lib.ptr(bgr), pm, _lib.ptr(mask), W, H, int(level), int(n_threads))
    if r < 0:
        _lib.check(r, "rr_host_png_write_batch")
    return r
