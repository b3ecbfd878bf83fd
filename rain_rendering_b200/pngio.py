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
    cv2.imread(path); depth_out (>= n, Hd, Wd) float32 receives cv2.imread(path, IMREAD_UNCHANGED) / 256., uint16 the samples.
    -> int32 status per frame: 0, or the reason the frame must take the caller's fallback decoder."""
    lib = _lib.load()
    n = len(image_paths if image_paths is not None else depth_paths)
    keep_i, pi = _paths(image_paths)
    keep_d, pd = _paths(depth_paths)
    Hi, Wi = (bgr_out.shape[1], bgr_out.shape[2]) if bgr_out is not None else (0, 0)
    Hd, Wd = (depth_out.shape[1], depth_out.shape[2]) if depth_out is not None else (0, 0)
    if bgr_out is not None:
        assert bgr_out.dtype == np.uint8 and bgr_out.flags["C_CONTIGUOUS"] and bgr_out.shape[0] >= n and bgr_out.shape[3] == 3
    fn = lib.rr_host_png_read_batch
    if depth_out is not None:
        assert depth_out.dtype in (np.float32, np.uint16) and depth_out.flags["C_CONTIGUOUS"] and depth_out.shape[0] >= n
        if depth_out.dtype == np.uint16:        # the file's samples; the division by 256 happens on the device (rr_frame_io)
            fn = lib.rr_host_png_read_batch_u16
    status = np.zeros(max(n, 1), np.int32)
    r = fn(n, pi, pd, _lib.ptr(bgr_out), Wi, Hi, _lib.ptr(depth_out), Wd, Hd, int(n_threads), _lib.ptr(status))
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


def write_batch_rgba(image_paths, bgr, mask_paths, mask_idx8, level=1, n_threads=8):
    """The reference's file formats (plt.imsave, common/generator.py:466-467): bgr[i] (H, W, 3) uint8 as an 8-bit RGBA PNG and
    mask_idx8[i] (H, W) uint8 -- the colormap index of the normalised mask -- as an RGBA PNG through matplotlib's viridis
    table.  -> number of files that could not be written."""
    lib = _lib.load()
    n = len(image_paths if image_paths is not None else mask_paths)
    if n == 0:
        return 0
    keep_i, pi = _paths(image_paths)
    keep_m, pm = _paths(mask_paths)
    ref = bgr if bgr is not None else mask_idx8
    H, W = ref.shape[1], ref.shape[2]
    if bgr is not None:
        assert bgr.dtype == np.uint8 and bgr.flags["C_CONTIGUOUS"] and bgr.shape[0] >= n and bgr.shape[1:] == (H, W, 3)
    if mask_idx8 is not None:
        assert mask_idx8.dtype == np.uint8 and mask_idx8.flags["C_CONTIGUOUS"] and mask_idx8.shape[0] >= n and mask_idx8.shape[1:] == (H, W)
    r = lib.rr_host_png_write_batch_rgba(n, pi, _lib.ptr(bgr), pm, _lib.ptr(mask_idx8), W, H, int(level), int(n_threads))
    if r < 0:
        _lib.check(r, "rr_host_png_write_batch_rgba")
    return r


def write_batch_u16(image_paths, bgr, mask_paths, mask_u16, level=1, n_threads=8):
    """Compact files: bgr[i] as an 8-bit RGB PNG, mask_u16[i] (the normalised 16-bit mask) as 16-bit gray."""
    lib = _lib.load()
    n = len(image_paths if image_paths is not None else mask_paths)
    if n == 0:
        return 0
    keep_i, pi = _paths(image_paths)
    keep_m, pm = _paths(mask_paths)
    ref = bgr if bgr is not None else mask_u16
    H, W = ref.shape[1], ref.shape[2]
    if bgr is not None:
        assert bgr.dtype == np.uint8 and bgr.flags["C_CONTIGUOUS"] and bgr.shape[0] >= n and bgr.shape[1:] == (H, W, 3)
    if mask_u16 is not None:
        assert mask_u16.dtype == np.uint16 and mask_u16.flags["C_CONTIGUOUS"] and mask_u16.shape[0] >= n and mask_u16.shape[1:] == (H, W)
    r = lib.rr_host_png_write_batch_u16(n, pi, _lib.ptr(bgr), pm, _lib.ptr(mask_u16), W, H, int(level), int(n_threads))
    if r < 0:
        _lib.check(r, "rr_host_png_write_batch_u16")
    return r


def write_streams(paths, streams, sizes, W, H, n_threads=8):
    """Frames finished zlib streams (made on the GPU: rr_frame_io.out_png_*; streams (n, stride) uint8, sizes (n,) uint32) as
    8-bit RGBA PNG files of W x H pixels.  -> number of files that could not be written."""
    lib = _lib.load()
    n = len(paths)
    if n == 0:
        return 0
    keep, pp = _paths(paths)
    assert streams.dtype == np.uint8 and streams.ndim == 2 and streams.flags["C_CONTIGUOUS"] and streams.shape[0] >= n
    assert sizes.dtype == np.uint32 and len(sizes) >= n
    r = lib.rr_host_png_write_streams(n, pp, _lib.ptr(streams), streams.shape[1], _lib.ptr(sizes), int(W), int(H), int(n_threads))
    if r < 0:
        _lib.check(r, "rr_host_png_write_streams")
    return r


def zlib_compress_fast(data: bytes) -> bytes:
    """The library's run + Huffman deflate encoder (csrc/rr_host_deflate.h) on a byte string -> zlib stream (tests)."""
    import ctypes as C
    lib = _lib.load()
    a = np.frombuffer(data, np.uint8) if len(data) else np.zeros(1, np.uint8)
    out = np.zeros(2 * len(data) + 4096, np.uint8)
    n = C.c_size_t()
    _lib.check(lib.rr_host_zlib_compress_fast(_lib.ptr(np.ascontiguousarray(a)), len(data), _lib.ptr(out), len(out), C.byref(n)),
               "rr_host_zlib_compress_fast")
    return out[:n.value].tobytes()


def zlib_decompress_fast(z: bytes, out_len: int):
    """The library's single-shot inflate (csrc/rr_host_inflate.h) -> bytes, or None when it refuses the stream (tests)."""
    lib = _lib.load()
    a = np.frombuffer(z, np.uint8) if len(z) else np.zeros(1, np.uint8)
    out = np.zeros(max(out_len, 1), np.uint8)
    r = lib.rr_host_zlib_decompress_fast(_lib.ptr(np.ascontiguousarray(a)), len(z), _lib.ptr(out), int(out_len))
    return out[:out_len].tobytes() if r == 0 else None


def viridis_rgb():
    """matplotlib's viridis as plt.imsave applies it, (256, 3) uint8 RGB (csrc/rr_viridis.h, tools/make_viridis_lut.py)."""
    import re
    here = os.path.dirname(os.path.abspath(__file__))
    txt = open(os.path.join(here, "csrc", "rr_viridis.h")).read()
    v = [int(t) for t in re.findall(r"\d+", txt[txt.index("{"):])]
    return np.array(v, np.uint8).reshape(256, 3)
