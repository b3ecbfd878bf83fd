"""Python face of the C ABI: one ``RainContext`` per GPU (thin; all rendering is in the library)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .streaks import BIG, STREAK_DTYPE, texture_buckets


def draw_randoms(seed: int, types: np.ndarray, buckets: np.ndarray, noise_std: float, noise_scale: float):
    """The frame's NumPy legacy-RNG draws (np.random.seed(seed); randint per streak; normal per
    non-Big streak), computed by the library's MT19937 mirror.  -> (tex_idx uint8, noise_deg float64)"""
    lib = _lib.load()
    n = len(types)
    types = np.ascontiguousarray(types, dtype=np.uint8)
    buckets = np.ascontiguousarray(buckets, dtype=np.int32)
    tex = np.zeros(n, np.uint8)
    noise = np.zeros(n, np.float64)
    _lib.check(lib.rr_host_draw_randoms(C.c_uint32(int(seed) & 0xFFFFFFFF), n, _lib.ptr(types), _lib.ptr(buckets),
                                        float(noise_std), float(noise_scale), _lib.ptr(tex), _lib.ptr(noise)),
               "rr_host_draw_randoms")
    return tex, noise


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy array (rr_host_alloc)."""

    def __init__(self, shape, dtype, write_combined=False):
        """write_combined: cudaHostAllocWriteCombined -- for buffers the host only ever WRITES sequentially (inputs);
        reading such memory from the CPU is very slow."""
        lib = _lib.load()
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        _lib.check(lib.rr_host_alloc_flags(C.byref(p), max(nbytes, 1), 1 if write_combined else 0), "rr_host_alloc_flags")
        self._p = p
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._p is not None:
            self.array = None
            _lib.load().rr_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def host_link_probe(device=0, mbytes=256, seconds=0.5, write_combined=False, direction=3):
    """rr_host_link_probe -> (h2d GB/s, d2h GB/s) of concurrent page-locked copies on this process's GPU."""
    lib = _lib.load()
    a, b = C.c_double(0), C.c_double(0)
    _lib.check(lib.rr_host_link_probe(int(device), int(mbytes) << 20, float(seconds), 1 if write_combined else 0, int(direction),
                                      C.byref(a), C.byref(b)), "rr_host_link_probe")
    return a.value, b.value


class RainContext:
    """rr_create / rr_set_streak_db / rr_set_camera / rr_render_frames."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self.lib.rr_create(int(device), C.byref(h)), "rr_create")
        self.h = h
        self.device = device
        self.W = self.H = 0
        self.max_batch = 0
        self.db_ratios = None

    def close(self):
        if self.h is not None:
            self.lib.rr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- streak DB ---------------------------------------------------------------------------
    def set_streak_db(self, textures, ratios=None):
        """textures: list of (h, w) uint8 gray arrays (or (h, w, 3) with identical channels) in
        natural-sort order, normalised like DBManager.load_streak_database leaves them."""
        gray = []
        for t in textures:
            t = np.asarray(t)
            if t.ndim == 3:
                assert (t[..., 0] == t[..., 1]).all() and (t[..., 0] == t[..., 2]).all(), "streak textures must be gray"
                t = t[..., 0]
            gray.append(np.ascontiguousarray(t, dtype=np.uint8))
        width = gray[0].shape[1]
        assert all(g.shape[1] == width for g in gray)
        heights = np.array([g.shape[0] for g in gray], dtype=np.int32)
        data = np.concatenate([g.reshape(-1) for g in gray])
        _lib.check(self.lib.rr_set_streak_db(self.h, len(gray), _lib.ptr(heights), width, _lib.ptr(data)), "rr_set_streak_db")
        self.db_heights, self.db_width = heights, width
        self.db_ratios = np.unique(width / heights.astype(np.float64)) if ratios is None else np.asarray(ratios)

    def alloc_streak_db(self, heights, width):
        heights = np.ascontiguousarray(heights, dtype=np.int32)
        _lib.check(self.lib.rr_alloc_streak_db(self.h, len(heights), _lib.ptr(heights), int(width)), "rr_alloc_streak_db")
        self.db_heights, self.db_width = heights, int(width)
        self.db_ratios = np.unique(width / heights.astype(np.float64))

    def streak_db_device_ptr(self):
        p, n = C.c_void_p(), C.c_size_t()
        _lib.check(self.lib.rr_streak_db_device_ptr(self.h, C.byref(p), C.byref(n)), "rr_streak_db_device_ptr")
        return p.value, n.value

    # -- camera ------------------------------------------------------------------------------
    def set_camera(self, W, H, focal_mm=6.0, f_number=6.0, exposure_ms=2.0, gain=20.0, fallrate=25.0,
                   opacity_attenuation=1.0, max_batch=1, focus_plane=6.0, pix_size=4.65e-06, radius=10.0, fov_deg=165.0,
                   render_scale=1):
        cam = _lib.Camera(int(W), int(H), focal_mm / 1000., float(f_number), float(exposure_ms), float(gain),
                          float(focus_plane), float(pix_size), float(radius), float(fov_deg),
                          float(opacity_attenuation), float(fallrate), int(render_scale), 0)
        _lib.check(self.lib.rr_set_camera(self.h, C.byref(cam), int(max_batch)), "rr_set_camera")
        self.W, self.H, self.max_batch, self.render_scale = int(W), int(H), int(max_batch), int(render_scale)
        he, we = C.c_int(), C.c_int()
        _lib.check(self.lib.rr_env_size(self.h, C.byref(he), C.byref(we)), "rr_env_size")
        self.H_env, self.W_env = he.value, we.value

    # -- hot path ----------------------------------------------------------------------------
    def png_stream_bound(self):
        """bytes to reserve per GPU-made PNG stream of this camera's frames (rr_png_stream_bound)"""
        return int(self.lib.rr_png_stream_bound(self.W, self.H))

    def _frame_io(self, bgr, depth, streaks, offsets, out_bgr, out_mask, out_u8, out_idx8, out_u16, out_range, png=None):
        """rr_frame_io of a batch (host arrays).  depth: float32 metres or the uint16 samples of the depth PNG."""
        n = bgr.shape[0]
        rs = self.render_scale
        assert bgr.shape == (n, self.H * rs, self.W * rs, 3) and bgr.dtype == np.uint8
        assert depth.shape == (n, self.H, self.W) and depth.dtype in (np.float32, np.uint16)
        assert streaks.dtype == STREAK_DTYPE
        assert offsets.dtype == np.int32 and offsets.flags["C_CONTIGUOUS"] and len(offsets) == n + 1
        for a, shape, dt in ((out_bgr, (n, self.H, self.W, 3), np.float32), (out_mask, (n, self.H, self.W), np.float32),
                             (out_u8, (n, self.H, self.W, 3), np.uint8), (out_idx8, (n, self.H, self.W), np.uint8),
                             (out_u16, (n, self.H, self.W), np.uint16), (out_range, (n, 2), np.float64)):
            assert a is None or (a.shape == shape and a.dtype == dt), "output array %s %s, expected %s %s" % (a.shape, a.dtype, shape, dt)
        p = lambda a: None if a is None else _lib.ptr(a).value
        io = _lib.FrameIO(p(bgr), p(depth), _lib.DEPTH_U16_256 if depth.dtype == np.uint16 else _lib.DEPTH_F32_M, 0, p(streaks), p(offsets),
                          p(out_bgr), p(out_mask), p(out_u8), p(out_idx8), p(out_u16), p(out_range), None, None, None, None, 0)
        if png is not None:
            # png: dict(image=(n, stride) uint8 or None, mask=..., image_sizes=(n,) uint32, mask_sizes=...): the saved files' image
            # data as finished zlib streams (rr_frame_io.out_png_*), page-locked host arrays
            stride = None
            for k in ("image", "mask"):
                a = png.get(k)
                if a is None:
                    continue
                sz = png[k + "_sizes"]
                assert a.dtype == np.uint8 and a.ndim == 2 and a.shape[0] >= n and a.flags["C_CONTIGUOUS"] and sz.dtype == np.uint32 and len(sz) >= n
                assert stride in (None, a.shape[1]), "image and mask streams share one stride"
                stride = a.shape[1]
                setattr(io, "out_png_" + k, p(a))
                setattr(io, "out_png_%s_sizes" % k, p(sz))
            io.png_stride = stride or 0
        return n, io

    def render_frames(self, bgr, depth, streaks, offsets, out_bgr=None, out_mask=None, out_u8=None,
                      want=("bgr", "mask", "u8"), out_idx8=None, out_u16=None, out_range=None, png=None):
        """bgr (n,H,W,3) uint8; depth (n,H,W) float32 metres or uint16 PNG samples; streaks STREAK_DTYPE (concatenated);
        offsets (n+1,) int32.  ``want`` names the outputs to allocate when no array is passed: bgr (float32), mask
        (float32), u8, idx8 (plt.imsave's colormap index of the mask), u16 (16-bit normalised mask), range ((min, max) of
        the float64 mask).  Returns a dict of the outputs (numpy arrays)."""
        n = bgr.shape[0]
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        if out_bgr is None and "bgr" in want:
            out_bgr = np.empty((n, self.H, self.W, 3), np.float32)
        if out_mask is None and "mask" in want:
            out_mask = np.empty((n, self.H, self.W), np.float32)
        if out_u8 is None and "u8" in want:
            out_u8 = np.empty((n, self.H, self.W, 3), np.uint8)
        if out_idx8 is None and "idx8" in want:
            out_idx8 = np.empty((n, self.H, self.W), np.uint8)
        if out_u16 is None and "u16" in want:
            out_u16 = np.empty((n, self.H, self.W), np.uint16)
        if out_range is None and "range" in want:
            out_range = np.empty((n, 2), np.float64)
        n, io = self._frame_io(bgr, depth, streaks, offsets, out_bgr, out_mask, out_u8, out_idx8, out_u16, out_range, png)
        _lib.check(self.lib.rr_render_frames_io(self.h, n, C.byref(io)), "rr_render_frames_io")
        return dict(bgr=out_bgr, mask=out_mask, u8=out_u8, idx8=out_idx8, u16=out_u16, range=out_range)

    def submit_frames(self, bgr, depth, streaks, offsets, out_bgr=None, out_mask=None, out_u8=None, out_idx8=None, out_u16=None,
                      out_range=None, png=None):
        """Asynchronous rr_submit_frames_io: all arrays (page-locked for the copies to overlap) must stay
        alive and untouched until the matching ``wait_frames``; at most two batches in flight."""
        n, io = self._frame_io(bgr, depth, streaks, offsets, out_bgr, out_mask, out_u8, out_idx8, out_u16, out_range, png)
        _lib.check(self.lib.rr_submit_frames_io(self.h, n, C.byref(io)), "rr_submit_frames_io")

    def wait_frames(self):
        _lib.check(self.lib.rr_wait_frames(self.h), "rr_wait_frames")

    # -- stage entry points (parity tests) -----------------------------------------------------
    def fog_only(self, bgr, depth):
        n = bgr.shape[0]
        out = np.empty((n, 3, self.H, self.W), np.float64)
        _lib.check(self.lib.rr_fog_only(self.h, n, _lib.ptr(bgr), _lib.ptr(depth), _lib.ptr(out)), "rr_fog_only")
        return out

    def envmap_only(self, planar_bgr):
        planar_bgr = np.ascontiguousarray(planar_bgr, dtype=np.float64)
        n = planar_bgr.shape[0]
        out = np.empty((n, self.H_env, self.W_env, 3), np.uint8)
        _lib.check(self.lib.rr_envmap_only(self.h, n, _lib.ptr(planar_bgr), _lib.ptr(out)), "rr_envmap_only")
        return out

    def streak_photometry_only(self, env_u8, streaks):
        env_u8 = np.ascontiguousarray(env_u8, dtype=np.uint8)
        assert env_u8.shape == (self.H_env, self.W_env, 3)
        out = np.empty((len(streaks), 3), np.float64)
        _lib.check(self.lib.rr_streak_photometry_only(self.h, _lib.ptr(env_u8), len(streaks), _lib.ptr(streaks), _lib.ptr(out)),
                   "rr_streak_photometry_only")
        return out

    def debug_read(self, what: str, frame: int = 0, count: int | None = None):
        np_img, np_env = self.H * self.W, self.H_env * self.W_env
        spec = dict(fog=((3, self.H, self.W), np.float64), rainy=((3, self.H, self.W), np.float64),
                    env=((self.H_env, self.W_env, 3), np.uint8), omega=((self.H_env, self.W_env), np.float64),
                    env_src=((self.H_env, self.W_env), np.int32), fext=((self.H, self.W), np.float32),
                    plans=((count or 0,), _lib.PLAN_DTYPE), arena=((count or 0,), np.float64))[what]
        out = np.zeros(spec[0], spec[1])
        _lib.check(self.lib.rr_debug_read(self.h, _lib.DBG[what], frame, _lib.ptr(out), out.nbytes), "rr_debug_read")
        return out

    # -- on-the-fly particle simulation (stand-in for the closed AHLSimulation binary) ------------
    def simulate_particles(self, first_frame, n_frames, W, H, fallrate, focal_mm=6.0, pix_size_um=4.65, exposure_ms=2.0,
                           sim_hz=2000.0, cam_speed_kmh=0.0, z_near=0.25, z_far=15.0, d_min_mm=0.1, d_max_mm=10.0,
                           min_width_px=1.0, seed=0, max_per_frame=None):
        """-> (list of SIM_STREAK_DTYPE arrays (one per frame, pid order), expected candidates per frame)."""
        from .streaks import SIM_STREAK_DTYPE
        p = _lib.SimParams(int(W), int(H), focal_mm / 1000., pix_size_um * 1e-6, float(exposure_ms), float(fallrate), float(sim_hz),
                           float(cam_speed_kmh), float(z_near), float(z_far), float(d_min_mm), float(d_max_mm), float(min_width_px), int(seed))
        exp = C.c_double(0)
        cap = int(max_per_frame or 1 << 16)
        out = np.zeros((n_frames, cap), SIM_STREAK_DTYPE)
        counts = np.zeros(n_frames, np.int32)
        _lib.check(self.lib.rr_simulate_particles(self.h, C.byref(p), int(first_frame), int(n_frames), cap, _lib.ptr(out), _lib.ptr(counts),
                                                  C.byref(exp)), "rr_simulate_particles")
        frames = []
        for f in range(n_frames):
            r = out[f, :counts[f]]
            frames.append(r[np.argsort(r["pid"], kind="stable")].copy())
        return frames, exp.value

    def _sim_params(self, W, H, fallrate, focal_mm, pix_size_um, exposure_ms, sim_hz, cam_speed_kmh, z_near, z_far, d_min_mm, d_max_mm,
                    min_width_px, seed):
        return _lib.SimParams(int(W), int(H), focal_mm / 1000., pix_size_um * 1e-6, float(exposure_ms), float(fallrate), float(sim_hz),
                              float(cam_speed_kmh), float(z_near), float(z_far), float(d_min_mm), float(d_max_mm), float(min_width_px), int(seed))

    def simulate_records_device(self, first_frame, n_frames, W, H, fallrate, db_ratios, render_scale=1, focal_mm=6.0, pix_size_um=4.65,
                                exposure_ms=2.0, sim_hz=2000.0, cam_speed_kmh=0.0, z_near=0.25, z_far=15.0, d_min_mm=0.1, d_max_mm=10.0,
                                min_width_px=1.0, seed=0):
        """rr_simulate_records_device: frames first_frame .. of the on-the-fly simulation as finished streak records that never
        leave the device (W, H: sensor size before render_scale).  -> (device pointer of the concatenated rr_streak_rec, int32
        offsets (n_frames + 1, host), expected candidates per frame).  The pointer stays valid until the next call."""
        p = self._sim_params(W, H, fallrate, focal_mm, pix_size_um, exposure_ms, sim_hz, cam_speed_kmh, z_near, z_far, d_min_mm, d_max_mm,
                             min_width_px, seed)
        ratios = np.ascontiguousarray(db_ratios, np.float64)
        ptr, exp = C.c_void_p(), C.c_double(0)
        offs = np.zeros(n_frames + 1, np.int32)
        _lib.check(self.lib.rr_simulate_records_device(self.h, C.byref(p), int(first_frame), int(n_frames), int(render_scale), _lib.ptr(ratios),
                                                       len(ratios), 0.0, 0.0, C.byref(ptr), _lib.ptr(offs), C.byref(exp)), "rr_simulate_records_device")
        return ptr.value, offs, exp.value

    def render_frames_device(self, d_bgr, d_depth, depth_u16, d_records, offsets, d_out_bgr=None, d_out_mask=None, d_out_u8=None,
                             d_out_idx8=None, d_out_u16=None, d_out_range=None, sync=True):
        """rr_render_frames_device_io: every pointer is a DEVICE address (int), offsets a host int32 array."""
        offsets = np.ascontiguousarray(offsets, np.int32)
        io = _lib.FrameIO(d_bgr, d_depth, _lib.DEPTH_U16_256 if depth_u16 else _lib.DEPTH_F32_M, 0, d_records, _lib.ptr(offsets).value,
                          d_out_bgr, d_out_mask, d_out_u8, d_out_idx8, d_out_u16, d_out_range, None, None, None, None, 0)
        _lib.check(self.lib.rr_render_frames_device_io(self.h, len(offsets) - 1, C.byref(io), 1 if sync else 0), "rr_render_frames_device_io")

    def set_option(self, name: str, value: int):
        _lib.check(self.lib.rr_set_option(self.h, name.encode(), int(value)), "rr_set_option")

    def timings(self):
        ms = np.zeros(len(_lib.T_NAMES), np.float32)
        _lib.check(self.lib.rr_timings(self.h, _lib.ptr(ms)), "rr_timings")
        return dict(zip(_lib.T_NAMES, ms.tolist()))

    def kernel_launches(self) -> int:
        v = C.c_longlong()
        _lib.check(self.lib.rr_kernel_launches(self.h, C.byref(v)), "rr_kernel_launches")
        return v.value

    def synchronize(self):
        _lib.check(self.lib.rr_synchronize(self.h), "rr_synchronize")


class RainLanes:
    """Several independent contexts ("lanes") on ONE GPU that take batches in turn.  Each lane has its own streams, device
    buffers and patch arena, so the kernels of batch i + 1 run beside those of batch i: the float64-bound frame stages of one
    batch fill the issue slots the integer / load-bound streak stages of another leave, and no stage's tail runs alone.
    Measured on C2, end to end from page-locked host buffers: 15.4 k frames/s through one context with two submissions in flight,
    17.2 k through two lanes with two each (profiles/r02d_lanes_e2e.txt).  Frames are independent (generator.py:318 reseeds per
    frame), so the results are those of a single context, bit for bit (tests/test_parity_gpu.py).

    The submission interface is RainContext's with a deeper queue: ``submit_frames`` / ``render_frames_device(sync=False)`` go
    to the lanes round-robin, ``wait_frames`` retires the OLDEST submission; at most ``capacity`` submissions may be in flight.
    A C caller does the same with ``lanes`` rr_context handles."""

    def __init__(self, device: int = 0, lanes: int = 2, context_factory=None):
        assert lanes >= 1
        make = context_factory or RainContext
        self.ctxs = [make(device) for _ in range(lanes)]
        self.device = device
        self.turn = 0                     # lane of the next submission
        self.fifo = []                    # lanes of the submissions in flight, oldest first
        self.per_lane = 2                 # rr_submit_frames keeps at most two batches in flight per context

    @property
    def capacity(self) -> int:
        return self.per_lane * len(self.ctxs)

    @property
    def inflight(self) -> int:
        return len(self.fifo)

    def close(self):
        for c in self.ctxs:
            c.close()

    # -- configuration: every lane gets the same streak DB and camera -----------------------------------------
    def set_streak_db(self, textures, ratios=None):
        for c in self.ctxs:
            c.set_streak_db(textures, ratios)

    def set_camera(self, *a, **k):
        for c in self.ctxs:
            c.set_camera(*a, **k)
        self.W, self.H, self.max_batch = self.ctxs[0].W, self.ctxs[0].H, self.ctxs[0].max_batch

    def set_option(self, name: str, value: int):
        for c in self.ctxs:
            c.set_option(name, value)

    @property
    def db_ratios(self):
        return self.ctxs[0].db_ratios

    # -- submissions ---------------------------------------------------------------------------------------------
    def _next_lane(self):
        if len(self.fifo) >= self.capacity:
            raise _lib.RainError("RainLanes: %d submissions in flight, wait_frames() first" % len(self.fifo))
        k = self.turn
        self.turn = (k + 1) % len(self.ctxs)
        return k

    def submit_frames(self, *a, **kw):
        """RainContext.submit_frames on the next lane (page-locked arrays, alive and untouched until their wait_frames)."""
        k = self._next_lane()
        self.ctxs[k].submit_frames(*a, **kw)
        self.fifo.append(k)

    def wait_frames(self):
        """Blocks until the OLDEST submission is complete (its outputs are in the caller's arrays)."""
        if not self.fifo:
            raise _lib.RainError("RainLanes.wait_frames: nothing in flight")
        k = self.fifo.pop(0)
        self.ctxs[k].wait_frames()

    def render_frames(self, *a, **kw):
        """Synchronous render on the next lane (drains nothing else)."""
        k = self.turn
        self.turn = (k + 1) % len(self.ctxs)
        return self.ctxs[k].render_frames(*a, **kw)

    def synchronize(self):
        while self.fifo:
            self.wait_frames()
        for c in self.ctxs:
            c.synchronize()

    def kernel_launches(self) -> int:
        return sum(c.kernel_launches() for c in self.ctxs)


def assemble_frame_records(sim_frame: np.ndarray, W: int, H: int, db_ratios, seed: int, noise_std: float = 0.0,
                           noise_scale: float = 0.0, mutate: bool = True) -> np.ndarray:
    """One image frame's records from a simulator frame (STREAK_DTYPE array, XML order):
    in-frame filter (generator.py:413-420), RNG draws, wind-noise write-back.  With
    ``mutate`` the write-back also lands in ``sim_frame`` (the reference mutates its shared Streak
    objects, so a simulator frame reused by a later image frame sees the rotated positions)."""
    from .streaks import apply_wind_noise, in_frame
    m = in_frame(sim_frame, W, H)
    rec = sim_frame[m].copy()
    buckets = texture_buckets(rec["ratio"], db_ratios)
    tex, noise = draw_randoms(seed, rec["type"], buckets, noise_std, noise_scale)
    rec["tex_idx"] = tex
    rec["noise_deg"] = noise
    apply_wind_noise(rec, noise)
    if mutate:
        sim_frame["ip1m"][m] = rec["ip1m"]
        sim_frame["ip2m"][m] = rec["ip2m"]
    return rec


def assemble_batch(sim_frames, seeds, W: int, H: int, db_ratios, noise_std: float = 0.0, noise_scale: float = 0.0, out=None):
    """Records of a batch of image frames in ONE native call (rr_host_assemble_batch): ``sim_frames[k]`` is the simulator
    frame (STREAK_DTYPE array) image frame k renders, ``seeds[k]`` its np.random.seed value (the frame index,
    generator.py:318).  -> (concatenated records, int32 offsets).  ``out``: optional preallocated STREAK_DTYPE array
    (e.g. a page-locked buffer) the records are written into.  With wind noise (noise_std and noise_scale non-zero) the
    frames depend on each other through the write-back of generator.py:152-161 and are assembled one by one, in
    order, by ``assemble_frame_records``."""
    n = len(sim_frames)
    if noise_std != 0.0 and noise_scale != 0.0:
        recs, offs = [], [0]
        for sim, seed in zip(sim_frames, seeds):
            r = assemble_frame_records(sim, W, H, db_ratios, int(seed), noise_std, noise_scale)
            recs.append(r)
            offs.append(offs[-1] + len(r))
        rec = np.concatenate(recs) if recs else np.zeros(0, STREAK_DTYPE)
        if out is not None:
            out[:len(rec)] = rec
            rec = out[:len(rec)]
        return rec, np.asarray(offs, np.int32)
    lib = _lib.load()
    total = int(sum(len(s) for s in sim_frames))
    if out is None:
        out = np.empty(max(total, 1), STREAK_DTYPE)
    assert out.dtype == STREAK_DTYPE and out.flags["C_CONTIGUOUS"]
    frames = [np.ascontiguousarray(s) for s in sim_frames]
    ptrs = (C.c_void_p * max(n, 1))(*[f.ctypes.data if len(f) else None for f in frames])
    counts = np.array([len(f) for f in frames], np.int32)
    seeds = np.asarray(seeds, np.uint32)
    ratios = np.ascontiguousarray(db_ratios, np.float64)
    offs = np.zeros(n + 1, np.int32)
    _lib.check(lib.rr_host_assemble_batch(n, C.cast(ptrs, C.c_void_p), _lib.ptr(counts), _lib.ptr(seeds), int(W), int(H), _lib.ptr(ratios),
                                          len(ratios), float(noise_std), float(noise_scale), _lib.ptr(out), len(out), _lib.ptr(offs), None),
               "rr_host_assemble_batch")
    return out[:offs[-1]], offs
