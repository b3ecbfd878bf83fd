"""common.generator.Generator of the reference (common/generator.py:22-473) driving the B200
library: same constructor, same ``run()`` loops (sequence x weather x frame), same output tree,
same conflict strategies and per-frame seeding -- but frames are rendered in batches by
``rr_submit_frames_io`` instead of streak by streak in Python.

Output files (``RAIN_B200_OUTPUT_FORMAT``):
  * ``reference`` (default): what ``plt.imsave`` writes at generator.py:466-467 -- 8-bit RGBA for the rainy image, and
    the rain mask min/max-normalised through matplotlib's default colormap (viridis) as RGBA; written by the native
    threaded encoder from the device's uint8 image and colormap index (no matplotlib needed);
  * ``compact``: 8-bit RGB image + the normalised mask as 16-bit gray (smaller, lossless on the mask's 16 bits);
  * ``matplotlib``: the per-file ``plt.imsave`` calls themselves (needs matplotlib; serial, slow -- for byte-for-byte
    comparisons with files the reference wrote on the same machine).
Only the bytes that are saved cross the host link: uint8 image + uint16 depth samples in, uint8 image + uint8 mask index
out (9 bytes per pixel instead of 14 with float32 depth and mask).
"""
import os
import sys
import time

import cv2
import numpy as np

from common import my_utils
from common.bad_weather import DBManager, RainRenderer, EnvironmentMapGenerator, FovComputation  # noqa: F401 (API parity)
from rain_rendering_b200 import api as _api
from rain_rendering_b200._lib import RainError as _RainError
from rain_rendering_b200.streaks import STREAK_DTYPE as _STREAK_DTYPE

FOG_ATT = 1
USE_DEPTH_WEIGHTING = 0

OUTPUT_FORMATS = ("reference", "compact", "matplotlib")


class _PinnedPool:
    """Page-locked staging buffers are expensive to make (cudaHostAlloc pins ~1.4 GB for three sets of KITTI-sized batches:
    the better part of a second) and identical from one (sequence, weather) to the next, so they are kept per process and
    handed out again by (shape, dtype); ``free`` on a pooled buffer returns it to the pool."""

    def __init__(self):
        self.free_lists = {}

    def __call__(self, shape, dtype):
        key = (tuple(int(v) for v in shape), np.dtype(dtype).str)
        lst = self.free_lists.setdefault(key, [])
        buf = lst.pop() if lst else _api.PinnedBuffer(shape, dtype)
        return _PooledBuffer(self, key, buf)

    def release_all(self):
        for lst in self.free_lists.values():
            for b in lst:
                b.free()
        self.free_lists = {}


class _PooledBuffer:
    def __init__(self, pool, key, buf):
        self._pool, self._key, self._buf = pool, key, buf
        self.array = buf.array

    def free(self):
        if self._buf is not None:
            self._pool.free_lists.setdefault(self._key, []).append(self._buf)
            self._buf, self.array = None, None


_POOL = _PinnedPool()
_CONTEXTS = {}          # device -> RainContext, kept for the life of the process (a new Generator per sequence reuses it)


def _release_process_state():
    try:
        _POOL.release_all()
        for c in _CONTEXTS.values():
            c.close()
        _CONTEXTS.clear()
    except Exception:
        pass


import atexit as _atexit        # noqa: E402
_atexit.register(_release_process_state)


class _FramePipeline:
    """Decode -> render -> encode with everything overlapped.  Batches go through ``rr_submit_frames_io`` /
    ``rr_wait_frames`` with two sets of page-locked buffers; PNG decoding and encoding run on native threads
    (``rr_host_png_read_batch[_u16]`` / ``rr_host_png_write_batch_*``, one ctypes call per batch, straight into / out of the
    page-locked buffers), driven from two helper threads so that, while batch k renders, batch k+1 is being decoded
    and batch k-1 is being written.  Files the native codec does not handle (or that need the reference's resize,
    generator.py:372-374) take ``fallback_decode`` (OpenCV).  The first batch is rendered synchronously (it sizes the patch
    arena).  ``stats`` accumulates where the calling thread waited (seconds)."""

    def __init__(self, ctx, batch, io_threads, fallback_decode=None, alloc=None, png_level=1, depth_u16=True, out_format="reference",
                 gpu_png=False):
        from concurrent.futures import ThreadPoolExecutor
        assert out_format in OUTPUT_FORMATS
        self.ctx, self.batch, self.io_threads, self.png_level = ctx, batch, max(1, io_threads), png_level
        self.fallback_decode = fallback_decode
        self.depth_u16, self.out_format = depth_u16, out_format
        # reference format only: the two files' image data (filter + deflate) is made on the GPU and arrives as finished zlib
        # streams (rr_frame_io.out_png_*); the host frames them and writes the files
        self.gpu_png = bool(gpu_png) and out_format == "reference"
        rs, W, H = ctx.render_scale, ctx.W, ctx.H
        alloc = alloc or _POOL                        # page-locked, so that the copies overlap the kernels; pooled per process
        # three sets of host buffers: one being decoded into, up to two submitted (rr_submit_frames keeps two batches in
        # flight), the outputs of the oldest being written
        self.sets = []
        for _ in range(3):
            s = dict(bgr=alloc((batch, H * rs, W * rs, 3), np.uint8), depth=alloc((batch, H, W), np.uint16 if depth_u16 else np.float32),
                     recs=None, writes=None, offs=None, paths=[], n_recs=0)
            if self.gpu_png:
                stride = ctx.png_stream_bound()
                s["png_image"], s["png_mask"] = alloc((batch, stride), np.uint8), alloc((batch, stride), np.uint8)
                s["png_image_sizes"], s["png_mask_sizes"] = alloc((batch,), np.uint32), alloc((batch,), np.uint32)
                self.sets.append(s)
                continue
            s["u8"] = alloc((batch, H, W, 3), np.uint8)
            if out_format == "reference":
                s["idx8"] = alloc((batch, H, W), np.uint8)
                s["range"] = alloc((batch, 2), np.float64)
            elif out_format == "compact":
                s["u16"] = alloc((batch, H, W), np.uint16)
            else:
                s["mask"] = alloc((batch, H, W), np.float32)
            self.sets.append(s)
        self._alloc = alloc
        self.pool = ThreadPoolExecutor(max_workers=2)      # one decode call and one encode call at a time; the threads are native
        self.inflight = []            # indices of the sets submitted and not yet waited for, oldest first
        self.turn = 0
        self.first = True
        self.frames_done = 0
        self.pending = None           # (set index, decode future) of the batch pushed last
        self.stats = dict(decode_wait=0.0, assemble=0.0, submit=0.0, gpu_wait=0.0, write_wait=0.0, first_batch=0.0)

    # ---- decode ------------------------------------------------------------------------------------------------
    def _decode_batch(self, si, entries):
        """entries: (image file, depth file, frame index for the records, rgb path, mask path).  Fills the set's input
        arrays, compacted over the frames that could be decoded; returns the surviving entries."""
        from rain_rendering_b200 import pngio
        s = self.sets[si]
        n = len(entries)
        if all(e[0].lower().endswith(".png") and e[1].lower().endswith(".png") for e in entries):
            status = pngio.read_batch([e[0] for e in entries], [e[1] for e in entries], s["bgr"].array, s["depth"].array, self.io_threads)
        else:
            status = np.full(n, -1, np.int32)
        ok = []
        for k, e in enumerate(entries):
            if status[k] != 0:
                bg, depth = self.fallback_decode(e[0], e[1], self.depth_u16)
                if bg is None:
                    continue                              # corrupt depth: the reference skips the frame (generator.py:361-363)
                if bg.shape != s["bgr"].array.shape[1:] or depth.shape != s["depth"].array.shape[1:]:
                    raise ValueError("frame %s: image %s / depth %s differ from the sequence's size (%d x %d, read once from its first "
                                     "frame like generator.py:250-258)" % (e[0], bg.shape, depth.shape, self.ctx.W, self.ctx.H))
                s["bgr"].array[k] = bg
                s["depth"].array[k] = depth
            ok.append(k)
        for j, k in enumerate(ok):
            if j != k:
                s["bgr"].array[j] = s["bgr"].array[k]
                s["depth"].array[j] = s["depth"].array[k]
        return [entries[k] for k in ok]

    # ---- encode ------------------------------------------------------------------------------------------------
    def _write_batch(self, si):
        from rain_rendering_b200 import pngio
        s = self.sets[si]
        paths = s["paths"]
        n = len(paths)
        for d in {os.path.dirname(p) for pair in paths for p in pair}:
            os.makedirs(d, exist_ok=True)
        if self.out_format == "matplotlib":              # the reference's own calls, file by file (generator.py:466-467)
            import matplotlib.pyplot as plt
            for k, (rgb_path, mask_path) in enumerate(paths):
                plt.imsave(rgb_path, s["u8"].array[k][..., ::-1])
                plt.imsave(mask_path, s["mask"].array[k])
            return
        if self.gpu_png:
            bad = pngio.write_streams([p[0] for p in paths], s["png_image"].array, s["png_image_sizes"].array, self.ctx.W, self.ctx.H, self.io_threads)
            bad += pngio.write_streams([p[1] for p in paths], s["png_mask"].array, s["png_mask_sizes"].array, self.ctx.W, self.ctx.H, self.io_threads)
        elif self.out_format == "reference":
            bad = pngio.write_batch_rgba([p[0] for p in paths], s["u8"].array[:n], [p[1] for p in paths], s["idx8"].array[:n], self.png_level, self.io_threads)
        else:
            bad = pngio.write_batch_u16([p[0] for p in paths], s["u8"].array[:n], [p[1] for p in paths], s["u16"].array[:n], self.png_level, self.io_threads)
        if bad:
            raise IOError("%d output files could not be written under %s" % (bad, os.path.dirname(paths[0][0])))

    def _schedule_writes(self, si):
        s = self.sets[si]
        s["writes"] = self.pool.submit(self._write_batch, si)
        self.frames_done += len(s["paths"])

    def _outputs(self, s, n):
        """keyword arguments of render_frames / submit_frames naming this format's output arrays"""
        if self.gpu_png:
            return dict(png=dict(image=s["png_image"].array, mask=s["png_mask"].array, image_sizes=s["png_image_sizes"].array,
                                 mask_sizes=s["png_mask_sizes"].array))
        kw = dict(out_u8=s["u8"].array[:n])
        if self.out_format == "reference":
            kw.update(out_idx8=s["idx8"].array[:n], out_range=s["range"].array[:n])
        elif self.out_format == "compact":
            kw.update(out_u16=s["u16"].array[:n])
        else:
            kw.update(out_mask=s["mask"].array[:n])
        return kw

    def _finish_oldest(self):
        si = self.inflight[0]
        t0 = time.perf_counter()
        try:
            self.ctx.wait_frames()
        except _RainError as e:
            if "arena" not in str(e):
                raise
            # the patch arena overflowed: drain, then render what was in flight synchronously (that call grows it)
            try:
                self.ctx.synchronize()
            except _RainError:
                pass
            for sj in self.inflight:
                t = self.sets[sj]
                n = len(t["paths"])
                self.ctx.render_frames(t["bgr"].array[:n], t["depth"].array[:n], t["recs"].array[:t["n_recs"]], t["offs"], want=(), **self._outputs(t, n))
                self._schedule_writes(sj)
            self.inflight = []
            return
        finally:
            self.stats["gpu_wait"] += time.perf_counter() - t0
        self.inflight.pop(0)
        self._schedule_writes(si)

    # ---- the pipeline ------------------------------------------------------------------------------------------
    def push(self, entries, assemble):
        """Renders the batch pushed before this one, then starts decoding ``entries`` (so that the decode overlaps the
        kernels just submitted).  assemble(list of frame indices, record_buffer) -> (records, offsets) of those frames, in
        order (the wind write-back is stateful); record_buffer(n) returns a page-locked array of at least n records that
        the records are to be written into."""
        nxt = None
        if entries:
            # the decode of this batch starts NOW, before the previous batch is assembled and submitted: the native decoder
            # threads never sit idle while this thread does its per-batch work
            si = self.turn
            self.turn = (si + 1) % len(self.sets)
            while si in self.inflight:                # never overwrite buffers the GPU still reads
                self._finish_oldest()
            nxt = (si, self.pool.submit(self._decode_batch, si, list(entries)))
        self._render_pending(assemble)
        self.pending = nxt

    def _render_pending(self, assemble):
        if self.pending is None:
            return
        si, fut = self.pending
        self.pending = None
        s = self.sets[si]
        t0 = time.perf_counter()
        kept = fut.result()
        t1 = time.perf_counter()
        self.stats["decode_wait"] += t1 - t0
        if s["writes"] is not None:
            s["writes"].result()                      # the previous user of this set's output arrays is on disk
            s["writes"] = None
        t2 = time.perf_counter()
        self.stats["write_wait"] += t2 - t1
        n = len(kept)
        if n == 0:
            return
        def record_buffer(n_upper):
            # the set's page-locked record buffer, grown with head-room when a batch needs more
            if s["recs"] is None or len(s["recs"].array) < n_upper:
                if s["recs"] is not None:
                    s["recs"].free()
                s["recs"] = self._alloc((max(2 * n_upper, 65536),), _STREAK_DTYPE)
            return s["recs"].array

        recs, offs = assemble([e[2] for e in kept], record_buffer)
        assert np.diff(offs).max(initial=0) <= 2 ** 16, "Assert that the number of drops doesn't overpass the uint16 rain_mask capacity"
        s["n_recs"], s["offs"] = len(recs), np.ascontiguousarray(offs, np.int32)
        s["paths"] = [(e[3], e[4]) for e in kept]
        t3 = time.perf_counter()
        self.stats["assemble"] += t3 - t2
        if self.first:
            self.ctx.render_frames(s["bgr"].array[:n], s["depth"].array[:n], s["recs"].array[:s["n_recs"]], s["offs"], want=(), **self._outputs(s, n))
            self.first = False
            self.stats["first_batch"] += time.perf_counter() - t3
            self._schedule_writes(si)
        else:
            self.ctx.submit_frames(s["bgr"].array[:n], s["depth"].array[:n], s["recs"].array[:s["n_recs"]], s["offs"], **self._outputs(s, n))
            self.stats["submit"] += time.perf_counter() - t3
            self.inflight.append(si)
            if len(self.inflight) == 2:
                self._finish_oldest()

    def finish(self, assemble=None):
        self._render_pending(assemble)
        while self.inflight:
            self._finish_oldest()
        t0 = time.perf_counter()
        for s in self.sets:
            if s["writes"] is not None:
                s["writes"].result()
                s["writes"] = None
        self.stats["write_wait"] += time.perf_counter() - t0

    def close(self):
        try:
            self.finish()
        finally:
            self.pool.shutdown(wait=True)
            for s in self.sets:
                for k in ("bgr", "depth", "mask", "u8", "idx8", "u16", "range", "recs", "png_image", "png_mask", "png_image_sizes", "png_mask_sizes"):
                    if s.get(k) is not None:
                        s[k].free()


class Generator:
    def __init__(self, args):
        self.conflict_strategy = args.conflict_strategy
        self.rendering_strategy = args.rendering_strategy
        if args.rendering_strategy is None:
            self.output_root = os.path.join(args.output, args.dataset)
        else:
            self.output_root = os.path.join(args.output, args.dataset + '_' + args.rendering_strategy)
        self.dataset = args.dataset
        self.dataset_root = args.dataset_root
        self.images = args.images
        self.sequences = args.sequences
        self.depth = args.depth
        self.particles = args.particles
        self.weather = args.weather
        self.texture = args.texture
        self.norm_coeff = args.norm_coeff
        self.save_envmap = args.save_envmap
        self.settings = args.settings
        self.calib = args.calib
        self.exposure = args.settings["cam_exposure"]
        self.camera_gain = args.settings["cam_gain"]
        self.focal = args.settings["cam_focal"] / 1000.
        self.f_number = args.settings["cam_f_number"]
        self.focus_plane = args.settings["cam_focus_plane"]
        self.noise_scale = args.noise_scale
        self.noise_std = args.noise_std
        self.opacity_attenuation = args.opacity_attenuation
        self.frame_start = args.frame_start
        self.frame_end = args.frame_end
        self.frame_step = args.frame_step
        self.frames = args.frames
        self.verbose = args.verbose
        self.env_type = 'ours'
        self.irrad_type = 'ambient'
        self.db = None
        self.renderer = None
        self.fov_comp = None
        self.BGR_env_map = None
        self.env_map_xyY = None
        self.solid_angle_map = None
        if self.rendering_strategy is not None:
            # 'white' skips tinting and defocus, 'naive_db' is broken upstream (bad_weather.py:349-360): neither is on the
            # accelerated path, and rendering the full model into a folder named after them would be a silent lie
            raise NotImplementedError("rendering_strategy=%r is not implemented on the B200 path (only the default photometric "
                                      "rendering is); see DESIGN.md 9" % (self.rendering_strategy,))
        self.batch = int(os.environ.get("RAIN_B200_BATCH", "64"))
        self.io_threads = int(os.environ.get("RAIN_B200_IO_THREADS", str(min(64, os.cpu_count() or 8))))
        self.device = int(os.environ.get("LOCAL_RANK", os.environ.get("RAIN_B200_DEVICE", "0")))
        self.output_format = os.environ.get("RAIN_B200_OUTPUT_FORMAT", "reference")
        self.gpu_png = os.environ.get("RAIN_B200_GPU_PNG", "1") != "0"      # reference format: filter + deflate on the device
        if self.output_format not in OUTPUT_FORMATS:
            raise ValueError("RAIN_B200_OUTPUT_FORMAT must be one of %s" % (OUTPUT_FORMATS,))
        self.last_stats = None        # _FramePipeline.stats of the last (sequence, weather) rendered
        self._ctx = None
        self.check_folders()

    def check_folders(self):
        print('Output directory: {}'.format(self.output_root))
        existing = []
        for sequence in self.sequences:
            for w in self.weather:
                out_dir = os.path.join(self.output_root, sequence, w["weather"], '{}mm'.format(w["fallrate"]))
                if os.path.exists(out_dir):
                    existing.append(out_dir)
        if len(existing) != 0 and self.conflict_strategy is None:
            print("\r\nFolders already exist: \n%s" % "\n".join(existing))
            while self.conflict_strategy not in ["overwrite", "skip", "rename_folder"]:
                self.conflict_strategy = input("\r\nWhat strategy to use (overwrite|skip|rename_folder):   ")
        assert (self.conflict_strategy in [None, "overwrite", "skip", "rename_folder"])

    def compute_drop(self, bg, drop_dict, rainy_bg, rainy_mask, rainy_saturation_mask):
        raise NotImplementedError("streaks are rendered in batches on the GPU (rr_render_frames); there is no per-streak CPU path")

    # ------------------------------------------------------------------------------------------
    def _decode(self, image_file, depth_file, depth_u16=False):
        """generator.py:352-381 on the decode side (I/O) for files the native codec does not take: uint8 BGR image (at
        sensor resolution: the render_scale reduction of generator.py:354-355 happens on the device) and the depth as
        float32 metres, or -- ``depth_u16`` -- as the PNG's uint16 samples (divided by 256 on the device)."""
        bg = cv2.imread(image_file)
        if bg is None:
            raise IOError("cannot read image %s" % image_file)
        rs = self.settings["render_scale"]
        if rs not in (1, 2) or bg.shape[0] % rs or bg.shape[1] % rs:
            raise NotImplementedError("render_scale %r: only 1 and an exact factor 2 (cv2's 2x2 area path) are implemented on the device" % (rs,))
        if depth_file.endswith(".png"):
            depth = cv2.imread(depth_file, cv2.IMREAD_UNCHANGED)
            if depth is None:
                print('Missing/Corrupted depth data (%s)' % depth_file)
                return None, None
            if not depth_u16:
                depth = depth.astype(np.float32) / 256.
            elif depth.dtype != np.uint16:
                depth = depth.astype(np.uint16)                  # 8-bit depth files: the same sample values
        elif depth_file.endswith(".npy"):
            depth = np.load(depth_file)
            if depth.dtype != np.float32:
                # the reference keeps the array's own dtype (generator.py:367) and a float64 depth makes its whole extinction
                # stage float64; the device stage is float32 like the PNG path, so anything else would silently change numbers
                raise NotImplementedError("depth %s is %s: only float32 .npy depth maps are supported (the device computes the "
                                          "extinction in float32 like the reference does for PNG depth)" % (depth_file, depth.dtype))
            assert not depth_u16
        else:
            raise Exception("Invalid extension")
        depthHW = np.array([int((depth.shape[0] * self.settings["depth_scale"]) // rs),
                            int((depth.shape[1] * self.settings["depth_scale"]) // rs)])
        if not np.all(depth.shape[:2] == depthHW):
            assert not depth_u16, "resized depth is float32"
            depth = cv2.resize(depth, (depthHW[1], depthHW[0]))
        rh, rw = bg.shape[0] // rs, bg.shape[1] // rs
        assert depth.shape[0] <= rh and depth.shape[1] <= rw, "Depth cannot be larger than the image"
        if (depth.shape[0], depth.shape[1]) != (rh, rw):
            # generator.py:379-381 crops the image to the depth map but keeps the environment-map tables, the streak
            # loader and the in-frame filter at the uncropped size (:265,281,413-420): the reference itself does not
            # get through generate_map with such a frame.  Refused by name rather than rendered at a guessed size.
            raise NotImplementedError("depth map %s (%d x %d after scaling) is smaller than the image (%d x %d): the crop-centre "
                                      "case of generator.py:379-381 is not supported" % (depth_file, depth.shape[1], depth.shape[0], rw, rh))
        return np.ascontiguousarray(bg), np.ascontiguousarray(depth)

    def _probe(self, files, depth_files):
        """Size of the sequence (generator.py:235-258) and the depth form its frames can be staged in."""
        rs = self.settings["render_scale"]
        if "nuscenes" in self.dataset:
            assert depth_files[0].endswith(".npy"), "nuscenes processing only works with .npy for depth"
            if "gan" in self.dataset:
                imW, imH = (1600, 900)                      # HARDCODED in the reference (generator.py:241-243)
            else:
                img = cv2.imread(files[0])                  # any extension (CAM_FRONT frames are .jpg), no render_scale division
                imH, imW = img.shape[0:2]
            if rs != 1:
                raise NotImplementedError("nuscenes with render_scale != 1: the reference sizes its tables without the division "
                                          "(generator.py:244-246) and then cannot render the reduced frame")
        else:
            im = files[0]
            if im.endswith(".png"):
                imH, imW = cv2.imread(im).shape[0:2]
            elif im.endswith(".npy"):
                imH, imW = np.load(im).shape[0:2]
            else:
                raise Exception("Invalid extension", im)
            imH, imW = imH // rs, imW // rs
        # uint16 staging needs PNG depth at the render size (no cv2.resize of the depth, generator.py:372-374)
        depth_u16 = False
        d0 = depth_files[0]
        if d0.lower().endswith(".png"):
            dd = cv2.imread(d0, cv2.IMREAD_UNCHANGED)
            if dd is not None and dd.ndim == 2:
                want = (int((dd.shape[0] * self.settings["depth_scale"]) // rs), int((dd.shape[1] * self.settings["depth_scale"]) // rs))
                depth_u16 = dd.shape[:2] == want
        return imW, imH, depth_u16

    def run(self):
        for folder_idx, sequence in enumerate(self.sequences):
            print('\nSequence: ' + sequence)
            depth_folder = self.depth[sequence]
            for sim_idx, sim_weather in enumerate(self.weather):
                weather, fallrate = sim_weather["weather"], sim_weather["fallrate"]
                out_seq_dir = os.path.join(self.output_root, sequence)
                out_dir = os.path.join(out_seq_dir, weather, '{}mm'.format(fallrate))
                sim_file = self.particles[sequence][sim_idx]
                if os.path.exists(out_dir):
                    if self.conflict_strategy in ("skip", "overwrite"):
                        pass
                    elif self.conflict_strategy == "rename_folder":
                        out_shift = 0
                        while os.path.exists(out_dir + '_copy%05d' % out_shift):
                            out_shift += 1
                        out_dir = out_dir + '_copy%05d' % out_shift
                    else:
                        raise NotImplementedError
                os.makedirs(out_dir, exist_ok=True)
                if "nuscenes" in self.dataset:
                    files = list(self.images[sequence])
                    depth_files = list(self.depth[sequence])
                else:
                    files = my_utils.natsorted([os.path.join(self.images[sequence], p) for p in my_utils.os_listdir(self.images[sequence])])
                    depth_files = my_utils.natsorted([os.path.join(depth_folder, d) for d in my_utils.os_listdir(depth_folder)])
                    files = [f for f in files if os.path.isfile(f)]
                    depth_files = [f for f in depth_files if os.path.isfile(f)]
                imW, imH, depth_u16 = self._probe(files, depth_files)
                print('Simulation: rain {}mm/hr'.format(fallrate))
                self.db = DBManager(streaks_path_xml=sim_file, streaks_path=self.texture, norm_coeff_path=self.norm_coeff)
                self.renderer = RainRenderer(focal=self.focal, f_number=self.f_number, focus_plane=6, radius=10, fov=165)
                self.db.load_streak_database()
                self.db.load_streaks_from_xml(self.dataset, self.settings, [imW, imH], use_pickle=False, verbose=self.verbose)
                frame_render_dict = list(self.db.streaks_simulator.values())
                if self._ctx is None:
                    if self.device not in _CONTEXTS:
                        _CONTEXTS[self.device] = _api.RainContext(self.device)
                    self._ctx = _CONTEXTS[self.device]
                ctx = self._ctx
                ctx.set_streak_db(self.db.streaks_light, self.db.ratio)
                gain = self.camera_gain if self.camera_gain else 20      # generator.py:232-233,260-261
                ctx.set_camera(imW, imH, focal_mm=self.focal * 1000., f_number=self.f_number, exposure_ms=self.exposure, gain=gain,
                               fallrate=fallrate, opacity_attenuation=self.opacity_attenuation, max_batch=self.batch,
                               render_scale=self.settings["render_scale"])
                f_start, f_end, f_step = self.frame_start, self.frame_end, self.frame_step
                f_end = len(files) if f_end is None else min(f_end, len(files))
                if self.frames:
                    idx = np.unique(np.clip(self.frames, 0, f_end - 1)).tolist()
                else:
                    idx = list(range(f_start, f_end, f_step))
                print("{} images".format(len(idx)))
                frames_exist_nb = 0
                t_setup = time.time()
                pipe = _FramePipeline(ctx, self.batch, self.io_threads, fallback_decode=self._decode,
                                      png_level=int(os.environ.get("RAIN_B200_PNG_LEVEL", "1")), depth_u16=depth_u16,
                                      out_format=self.output_format, gpu_png=self.gpu_png)
                t0 = time.time()      # the frame loop proper; what came before is per-(sequence, weather) set-up
                queue = []            # (image file, depth file, frame index for the records, out paths)
                env_jobs = []         # --save_envmap: (frame position in its batch's queue, path)

                def assemble(f_name_indices, record_buffer):
                    # np.random.seed(f_name_idx) + the in-frame filter + the per-streak draws (generator.py:318,413-420,136) for the
                    # whole batch in one native call; the wind write-back (:152-161) makes frames depend on each other, so with
                    # noise they are assembled one by one
                    sims = [frame_render_dict[k % len(frame_render_dict)].records for k in f_name_indices]
                    out = record_buffer(int(sum(len(r) for r in sims)))
                    return _api.assemble_batch(sims, f_name_indices, imW, imH, self.db.ratio, self.noise_std, self.noise_scale, out=out)

                try:
                    for f_idx, i in enumerate(idx):
                        image_file, depth_file = files[i], depth_files[i]
                        if self.dataset == 'nuscenes':
                            render_ix = np.linspace(0, len(frame_render_dict), len(files), endpoint=False, dtype=int)
                            f_name_idx = int(render_ix[i])
                        else:
                            f_name_idx = i
                        assert os.path.exists(image_file), "Image file {} does not exist".format(image_file)
                        assert os.path.exists(depth_file), "Depth file {} does not exist".format(depth_file)
                        file_name = os.path.split(image_file)[-1]
                        out_rainy_path = os.path.join(out_dir, 'rainy_image', '{}.png'.format(file_name[:-4]))
                        out_rainy_mask_path = os.path.join(out_dir, 'rain_mask', '{}.png'.format(file_name[:-4]))
                        if os.path.exists(out_rainy_path) or os.path.exists(out_rainy_mask_path):
                            if self.conflict_strategy == "skip":
                                frames_exist_nb += 1
                                continue
                            elif self.conflict_strategy == "overwrite":
                                pass
                            else:
                                raise NotImplementedError
                        queue.append((image_file, depth_file, f_name_idx, out_rainy_path, out_rainy_mask_path))
                        if self.save_envmap:
                            env_jobs.append((image_file, depth_file, f_name_idx, os.path.join(out_seq_dir, 'envmap', '{}.png'.format(file_name[:-4]))))
                        if len(queue) == self.batch:
                            pipe.push(queue, assemble)
                            queue = []
                            if self.verbose:
                                sys.stdout.write('\r%d/%d frames, %.1f frames/s   ' % (f_idx + 1, len(idx), (f_idx + 1) / (time.time() - t0)))
                    pipe.push(queue, assemble)
                    pipe.finish(assemble)
                finally:
                    pipe.close()
                    self.last_stats = dict(pipe.stats, frames=pipe.frames_done, seconds=time.time() - t0, buffers_s=t0 - t_setup)
                if env_jobs:
                    self._save_envmaps(ctx, env_jobs, depth_u16, frame_render_dict, imW, imH)
                if frames_exist_nb > 0:
                    print("Skipped {}/{} already existing renderings".format(frames_exist_nb, len(idx)))
            print("\n\nEnd of the simulation")

    def _save_envmaps(self, ctx, jobs, depth_u16, frame_render_dict, imW, imH):
        """--save_envmap (generator.py:468-469): a debugging aid, so done the simple way -- each frame once more through the
        synchronous call, the device's uint8 environment map read back (rr_debug_read) and saved as plt.imsave would save
        BGR_env_map[..., ::-1], a float RGB array k / 255: (k / 255.0 * 255).astype(uint8), RGBA."""
        from rain_rendering_b200 import pngio
        lut = (np.arange(256) / 255.0 * 255).astype(np.uint8)
        for image_file, depth_file, f_name_idx, path in jobs:
            bg, depth = self._decode(image_file, depth_file, depth_u16)
            if bg is None:
                continue
            sim = frame_render_dict[f_name_idx % len(frame_render_dict)].records
            recs, offs = _api.assemble_batch([sim], [f_name_idx], imW, imH, self.db.ratio, 0.0, 0.0)
            ctx.render_frames(bg[None], depth[None], recs, offs, want=("u8",))
            env = lut[ctx.debug_read("env", 0)]
            os.makedirs(os.path.dirname(path), exist_ok=True)
            if pngio.write_batch_rgba([path], np.ascontiguousarray(env[None]), None, None, 1, 1):
                raise IOError("cannot write %s" % path)
