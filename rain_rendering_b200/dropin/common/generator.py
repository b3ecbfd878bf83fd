"""common.generator.Generator of the reference (common/generator.py:22-473) driving the B200
library: same constructor, same ``run()`` loops (sequence x weather x frame), same output tree,
same conflict strategies and per-frame seeding -- but frames are rendered in batches by
``rr_render_frames`` instead of streak by streak in Python.

Differences that are presentation only: PNGs are written with OpenCV (the reference uses
``matplotlib.pyplot.imsave``; matplotlib is used here too when it is importable so the files are
byte-compatible), and progress is printed per batch.
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

FOG_ATT = 1
USE_DEPTH_WEIGHTING = 0


try:
    import matplotlib.pyplot as _plt        # noqa: F401  (identical mask files to the reference when matplotlib exists)
    _HAVE_MATPLOTLIB = True
except Exception:
    _HAVE_MATPLOTLIB = False


def _imsave_rgb(path, bgr_u8):
    cv2.imwrite(path, bgr_u8)


def _imsave_mask(path, mask):
    try:
        import matplotlib.pyplot as plt        # identical to the reference when matplotlib exists
        plt.imsave(path, mask)
        return
    except Exception:
        pass
    lo, hi = float(mask.min()), float(mask.max())
    norm = (mask - lo) / (hi - lo) if hi > lo else np.zeros_like(mask)
    cv2.imwrite(path, (norm * 65535.0 + 0.5).astype(np.uint16))     # 16-bit gray, min/max normalised like imsave


class _FramePipeline:
    """Decode -> render -> encode with everything overlapped.  Batches go through ``rr_submit_frames`` /
    ``rr_wait_frames`` with two sets of page-locked buffers; PNG decoding and encoding run on native threads
    (``rr_host_png_read_batch`` / ``rr_host_png_write_batch``, one ctypes call per batch, straight into / out of the
    page-locked buffers), driven from two helper threads so that, while batch k renders, batch k+1 is being decoded
    and batch k-1 is being written.  Files the native codec does not handle (or that need the reference's resize /
    crop, generator.py:372-381) take ``fallback_decode`` (OpenCV).  The first batch is rendered synchronously (it
    sizes the patch arena)."""

    def __init__(self, ctx, batch, io_threads, fallback_decode=None, alloc=None, png_level=1):
        from concurrent.futures import ThreadPoolExecutor
        self.ctx, self.batch, self.io_threads, self.png_level = ctx, batch, max(1, io_threads), png_level
        self.fallback_decode = fallback_decode
        rs, W, H = ctx.render_scale, ctx.W, ctx.H
        alloc = alloc or _api.PinnedBuffer            # page-locked, so that the copies overlap the kernels
        self.sets = []
        for _ in range(2):
            self.sets.append(dict(bgr=alloc((batch, H * rs, W * rs, 3), np.uint8), depth=alloc((batch, H, W), np.float32),
                                  mask=alloc((batch, H, W), np.float32), u8=alloc((batch, H, W, 3), np.uint8),
                                  writes=None, recs=None, offs=None, paths=[]))
        self.pool = ThreadPoolExecutor(max_workers=2)      # one decode call and one encode call at a time; the threads are native
        self.inflight = []            # indices of the sets submitted and not yet waited for, oldest first
        self.turn = 0
        self.first = True
        self.frames_done = 0
        self.pending = None           # (set index, decode future) of the batch pushed last

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
                bg, depth = self.fallback_decode(e[0], e[1])
                if bg is None:
                    continue                              # corrupt depth: the reference skips the frame (generator.py:361-363)
                if bg.shape != s["bgr"].array.shape[1:] or depth.shape != s["depth"].array.shape[1:]:
                    raise AssertionError("frame size %s differs from the sequence size (%d, %d) (generator.py:250-258 reads it once)" % (bg.shape, self.ctx.W, self.ctx.H))
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
        for d in {os.path.dirname(p) for pair in paths for p in pair}:
            os.makedirs(d, exist_ok=True)
        if _HAVE_MATPLOTLIB:              # byte-compatible mask files need matplotlib's colormap: per-file path
            for k, (rgb_path, mask_path) in enumerate(paths):
                _imsave_rgb(rgb_path, s["u8"].array[k])
                _imsave_mask(mask_path, s["mask"].array[k])
            return
        bad = pngio.write_batch([p[0] for p in paths], s["u8"].array, [p[1] for p in paths], s["mask"].array, self.png_level, self.io_threads)
        if bad:
            raise IOError("%d output files could not be written under %s" % (bad, os.path.dirname(paths[0][0])))

    def _schedule_writes(self, si):
        s = self.sets[si]
        s["writes"] = self.pool.submit(self._write_batch, si)
        self.frames_done += len(s["paths"])

    def _finish_oldest(self):
        si = self.inflight[0]
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
                self.ctx.render_frames(t["bgr"].array[:n], t["depth"].array[:n], t["recs"], t["offs"], None, t["mask"].array[:n], t["u8"].array[:n],
                                       want=("mask", "u8"))
                self._schedule_writes(sj)
            self.inflight = []
            return
        self.inflight.pop(0)
        self._schedule_writes(si)

    # ---- the pipeline ------------------------------------------------------------------------------------------
    def push(self, entries, assemble):
        """Renders the batch pushed before this one, then starts decoding ``entries`` (so that the decode overlaps the
        kernels just submitted).  assemble(frame index) -> records of that frame; it is called in frame order (the wind
        write-back is stateful)."""
        self._render_pending(assemble)
        if not entries:
            return
        si = self.turn
        self.turn ^= 1
        s = self.sets[si]
        if si in self.inflight:                       # never overwrite buffers the GPU still reads
            self._finish_oldest()
        self.pending = (si, self.pool.submit(self._decode_batch, si, list(entries)))

    def _render_pending(self, assemble):
        if self.pending is None:
            return
        si, fut = self.pending
        self.pending = None
        s = self.sets[si]
        kept = fut.result()
        if s["writes"] is not None:
            s["writes"].result()                      # the previous user of this set's output arrays is on disk
            s["writes"] = None
        n = len(kept)
        if n == 0:
            return
        recs, offs = [], [0]
        for e in kept:
            r = assemble(e[2])
            assert len(r) <= 2 ** 16, "Assert that the number of drops doesn't overpass the uint16 rain_mask capacity"
            recs.append(r)
            offs.append(offs[-1] + len(r))
        s["recs"] = np.concatenate(recs)
        s["offs"] = np.ascontiguousarray(offs, np.int32)
        s["paths"] = [(e[3], e[4]) for e in kept]
        if self.first:
            self.ctx.render_frames(s["bgr"].array[:n], s["depth"].array[:n], s["recs"], s["offs"], None, s["mask"].array[:n], s["u8"].array[:n],
                                   want=("mask", "u8"))
            self.first = False
            self._schedule_writes(si)
        else:
            self.ctx.submit_frames(s["bgr"].array[:n], s["depth"].array[:n], s["recs"], s["offs"], None, s["mask"].array[:n], s["u8"].array[:n])
            self.inflight.append(si)
            if len(self.inflight) == 2:
                self._finish_oldest()

    def finish(self, assemble=None):
        self._render_pending(assemble)
        while self.inflight:
            self._finish_oldest()
        for s in self.sets:
            if s["writes"] is not None:
                s["writes"].result()
                s["writes"] = None

    def close(self):
        try:
            self.finish()
        finally:
            self.pool.shutdown(wait=True)
            for s in self.sets:
                for k in ("bgr", "depth", "mask", "u8"):
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
        self.batch = int(os.environ.get("RAIN_B200_BATCH", "16"))
        self.io_threads = int(os.environ.get("RAIN_B200_IO_THREADS", str(min(32, os.cpu_count() or 8))))
        self.device = int(os.environ.get("LOCAL_RANK", os.environ.get("RAIN_B200_DEVICE", "0")))
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
    def _decode(self, image_file, depth_file):
        """generator.py:352-381 on the decode side (I/O): uint8 BGR image (at sensor resolution: the
        render_scale reduction of generator.py:354-355 happens on the device) and float32 depth."""
        bg = cv2.imread(image_file)
        rs = self.settings["render_scale"]
        if rs not in (1, 2) or bg.shape[0] % rs or bg.shape[1] % rs:
            raise NotImplementedError("render_scale %r: only 1 and an exact factor 2 (cv2's 2x2 area path) are implemented on the device" % (rs,))
        if depth_file.endswith(".png"):
            depth = cv2.imread(depth_file, cv2.IMREAD_UNCHANGED)
            if depth is None:
                print('Missing/Corrupted depth data (%s)' % depth_file)
                return None, None
            depth = depth.astype(np.float32) / 256.
        elif depth_file.endswith(".npy"):
            depth = np.load(depth_file).astype(np.float32)
        else:
            raise Exception("Invalid extension")
        depthHW = np.array([int((depth.shape[0] * self.settings["depth_scale"]) // rs),
                            int((depth.shape[1] * self.settings["depth_scale"]) // rs)])
        if not np.all(depth.shape[:2] == depthHW):
            depth = cv2.resize(depth, (depthHW[1], depthHW[0]))
        rh, rw = bg.shape[0] // rs, bg.shape[1] // rs
        assert depth.shape[0] <= rh and depth.shape[1] <= rw, "Depth cannot be larger than the image"
        if (depth.shape[0], depth.shape[1]) != (rh, rw):
            # crop_center of the reduced image (generator.py:379-381) == the same crop at sensor resolution
            x1 = int((rh - depth.shape[0]) / 2) * rs
            y1 = int((rw - depth.shape[1]) / 2) * rs
            bg = bg[x1:x1 + depth.shape[0] * rs, y1:y1 + depth.shape[1] * rs]
        return np.ascontiguousarray(bg), np.ascontiguousarray(depth, dtype=np.float32)

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
                im = files[0]
                if im.endswith(".png"):
                    imH, imW = cv2.imread(im).shape[0:2]
                elif im.endswith(".npy"):
                    imH, imW = np.load(im).shape[0:2]
                else:
                    raise Exception("Invalid extension", im)
                imH, imW = imH // self.settings["render_scale"], imW // self.settings["render_scale"]
                print('Simulation: rain {}mm/hr'.format(fallrate))
                self.db = DBManager(streaks_path_xml=sim_file, streaks_path=self.texture, norm_coeff_path=self.norm_coeff)
                self.renderer = RainRenderer(focal=self.focal, f_number=self.f_number, focus_plane=6, radius=10, fov=165)
                self.db.load_streak_database()
                self.db.load_streaks_from_xml(self.dataset, self.settings, [imW, imH], use_pickle=False, verbose=self.verbose)
                frame_render_dict = list(self.db.streaks_simulator.values())
                if self._ctx is None:
                    self._ctx = _api.RainContext(self.device)
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
                t0 = time.time()
                pipe = _FramePipeline(ctx, self.batch, self.io_threads, fallback_decode=self._decode,
                                      png_level=int(os.environ.get("RAIN_B200_PNG_LEVEL", "1")))
                queue = []            # (image file, depth file, frame index for the records, out paths)

                def assemble(f_name_idx):
                    sim = frame_render_dict[f_name_idx % len(frame_render_dict)]
                    # np.random.seed(f_name_idx) + the per-streak draws + the wind write-back (generator.py:318,136,152-161)
                    return _api.assemble_frame_records(sim.records, imW, imH, self.db.ratio, f_name_idx, self.noise_std, self.noise_scale)

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
                        if len(queue) == self.batch:
                            pipe.push(queue, assemble)
                            queue = []
                            if self.verbose:
                                sys.stdout.write('\r%d/%d frames, %.1f frames/s   ' % (f_idx + 1, len(idx), (f_idx + 1) / (time.time() - t0)))
                    pipe.push(queue, assemble)
                    pipe.finish(assemble)
                finally:
                    pipe.close()
                if frames_exist_nb > 0:
                    print("Skipped {}/{} already existing renderings".format(frames_exist_nb, len(idx)))
            print("\n\nEnd of the simulation")
