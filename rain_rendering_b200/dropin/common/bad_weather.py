"""common.bad_weather of the reference, re-hosted on the B200 library.

Kept importable with the reference's names and call signatures:
``DBManager``, ``RainRenderer``, ``FovComputation``, ``EnvironmentMapGenerator``, ``DropType``,
``Streak``, ``Frame``.  The accelerated unit is the frame batch (``Generator.run`` ->
``rr_render_frames``); the reference's per-streak *rendering* entry points whose work now happens inside
the CUDA kernels (``add_drop_to_image``, ``circle_of_confusion``, ``make_rain_layer``) raise
``NotImplementedError`` naming the call that replaces them -- there is no CPU rendering path here.
The per-streak *geometry* calls (``compute_circle``, ``warping_points``,
``FovComputation.compute_fov_plane_points``) answer with the reference's values: the polygon through
``rr_host_fov_polygon`` -- the header code ``k_plan`` compiles for the device, run on the host.
"""
import os
from enum import Enum

import cv2
import numpy as np

from common import my_utils
from rain_rendering_b200 import streaks as _S

cache = {}


class DropType(Enum):
    Big = 0
    Medium = 1
    Small = 2


class Streak:
    """Field-for-field the reference's Streak (common/bad_weather.py:46-60)."""

    def __init__(self):
        self.pid = None
        self.world_position_start = None
        self.world_position_end = None
        self.world_diameter_start = None
        self.world_diameter_end = None
        self.image_position_start = None
        self.image_position_end = None
        self.image_diameter_start = None
        self.image_diameter_end = None
        self.ratio = None
        self.max_width = None
        self.length = None
        self.drop_type = None

    def __repr__(self):
        return str(self.__dict__).replace(',', '\n')


class Frame:
    def __init__(self):
        self.id = None
        self.starting_time = None
        self.exposure_time = None
        self.streaks_count = None
        self.streaks = None
        self.records = None      # STREAK_DTYPE array of this simulator frame (B200 path)

    def __repr__(self):
        return str({k: v for k, v in self.__dict__.items() if k != "records"}).replace(',', '\n')


class DBManager:
    def __init__(self, streaks_path=None, streaks_path_xml=None, norm_coeff_path=None):
        self.streaks_path = streaks_path
        self.streaks_path_xml = streaks_path_xml
        self.streaks_light = []
        self.norm_coeff_path = norm_coeff_path
        self.streaks_simulator = {}
        self.ratio = np.array([])

    @staticmethod
    def classify_drop(w):
        if w >= 4:
            return DropType(0)
        if w > 1:
            return DropType(1)
        return DropType(2)

    def load_streak_database(self):
        """common/bad_weather.py:108-146: normalised uint8 BGR textures in natural-sort order."""
        if not os.path.exists(self.streaks_path):
            print("No existing path for streak database (", self.streaks_path, ")")
            exit(-1)
        norm_coeffs, coeff = {}, None
        with open(self.norm_coeff_path, 'r') as f:
            for line in f.readlines():
                if line[:2] == 'cv':
                    coeff = int(line[2:])
                    continue
                norm_coeffs[coeff] = [float(v) for v in line.split('\n')[0].split(' ')[:-1]]
        tmp, ratios = [], []
        for file_name in my_utils.os_listdir(self.streaks_path):
            name = os.path.splitext(file_name)[0]
            coeff, osc = name.split('_')
            coeff = int(coeff[-1:]) if len(coeff) == 3 else int(coeff[-2:])
            osc = int(osc[-1:])
            img = cv2.imread(os.path.join(self.streaks_path, file_name), cv2.IMREAD_ANYDEPTH)
            img = cv2.cvtColor(img, cv2.COLOR_GRAY2BGR)
            tmp.append(((255.0 * norm_coeffs[coeff][osc] * img) / 65535.0).astype(np.uint8))
            ratios.append(tmp[-1].shape[1] / tmp[-1].shape[0])
        self.ratio = np.unique(np.array(ratios))
        self.streaks_light = tmp       # ragged list (the reference relies on numpy<1.24 ragged arrays)

    def load_streaks_from_xml(self, dataset, settings, image_shape_WH, use_pickle=True, verbose=True):
        """common/bad_weather.py:148-248 -> self.streaks_simulator {frame id: Frame}; every Frame also
        carries the packed records the GPU consumes."""
        print('Reading particles file {}'.format(self.streaks_path_xml))
        if not os.path.exists(self.streaks_path_xml):
            my_utils.print_error("No existing path for XML file (" + self.streaks_path_xml + ")")
            exit(-1)
        if dataset == 'nuscenes_gan':
            raise NotImplementedError("nuscenes_gan rescaling (bad_weather.py:213-219) is out of scope")
        try:
            # use_pickle: the reference's cache of the parsed simulation (:155-178), here a binary file of packed records
            frames = _S.load_streaks_from_xml(self.streaks_path_xml, settings["render_scale"], image_shape_WH[0], image_shape_WH[1],
                                              use_cache=bool(use_pickle) or os.environ.get("RAIN_B200_PARTICLES_CACHE", "0") == "1")
        except Exception:
            raise Exception("Reading XML file {} crashed, which is likely due to corrupted particles simulation files. If so, delete this simulation folder manually and re-run to allow generation of new simulation.".format(self.streaks_path_xml))
        for i, rec in enumerate(frames):
            f = Frame()
            f.id = i
            f.streaks_count = len(rec)
            f.records = rec
            f.streaks = _LazyStreaks(rec, self)
            self.streaks_simulator[f.id] = f

    def take_drop_texture(self, drop):
        b = int(_S.texture_buckets(np.array([drop.ratio]), self.ratio)[0])
        return self.streaks_light[np.random.randint(10 * b, 10 * b + 10)] / 255.0

    @staticmethod
    def normalize(v):
        return v / np.linalg.norm(v)


class _LazyStreaks(dict):
    """pid -> Streak view of a record array, materialised on first access."""

    def __init__(self, rec, db):
        super().__init__()
        self._rec = rec
        self._db = db
        self._done = False

    def _fill(self):
        if self._done:
            return
        self._done = True
        for r in self._rec:
            s = Streak()
            s.pid = int(r["pid"])
            s.world_position_start, s.world_position_end = r["wp1"].copy(), r["wp2"].copy()
            s.image_position_start, s.image_position_end = r["ip1m"].astype(int), r["ip2m"].astype(int)
            s.image_diameter_start, s.image_diameter_end = float(r["iw1"]), float(r["iw2"])
            s.ratio, s.max_width, s.length = float(r["ratio"]), int(r["max_width"]), int(r["length"])
            s.drop_type = DropType(int(r["type"]))
            dict.__setitem__(self, s.pid, s)

    def __len__(self):
        return len(self._rec)

    def __iter__(self):
        self._fill()
        return dict.__iter__(self)

    def items(self):
        self._fill()
        return dict.items(self)

    def values(self):
        self._fill()
        return dict.values(self)

    def __getitem__(self, k):
        self._fill()
        return dict.__getitem__(self, k)


class RainRenderer:
    def __init__(self, focal, f_number, focus_plane, radius, fov):
        self.f = focal
        self.N = f_number
        self.focus_plane = focus_plane
        self.radius = radius
        self.fov = fov

    def compute_circle(self, o, is_infinity=False):
        """Thin-lens blur circle of an object at distance ``o`` (common/bad_weather.py:464-469), in pixels of the
        reference's hard-wired 4.65 um pitch unless ``is_infinity``.  The render path evaluates the same expression per
        streak on the device (rr_plan_patch, csrc/rr_streak_geom.h: sig_y = |c|, sig_x = |c| / 2)."""
        f2 = self.f ** 2
        if is_infinity:
            return f2 / (self.N * o)
        return ((o - self.focus_plane) * f2) / (o * (self.focus_plane - self.f) * self.N) / 4.65e-06

    @staticmethod
    def warping_points(drop, drop_texture, image_width, image_height):
        """Source and destination quads of a Big drop's perspective warp, and the patch's corners in the image
        (common/bad_weather.py:300-328) -> (p1, p2, maxC, minC).  The render path builds the same quad on the device
        (rr_plan_patch)."""
        xs, ys = round(drop.image_position_start[0]), round(drop.image_position_start[1])
        xe, ye = round(drop.image_position_end[0]), round(drop.image_position_end[1])
        ds, de = np.floor(drop.image_diameter_start), np.floor(drop.image_diameter_end)
        lo = np.array([max(min(xs, xe), 0), max(min(ys, ye), 0)])
        hi = np.array([min(max(xs + ds, xe + de), image_width), min(max(ys, ye), image_height)])
        th, tw = drop_texture.shape[:2]
        eps = 0.001                                   # keeps the perspective matrix regular (:313)
        src = np.float32([[0, 0], [tw, 0], [tw, th], [0, th]])
        dst = np.float32([[xs - lo[0], ys - lo[1]], [xs - lo[0] + ds, ys - lo[1]],
                          [xe - lo[0] + de + eps, ye - lo[1]], [xe - lo[0] + eps, ye - lo[1]]])
        return src, dst, hi, lo

    def circle_of_confusion(self, drop, drop_distance, drop_dict):
        raise NotImplementedError("defocus runs inside the CUDA path (k_blur_v/k_blur_h); use Generator.run / RainContext.render_frames")

    def add_drop_to_image(self, *a, **k):
        raise NotImplementedError("per-streak photometry and blending run inside the CUDA path (k_setup, k_composite); "
                                  "use Generator.run / RainContext.render_frames")

    @staticmethod
    def make_rain_layer(*a, **k):
        raise NotImplementedError("rain_layer is never saved by the reference (generator.py:389,438) and is not produced")


class FovComputation:
    def __init__(self, camera):
        self.camera = camera

    def compute_fov_plane_points(self, drop_dict, radius, fov, N, env_shape):
        """Field-of-view polygon of one drop in environment-map pixels (common/bad_weather.py:596-704) ->
        (polygon (n, 2), cone points on the sphere, drop position, drop direction).  The polygon comes from
        ``rr_host_fov_polygon``: the header code k_plan compiles for the device (csrc/rr_streak_geom.h), evaluated on the
        host -- 20 cone rays, camera at the origin (what Generator uses, generator.py:268); other arguments are refused.
        The cone points (unused by the reference's caller, generator.py:175) are returned as an empty (0, 3) array.  A drop
        the reference would skip prints 'Drop skipped' and returns its except-branch tuple (:699-704)."""
        import ctypes as C
        from rain_rendering_b200 import _lib
        if N != 20 or np.any(np.asarray(self.camera, np.float64) != 0):
            raise NotImplementedError("compute_fov_plane_points: the B200 path evaluates N = 20 rays for a camera at the origin")
        w0 = np.asarray(drop_dict.world_position_start, np.float64)
        w1 = np.asarray(drop_dict.world_position_end, np.float64)
        position = (w0 + w1) / 2
        position[1], position[2] = position[2], position[1].copy()        # y <-> z (:599)
        direction = position / np.linalg.norm(position)
        rec = np.zeros(1, _S.STREAK_DTYPE)
        rec["wp1"][0], rec["wp2"][0] = w0, w1
        xy = np.zeros(48)
        n = C.c_int32(0)
        _lib.check(_lib.load().rr_host_fov_polygon(_lib.ptr(rec), float(radius), float(fov), int(env_shape[0]), int(env_shape[1]),
                                                   _lib.ptr(xy), C.byref(n)), "rr_host_fov_polygon")
        if n.value == 0:
            print('Drop skipped')
            return np.array([]), np.array([]), direction, position
        return xy[:2 * n.value].reshape(-1, 2).copy(), np.zeros((0, 3)), position, direction


class EnvironmentMapGenerator:
    """generate_map (common/bad_weather.py:742-819) through rr_envmap_only."""

    def __init__(self, f, image_width, image_height):
        self.image_width = image_width
        self.image_height = image_height
        self.f = f
        self.focal = int(((f * 1000) / 12.7) * image_width)
        self._ctx = None

    def generate_map(self, background):
        from rain_rendering_b200.api import RainContext
        H, W = background.shape[:2]
        if self._ctx is None:
            self._ctx = RainContext(0)
            self._ctx.set_camera(W, H, focal_mm=self.f * 1000., max_batch=1)
        planar = np.ascontiguousarray(np.moveaxis(np.asarray(background, np.float64), -1, 0))[None]
        return self._ctx.envmap_only(planar)[0] / 255.0
