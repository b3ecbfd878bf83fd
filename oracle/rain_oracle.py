"""CPU restatement of the rain-rendering hot path (TEST INFRASTRUCTURE -- never shipped).

Every function cites the reference file:line it follows (paths relative to the reference
root, astra-vision/rain-rendering).  The restatement uses the reference's own numeric
libraries (numpy, cv2, scipy.ndimage) at the same call sites so that third-party
behaviour (OpenCV fixed-point resampling, SciPy reflect-mode Gaussian, numpy summation
order) is inherited, not re-guessed.  pyclipper and imutils are absent from the image and
are restated in ``oracle/clipper_rect.py`` / ``rotate_bound`` below (parity unpinned there).

Two arithmetic modes for the two float32 stages of the fog model:

* ``f32_mode="native"``  -- numpy's float32 ``exp`` and OpenCV's float32 ``GaussianBlur``,
  exactly what the reference executes.  Bit-exact against the live reference (pin).
* ``f32_mode="canonical"`` -- the same two stages defined platform-independently: exp is the
  correctly rounded float32 of the real exp, the 25x25 blur accumulates the exact float32
  products in float64 (ascending tap order) and rounds once per pass.  numpy's float32 exp
  is a <=2.52-ULP SIMD routine that changes with the host CPU and OpenCV's float32 filter
  uses FMA/SIMD orderings that cannot be reproduced elsewhere; "canonical" is what the CUDA
  path is held to (tests/test_parity_gpu.py), and tests/test_oracle.py bounds
  canonical-vs-native.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from xml.etree.ElementTree import parse

import cv2
import numpy as np
from scipy.ndimage import gaussian_filter

from . import clipper_rect

BIG, MEDIUM, SMALL = 0, 1, 2


@dataclass
class Camera:
    """Per (sequence, weather) constants.  Reference: common/generator.py:51-55,232-233,267
    and config/*.py settings()."""
    W: int
    H: int
    focal_mm: float = 6.0
    f_number: float = 6.0
    exposure_ms: float = 2.0
    gain: float = 20.0
    fallrate: float = 25
    opacity_attenuation: float = 1.0
    noise_scale: float = 0.0
    noise_std: float = 0.0
    focus_plane: float = 6.0      # hard-coded at common/generator.py:267
    radius: float = 10.0          # idem
    fov_deg: float = 165.0        # idem
    pix_size: float = 4.65e-06    # hard-coded at common/bad_weather.py:469

    @property
    def focal_m(self):
        return self.focal_mm / 1000.0   # common/generator.py:53


# --------------------------------------------------------------------------------------
# row 2: fog-like rain attenuation            (common/add_attenuation.py:26-95)
# --------------------------------------------------------------------------------------

def gauss_kernel_f32_25():
    return cv2.getGaussianKernel(25, 25, cv2.CV_32F).reshape(-1)


def _blur25_f32_canonical(img32: np.ndarray) -> np.ndarray:
    """25x25, sigma 25, BORDER_REFLECT_101 on a float32 image; float64 accumulation of the
    exact float32 products in ascending tap order, one rounding to float32 per pass."""
    k = gauss_kernel_f32_25().astype(np.float64)
    H, W = img32.shape

    def one_pass(a32, axis):
        n = a32.shape[axis]
        idx = np.arange(-12, n + 12)
        idx = np.where(idx < 0, -idx, idx)
        idx = np.where(idx >= n, 2 * (n - 1) - idx, idx)
        pad = np.take(a32, idx, axis=axis).astype(np.float64)
        acc = None
        for t in range(25):
            sl = [slice(None)] * 2
            sl[axis] = slice(t, t + n)
            term = k[t] * pad[tuple(sl)]
            acc = term if acc is None else acc + term
        return acc.astype(np.float32)

    return one_pass(one_pass(img32, 1), 0)


def fog_rain_layer(bg: np.ndarray, depth: np.ndarray, cam: Camera, f32_mode: str = "native") -> np.ndarray:
    """bg: (H,W,3) float64 BGR in [0,1]; depth: (H,W) float32 metres -> (H,W,3) float64.
    Follows FogRain.fog_rain_layer -> calc_l (add_attenuation.py:75-95)."""
    beta_ext = 0.312 * cam.fallrate ** 0.67                               # :43
    x = (-beta_ext) * (depth / 1000)                                      # :48 (float32 when depth is float32)
    if f32_mode == "native" or depth.dtype != np.float32:
        f_ext = np.exp(x)
    else:
        f_ext = np.exp(x.astype(np.float64)).astype(np.float32)
    f_ext3 = np.tile(np.expand_dims(f_ext, axis=-1), (1, 1, 3))           # :49
    exposure_time = cam.exposure_ms * 1e-3                                # :33
    irradiance = (4 * (cam.f_number ** 2) * bg) / (exposure_time * cam.gain * np.pi)   # :53
    irradiance_mean = np.mean(irradiance.reshape(-1, 3), axis=0)          # :70
    g = 0.97
    cos_term = math.cos(math.radians(90))                                 # angle=90 (generator.py:232)
    beta_hg = (1 - (g ** 2)) / (4 * np.pi * ((1 + g ** 2 - 2 * g * cos_term) ** 1.5))   # :64
    l_in = np.clip(beta_hg * irradiance_mean * (1 - f_ext3), 0, 1)        # :71-72
    if f32_mode == "native" or depth.dtype != np.float32:
        f_blur = cv2.GaussianBlur(f_ext3, (25, 25), 25)                   # :79
    else:
        f_blur = np.tile(_blur25_f32_canonical(f_ext)[..., None], (1, 1, 3))
    l_in = cv2.GaussianBlur(l_in, (25, 25), 25)                           # :80
    l = np.clip(bg * f_blur + l_in, 0, 1)                                 # :85-86
    return np.clip(l, 0, 1)                                               # :93


# --------------------------------------------------------------------------------------
# row 3: environment map                        (common/bad_weather.py:707-853)
# --------------------------------------------------------------------------------------

@dataclass
class EnvTables:
    """Geometry-only tables of EnvironmentMapGenerator.generate_map: they depend on (W, H, focal)
    and not on the pixel values, so the whole map is one gather + a blur on the hole pixels."""
    H: int
    W_env: int
    cyl_w: int
    src: np.ndarray        # (H, W_env) int32: flat source pixel index into the H*W image, -1 = black
    written: np.ndarray    # (H, W_env) bool: mask_result != 0  (pixel keeps its gathered value)


def env_focal_px(focal_m: float, W: int) -> int:
    return int(((focal_m * 1000) / 12.7) * W)                             # :712


def build_env_tables(W: int, H: int, focal_m: float) -> EnvTables:
    f = env_focal_px(focal_m, W)
    cx, cy = int(W // 2), int(H // 2)                                     # :745
    max_x = round(f * np.arctan(cx / f) + cx)                             # :730-734
    min_x = round(f * np.arctan(-cx / f) + cx)                            # :736-740
    cyl_w = int(max_x - min_x) + 1                                        # :749
    xx, yy = np.meshgrid(np.linspace(0, W - 1, W), np.linspace(0, H - 1, H))   # :753-755
    hh = xx - cx
    vv = yy - cy
    rowp = np.round((f * (vv / (np.sqrt(hh ** 2 + f ** 2)))) + cy)        # :724-725,760
    colp = np.round((f * np.arctan(hh / f)) + cx) - min_x                 # :726,760-761
    rowp = rowp.astype(np.int32).reshape(-1)
    colp = colp.astype(np.int32).reshape(-1)
    # first writer (lowest flat source index) wins: np.unique(..., return_index=True)  (:762-767)
    src = np.full((H, cyl_w), -1, dtype=np.int64)
    order = np.arange(H * W - 1, -1, -1)
    src[rowp[order], colp[order]] = order        # later assignments (smaller index) overwrite
    mask = src >= 0
    filled = src.copy()
    half = H // 2
    # bottom half first (:776-781, fill table from mask[H//2:] :839-851)
    rows_dn = H - half
    flipped = mask[half:][::-1]                                           # (rows_dn, cyl_w), row 0 = image bottom
    y_fill_dn = np.argmax(flipped, axis=0)
    src_dn = src[H - 1 - y_fill_dn, np.arange(cyl_w)]
    # holes of the centre row (odd H) would break the reference's index bookkeeping
    assert rows_dn == half or mask[half].all(), "centre row of the cylindrical map has holes"
    for y in range(H - half, H):
        holes = ~mask[y]
        filled[y, holes] = src_dn[holes]
    # then the top half (:785-789, table :825-837)
    y_fill_up = np.argmax(mask[:half], axis=0)
    src_up = src[y_fill_up, np.arange(cyl_w)]
    for y in range(half):
        holes = ~mask[y]
        filled[y, holes] = src_up[holes]
    pad = int(cyl_w / 2)                                                  # :791
    W_env = cyl_w + 2 * pad
    src_env = np.full((H, W_env), -1, dtype=np.int64)
    wr_env = np.zeros((H, W_env), dtype=bool)
    src_env[:, pad:pad + cyl_w] = filled
    wr_env[:, pad:pad + cyl_w] = mask
    src_env[:, :pad] = filled[:, :pad][:, ::-1]                           # :797-803
    wr_env[:, :pad] = mask[:, :cyl_w // 2][:, ::-1]
    right = filled[:, cyl_w // 2:][:, ::-1]                               # :806-812
    src_env[:, W_env - right.shape[1]:] = right
    wr_env[:, W_env - right.shape[1]:] = mask[:, cyl_w // 2:][:, ::-1]
    return EnvTables(H=H, W_env=W_env, cyl_w=cyl_w, src=src_env.astype(np.int32), written=wr_env)


def generate_map(rainy_bg: np.ndarray, tab: EnvTables) -> np.ndarray:
    """(H,W,3) float64 BGR -> (H,W_env,3) float64 BGR env map.  generate_map :742-819."""
    bg8 = (rainy_bg * 255).astype(np.uint8).reshape(-1, 3)               # :744
    src = tab.src
    result = np.where((src >= 0)[..., None], bg8[np.maximum(src, 0)], np.uint8(0)).astype(np.uint8)
    blur = cv2.GaussianBlur(result, (15, 15), 0)                          # :815
    out = np.where(tab.written[..., None], result, blur)                  # :816-817 (uint8 wrap-around identity)
    return out / 255.0                                                    # :819


# --------------------------------------------------------------------------------------
# rows 4-5: colour conversion and solid angles (common/my_utils.py:55-85, common/solid_angle.py)
# --------------------------------------------------------------------------------------

_M_RGB2XYZ = np.array([[0.49000, 0.31000, 0.20000], [0.17697, 0.81240, 0.01063], [0.00000, 0.01000, 0.99000]])
_M_XYZ2RGB = np.array([[0.41847, -0.15866, -0.082835], [-0.091169, 0.25243, 0.015708], [0.0009209, -0.0025498, 0.1786]])


def rgb_to_xyY(array: np.ndarray) -> np.ndarray:
    XYZ = np.dot(array, _M_RGB2XYZ) / 0.17697                             # my_utils.py:59
    X, Y, Z = XYZ[..., 0], XYZ[..., 1], XYZ[..., 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        x = X / (X + Y + Z)
        y = Y / (X + Y + Z)
    return np.concatenate([x[..., None], y[..., None], Y[..., None]], axis=-1)


def xyY_to_rgb(xyY: np.ndarray) -> np.ndarray:
    x, y, Y = xyY[..., 0], xyY[..., 1], xyY[..., 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        X = (Y * x) / y                                                   # my_utils.py:77
        Z = (Y * (1 - x - y)) / y                                         # :78
    XYZ = np.concatenate([X[..., None], Y[..., None], Z[..., None]], axis=-1)
    return np.dot(XYZ, _M_XYZ2RGB)                                        # :83


def solid_angles(H_env: int, W_env: int) -> np.ndarray:
    """Per-pixel solid angle of a lat-long map (solid_angle.py:5-29,32-44,66-102)."""
    cols = np.linspace(0, 1, W_env + 1)
    rows = np.linspace(0, 1, H_env + 1)
    u, v = np.meshgrid(cols, rows)
    u = u * 2
    theta = np.pi * (u - 1)
    phi = np.pi * v
    dx = np.sin(phi) * np.sin(theta)
    dy = np.cos(phi)
    dz = -np.sin(phi) * np.cos(theta)

    def stack(sl0, sl1):
        return np.vstack((dx[sl0, sl1].ravel(), dy[sl0, sl1].ravel(), dz[sl0, sl1].ravel()))

    a = stack(slice(None, -1), slice(None, -1))
    b = stack(slice(None, -1), slice(1, None))
    c = stack(slice(1, None), slice(None, -1))
    d = stack(slice(1, None), slice(1, None))

    def tetra(a, b, c):
        ta = np.arccos(np.sum(b * c, 0))
        tb = np.arccos(np.sum(a * c, 0))
        tc = np.arccos(np.sum(a * b, 0))
        ts = (ta + tb + tc) / 2
        product = np.tan(ts / 2) * np.tan((ts - ta) / 2) * np.tan((ts - tb) / 2) * np.tan((ts - tc) / 2)
        product[product < 0] = 0
        return 4 * np.arctan(np.sqrt(product))

    omega = tetra(a, b, c)
    omega += tetra(b, c, d)
    return omega.reshape(H_env, W_env)


# --------------------------------------------------------------------------------------
# rows 7-8: streak database and particles XML  (common/bad_weather.py:108-265)
# --------------------------------------------------------------------------------------

@dataclass
class Streak:
    pid: int
    wp1: np.ndarray
    wp2: np.ndarray
    iw1: float
    iw2: float
    ip1: np.ndarray      # int, (x, y) -- mutated in place by the wind-noise rotation like the reference
    ip2: np.ndarray
    ratio: float
    max_width: int
    length: int
    drop_type: int


def classify_drop(w: int) -> int:
    if w >= 4:                                                            # :100-106
        return BIG
    if w > 1:
        return MEDIUM
    return SMALL


def load_streaks_from_xml(path: str, render_scale: int, W: int, H: int):
    """-> list of frames (XML order), each a list of Streak in XML order.  :185-239"""
    frames = []
    for frame in parse(path).getroot():
        streaks = {}
        for drop in frame:
            a = drop.attrib
            wp1 = np.array(a["wp1"][1:-1].split(";"), dtype=float)
            wp2 = np.array(a["wp2"][1:-1].split(";"), dtype=float)
            ip1 = np.array(a["ip1"][1:-1].split(";"), dtype=float) / render_scale
            ip2 = np.array(a["ip2"][1:-1].split(";"), dtype=float) / render_scale
            iw1 = float(a["iw1"]) / render_scale
            iw2 = float(a["iw2"]) / render_scale
            ip1[1] = H - ip1[1]                                           # :221-222
            ip2[1] = H - ip2[1]
            wp1[2] *= -1                                                  # :223-224
            wp2[2] *= -1
            diff = abs(ip1 - ip2)
            max_width = int(max(iw1, iw2))                                # :226
            with np.errstate(divide="ignore", invalid="ignore"):
                dir2 = diff / np.linalg.norm(diff)
                dir2[1] = -dir2[1]
                cos_theta = np.dot(np.array([0, -1]), dir2)
                ratio = max_width / (diff[1] / cos_theta)                 # :232-233
            ip2 = ip2.round().astype(int)                                 # :234-235
            ip1 = ip1.round().astype(int)
            length = int(np.ceil(np.linalg.norm(ip1 - ip2)).astype(int))  # :236
            if max_width >= 1 and length >= 1:                            # :238
                streaks[int(a["pid"])] = Streak(int(a["pid"]), wp1, wp2, iw1, iw2, ip1, ip2, float(ratio),
                                                max_width, length, classify_drop(max_width))
        frames.append(list(streaks.values()))
    return frames


def filter_in_frame(streaks, W: int, H: int):
    """common/generator.py:413-420."""
    m = max(H, W)
    return [s for s in streaks if 1 <= s.max_width < m and 1 <= s.length < m and
            ((0 <= s.ip1[0] < W and 0 <= s.ip1[1] < H) or (0 <= s.ip2[0] < W and 0 <= s.ip2[1] < H))]


def texture_bucket(ratio: float, ratios: np.ndarray) -> int:
    for i in range(4):                                                    # :251-265
        if ratio < ratios[i]:
            return i
    return 4


# --------------------------------------------------------------------------------------
# row 9: streak patch                           (common/generator.py:119-174)
# --------------------------------------------------------------------------------------

def rotate_bound(image: np.ndarray, angle: float) -> np.ndarray:
    """imutils.rotate_bound restated (imutils absent; see oracle/ref_shims/imutils.py)."""
    (h, w) = image.shape[:2]
    (cX, cY) = (w / 2, h / 2)
    M = cv2.getRotationMatrix2D((cX, cY), -angle, 1.0)
    cos = np.abs(M[0, 0])
    sin = np.abs(M[0, 1])
    nW = int((h * sin) + (w * cos))
    nH = int((h * cos) + (w * sin))
    M[0, 2] += (nW / 2) - cX
    M[1, 2] += (nH / 2) - cY
    return cv2.warpAffine(image, M, (nW, nH))


def warping_points(s: Streak, tex_w: int, tex_h: int, W: int, H: int):
    """RainRenderer.warping_points, common/bad_weather.py:300-329."""
    x0, x1 = round(s.ip1[0]), round(s.ip2[0])
    y0, y1 = round(s.ip1[1]), round(s.ip2[1])
    d0, d1 = np.floor(s.iw1), np.floor(s.iw2)
    minx = max(min(x0, x1), 0)
    miny = max(min(y0, y1), 0)
    maxx = min(max(x0 + d0, x1 + d1), W)
    maxy = min(max(y0, y1), H)
    eps = 0.001
    p1 = np.float32([[0, 0], [tex_w, 0], [tex_w, tex_h], [0, tex_h]])
    p2 = np.float32([[x0 - minx, y0 - miny], [x0 - minx + d0, y0 - miny],
                     [x1 - minx + d1 + eps, y1 - miny], [x1 - minx + eps, y1 - miny]])
    return p1, p2, np.array([maxx, maxy]), np.array([minx, miny])


def make_patch(s: Streak, tex_gray: np.ndarray, cam: Camera, noise: float):
    """-> (gray patch float64 (h,w), minC int (x,y)).  The database textures are gray, so the
    reference's three BGR channels and its alpha (= channel 0, generator.py:174) are one array.
    NB mutates s.ip1 / s.ip2 exactly like generator.py:152-161."""
    W, H = cam.W, cam.H
    tex = tex_gray / 255.0                                                # bad_weather.py:252
    if s.drop_type == BIG:
        p1, p2, maxC, minC = warping_points(s, tex.shape[1], tex.shape[0], W, H)
        shape = np.subtract(maxC, minC).astype(int)
        M = cv2.getPerspectiveTransform(p1, p2)
        drop = cv2.warpPerspective(tex, M, (max(shape[0], 1), max(shape[1], 1)), flags=cv2.INTER_CUBIC)
        return np.clip(drop, 0, 1), minC
    dir1 = s.ip1 - s.ip2
    dir1 = dir1 / np.linalg.norm(dir1)
    theta = np.rad2deg(np.arccos(np.dot(dir1, np.array([0, -1]))))       # generator.py:144
    nx, ny = np.cos(np.deg2rad(noise)), np.sin(np.deg2rad(noise))
    mean_x = (s.ip2[0] + s.ip1[0]) / 2
    mean_y = (s.ip2[1] + s.ip1[1]) / 2
    s.ip1[:] = (s.ip1[0] - mean_x) * nx - (s.ip1[1] - mean_y) * ny + mean_x, \
               (s.ip1[0] - mean_x) * ny + (s.ip1[1] - mean_y) * nx + mean_y
    s.ip2[:] = (s.ip2[0] - mean_x) * nx - (s.ip2[1] - mean_y) * ny + mean_x, \
               (s.ip2[0] - mean_x) * ny + (s.ip2[1] - mean_y) * nx + mean_y
    drop = rotate_bound(tex, theta + noise)                               # :163
    if s.ip2[0] > W // 2:                                                 # :165
        drop = cv2.flip(drop, 0)
    height = max(abs(s.ip2[1] - s.ip1[1]), 2)                             # :166
    width = max(abs(s.ip2[0] - s.ip1[0]), s.max_width + 2)                # :167-168
    drop = cv2.resize(drop, (int(width), int(height)), interpolation=cv2.INTER_AREA)
    return np.clip(drop, 0, 1), s.ip1.copy()                              # :170-171


# --------------------------------------------------------------------------------------
# row 10: streak field-of-view polygon          (common/bad_weather.py:532-704)
# --------------------------------------------------------------------------------------

def _rotation_matrix(axis, theta):
    axis = np.asarray(axis)
    c, s = np.cos(theta), np.sin(theta)
    skv = np.roll(np.roll(np.diag(axis.flatten()), 1, 1), -1, 0)
    return (c * np.identity(3)) + s * (skv - skv.T) + ((1 - c) * np.outer(axis, axis))


def fov_polygon(s: Streak, cam: Camera, env_shape, N: int = 20) -> np.ndarray:
    """compute_fov_plane_points :596-704 -> (20|24, 2) float64 pixel coords, or empty on failure."""
    try:
        with np.errstate(all="ignore"):
            P = np.array((s.wp1 + s.wp2) / 2)
            P[1], P[2] = P[2], P[1].copy()                                # :599
            n = P / np.linalg.norm(P)                                     # camera at origin
            theta = np.deg2rad(cam.fov_deg / 2)
            a, b, c = n[0], n[1], n[2]
            d = np.dot(P, n)
            if b == 0:
                b = 0.001
            px = P[1]
            pz = 0
            py = (-a * px + d - c * pz) / b                               # :613-615
            point = np.array([px, py, pz])
            u = (P - point) / np.linalg.norm(P - point)
            assert np.all(~np.isnan(u))
            rot_vec = np.cross(u, n)
            v = np.dot(n, _rotation_matrix(rot_vec, -theta))
            pts = []
            az_list = []
            R = cam.radius
            for angle in np.arange(0, 2 * np.pi, (2 * np.pi) / N):
                dirv = np.dot(v, _rotation_matrix(n, angle))
                dx, dy, dz = dirv
                x0, y0, z0 = P
                qa = dx * dx + dy * dy + dz * dz
                qb = 2 * dx * (x0 - 0) + 2 * dy * (y0 - 0) + 2 * dz * (z0 - 0)
                qc = 0 * 0 + 0 * 0 + 0 * 0 + x0 * x0 + y0 * y0 + z0 * z0 + -2 * (0 * x0 + 0 * y0 + 0 * z0) - R * R
                disc = qb ** 2 - 4 * qa * qc
                t1 = (-qb + np.sqrt(disc)) / (2 * qa)
                p = P + (t1 * dirv)
                x, y, z = p
                el = np.arctan2(z, np.sqrt(x ** 2 + y ** 2))
                az = np.arctan2(y, x)
                if az < 0:
                    az += 2 * np.pi
                if el < 0:
                    el += 2 * np.pi
                if az > np.pi * 2:
                    az -= 2 * np.pi
                if el > np.pi * 2:
                    el -= 2 * np.pi
                az = ((2 * np.pi - az) - np.pi / 2)                       # :651
                az = az % (2 * np.pi)
                uu = az / (2 * np.pi)
                el = (el + np.pi / 2)
                el = el % (2 * np.pi)
                vv = 1. - el / np.pi
                az_list.append(az)
                pts.append([uu * env_shape[1], vv * env_shape[0]])
            pts = np.array(pts)
            azs = np.array(az_list + [az_list[0]])
            cond = np.bitwise_or(np.isclose(np.diff(azs), 0), np.diff(azs) < 0)   # :669
            count_true = np.sum(cond)
            count_false = np.sum(~cond)
            pos_true = np.where(cond)[0][0]
            pos_false = np.where(~cond)[0][0]
            rows, cols = env_shape[:2]
            if count_true == 1:                                           # :678-684
                final = np.vstack([pts[:pos_true + 1], [cols, pts[pos_true][1]], [cols, 0], [0, 0],
                                   [0, pts[np.mod(pos_true + 1, N)][1]], pts[pos_true + 1:]])
            elif count_false == 1:                                        # :686-692
                final = np.vstack([pts[:pos_false + 1], [0, pts[pos_false][1]], [0, rows], [cols, rows],
                                   [cols, pts[np.mod(pos_false + 1, N)][1]], pts[pos_false + 1:]])
            else:
                final = pts
            return np.array(final)
    except Exception:
        return np.array([])


# --------------------------------------------------------------------------------------
# rows 11-12: photometry, defocus and ordered blend   (common/bad_weather.py:286-298,336-469)
# --------------------------------------------------------------------------------------

def fov_mask(poly: np.ndarray, rows: int, cols: int) -> np.ndarray:
    """pyclipper intersection with the env rectangle, first solution, fillConvexPoly  (:363-390)."""
    solution = clipper_rect.intersect_with_rect([tuple(p) for p in poly], cols, rows)
    s = np.asarray(solution[0]).reshape((-1, 2))                          # IndexError when empty -> streak skipped
    s = np.vstack([s, s[0]])
    m = np.zeros((rows, cols), dtype=np.float64)
    cv2.fillConvexPoly(m, s, 1)
    return m.astype(bool), s


def circle_of_confusion_px(o: float, cam: Camera) -> float:
    f = cam.focal_m
    return ((o - cam.focus_plane) * f ** 2) / (o * (cam.focus_plane - f) * cam.f_number) / cam.pix_size   # :468-469


@dataclass
class FrameState:
    rainy_bg: np.ndarray
    rainy_mask: np.ndarray
    env_xyY: np.ndarray
    omega: np.ndarray
    skipped: list = field(default_factory=list)
    per_streak: list = field(default_factory=list)


def add_drop_to_image(st: FrameState, s: Streak, patch_gray: np.ndarray, minC, poly, cam: Camera):
    """RainRenderer.add_drop_to_image default branch (:361-462).  Raises on a degenerate polygon
    (the reference's try/except then skips the streak, generator.py:180-189)."""
    exposure_time = cam.exposure_ms / 1000.                               # :344
    rows, cols = st.env_xyY.shape[:2]
    mask_env, _ = fov_mask(poly, rows, cols)
    d_avg = (s.iw1 + s.iw2) / 2.                                          # :376
    drop = np.dstack([patch_gray, patch_gray, patch_gray, patch_gray])    # generator.py:174
    drop_xyY = rgb_to_xyY(drop[..., :3])                                  # :379
    drop_xyY[np.isnan(drop_xyY)] = 0
    fov_sa = st.omega[mask_env].copy()                                    # :393
    fov_env = st.env_xyY[mask_env].copy()
    fov_xyY = (fov_env * np.expand_dims(fov_sa, axis=-1)).sum(axis=0)     # :395
    with np.errstate(divide="ignore", invalid="ignore"):
        fov_xy_avg = fov_xyY[:2] / (np.sum(fov_sa))                       # :397
    col = drop_xyY.copy()
    col[..., :2] = fov_xy_avg
    ambient = np.sum(st.env_xyY[..., 2] * st.omega) / np.sum(st.omega)    # :403-404
    avg_fov_lum = fov_xyY[..., 2] / np.sum(st.omega)                      # :407
    drop_Y = 0.94 * avg_fov_lum + 0.06 * ambient                          # :408
    col[..., 2] *= drop_Y
    bgr = xyY_to_rgb(col)[..., ::-1]                                      # :411-412
    sel = drop[..., 3] > 0
    drop[..., :3][sel] = bgr[sel]                                         # :413
    # defocus (:286-298)
    c = abs(circle_of_confusion_px(abs(s.wp1[2]), cam))
    shift = int(10 * c)
    drop2 = cv2.copyMakeBorder(drop, shift, shift, shift, shift, cv2.BORDER_CONSTANT, value=(0, 0, 0, 0))
    drop2 = gaussian_filter(drop2, [c, c / 2, 0])
    minC_tmp = np.asarray(minC) - shift                                   # :418
    H, W = st.rainy_bg.shape[:2]
    minC2 = np.array([np.clip(minC_tmp[0], 0, W), np.clip(minC_tmp[1], 0, H)])
    delta = minC2 - minC_tmp
    drop2 = drop2[:delta[1]] if delta[1] < 0 else drop2[delta[1]:]
    drop2 = drop2[:, :delta[0]] if delta[0] < 0 else drop2[:, delta[0]:]
    tau_zero = np.sqrt(1.16 * 1e-3) / 50                                  # :425
    length_opacity = cam.opacity_attenuation * d_avg / (s.length + d_avg)  # :426
    tau_one = exposure_time * length_opacity
    y0, x0 = int(minC2[1]), int(minC2[0])
    occ = st.rainy_bg[y0:y0 + drop2.shape[0], x0:x0 + drop2.shape[1], :]
    vis = drop2[:occ.shape[0], :occ.shape[1]]
    alpha = vis[:, :, 3]
    a_ = np.expand_dims(alpha, axis=-1)
    blended = ((1. - ((a_ * tau_one) / exposure_time)) * occ) + vis[:, :, :3] * (tau_one / tau_zero)   # :443-444
    st.rainy_bg[y0:y0 + vis.shape[0], x0:x0 + vis.shape[1]] = np.clip(blended, 0, 1)
    st.rainy_mask[y0:y0 + vis.shape[0], x0:x0 + vis.shape[1]] += alpha   # :450
    st.per_streak.append(dict(pid=s.pid, fov_xy_avg=fov_xy_avg, drop_Y=float(drop_Y), minC=(x0, y0),
                              shape=vis.shape[:2], shift=shift, c=float(c)))
    return vis, (x0, y0)


# --------------------------------------------------------------------------------------
# rows 1,6,14: the per-frame driver            (common/generator.py:299-469)
# --------------------------------------------------------------------------------------

@dataclass
class FrameResult:
    rainy_bg: np.ndarray        # (H,W,3) float64 BGR after the streak loop (before the mean shift)
    out_bgr: np.ndarray         # (H,W,3) float64 BGR mean-shifted (what imsave receives, channel-reversed, clipped)
    out_u8: np.ndarray          # (H,W,3) uint8 BGR  = floor(clip(out,0,1)*255)  (matplotlib float->u8 rule)
    rain_mask: np.ndarray       # (H,W) float64
    fog: np.ndarray
    env: np.ndarray
    n_streaks: int
    skipped: list
    per_streak: list
    tex_idx: list
    noise: list


def render_frame(bg_u8: np.ndarray, depth: np.ndarray, streaks, textures, ratios, cam: Camera, seed: int,
                 tables: EnvTables | None = None, omega: np.ndarray | None = None,
                 f32_mode: str = "native", keep_patches: bool = False, render_scale: int = 1) -> FrameResult:
    """One frame of Generator.run (common/generator.py:318-469) on decoded arrays.
    ``streaks``: the simulator frame's Streak list (XML order, unfiltered).  With ``render_scale`` > 1
    ``bg_u8`` is the full-resolution frame and is reduced like generator.py:354-355."""
    np.random.seed(seed)                                                  # :318
    bg = bg_u8 / 255.0                                                    # :352
    if render_scale != 1:                                                 # :354-355
        bg = cv2.resize(bg, (int(bg.shape[1] // render_scale), int(bg.shape[0] // render_scale)))
    rainy_bg = fog_rain_layer(bg, depth, cam, f32_mode)                   # :386
    fog = rainy_bg.copy()
    tables = tables or build_env_tables(cam.W, cam.H, cam.focal_m)
    env = generate_map(rainy_bg, tables)                                  # :400
    with np.errstate(divide="ignore", invalid="ignore"):
        env_xyY = rgb_to_xyY(env[..., ::-1])                              # :407
    env_xyY[np.isnan(env_xyY)] = 0
    if omega is None:
        omega = solid_angles(env.shape[0], env.shape[1])                  # :410
    st = FrameState(rainy_bg=rainy_bg, rainy_mask=np.zeros(bg.shape[:2], np.float64), env_xyY=env_xyY, omega=omega)
    todo = filter_in_frame(streaks, cam.W, cam.H)                         # :413-420
    tex_idx, noises = [], []
    for s in todo:                                                        # :431
        b = texture_bucket(s.ratio, ratios)
        ti = np.random.randint(10 * b, 10 * b + 10)                       # bad_weather.py:252-264
        tex_idx.append(ti)
        noise = 0.0
        if s.drop_type != BIG:
            noise = np.random.normal(0.0, cam.noise_std) * cam.noise_scale   # generator.py:136
        noises.append(noise)
        patch, minC = make_patch(s, textures[ti], cam, noise)
        poly = fov_polygon(s, cam, env.shape)
        try:
            vis, pos = add_drop_to_image(st, s, patch, minC, poly, cam)
            if keep_patches:
                st.per_streak[-1]["patch"] = vis.copy()
        except Exception as e:                                            # generator.py:185-189
            st.skipped.append((s.pid, repr(e)))
    diff = np.mean(st.rainy_bg) - np.mean(bg)                             # :461-463
    out = st.rainy_bg - diff
    out_u8 = (np.clip(out, 0, 1) * 255).astype(np.uint8)                  # plt.imsave float RGB -> uint8
    return FrameResult(rainy_bg=st.rainy_bg, out_bgr=out, out_u8=out_u8, rain_mask=st.rainy_mask, fog=fog, env=env,
                       n_streaks=len(todo), skipped=st.skipped, per_streak=st.per_streak,
                       tex_idx=tex_idx, noise=noises)


# --------------------------------------------------------------------------------------
# row 14: what plt.imsave(path, rainy_mask) stores (common/generator.py:467)
# --------------------------------------------------------------------------------------

def imsave_mask_index(mask: np.ndarray):
    """plt.imsave of a 2-D float64 array, up to the colour table (matplotlib is absent here and unpinned by the
    reference: PARITY UNPINNED, restated from matplotlib 3.x -- image.imsave -> ScalarMappable.to_rgba(bytes=True) ->
    colors.Normalize.__call__ (autoscaled to the array's min / max: ``(x - vmin) / (vmax - vmin)`` in float64, zeros for
    a flat array) -> Colormap.__call__ (``xa *= N; xa[xa == N] = N - 1; xa.astype(int)`` with N = 256)).
    -> (uint8 index into the 256-entry colormap, (vmin, vmax))."""
    m = np.asarray(mask, np.float64)
    lo, hi = float(m.min()), float(m.max())
    if lo == hi:
        return np.zeros(m.shape, np.uint8), (lo, hi)
    t = (m - lo) / (hi - lo)
    xa = t * 256
    xa[xa == 256] = 255
    return xa.astype(int).astype(np.uint8), (lo, hi)


def mask_u16(mask: np.ndarray) -> np.ndarray:
    """The 16-bit gray form of the mask file this repo offers beside the colormapped one: int(t * 65535 + 0.5) with the
    same normalisation (not a reference format)."""
    m = np.asarray(mask, np.float64)
    lo, hi = float(m.min()), float(m.max())
    if lo == hi:
        return np.zeros(m.shape, np.uint16)
    return (((m - lo) / (hi - lo)) * 65535.0 + 0.5).astype(np.uint16)


# --------------------------------------------------------------------------------------
# file-level helpers (decode side of common/generator.py:352-367, bad_weather.py:108-146)
# --------------------------------------------------------------------------------------

def load_streak_database(tex_dir: str, norm_path: str):
    """DBManager.load_streak_database (:108-146) -> (list of (h,32) uint8 gray textures, sorted unique ratios)."""
    import os
    import re

    def natkey(s):
        return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)]

    norm = {}
    coeff = None
    with open(norm_path) as f:
        for line in f.readlines():
            if line[:2] == "cv":
                coeff = int(line[2:])
                continue
            norm[coeff] = [float(v) for v in line.split("\n")[0].split(" ")[:-1]]
    textures, ratios = [], []
    for name in sorted(os.listdir(tex_dir), key=natkey):
        stem = os.path.splitext(name)[0]
        cv_s, osc_s = stem.split("_")
        k = int(cv_s[-1:]) if len(cv_s) == 3 else int(cv_s[-2:])
        j = int(osc_s[-1:])
        img = cv2.imread(os.path.join(tex_dir, name), cv2.IMREAD_ANYDEPTH)
        tex = ((255.0 * norm[k][j] * img) / 65535.0).astype(np.uint8)
        textures.append(tex)
        ratios.append(tex.shape[1] / tex.shape[0])
    return textures, np.unique(np.array(ratios))


def read_frame(image_path: str, depth_path: str):
    bg = cv2.imread(image_path)
    depth = cv2.imread(depth_path, cv2.IMREAD_UNCHANGED).astype(np.float32) / 256.
    return bg, depth
