"""Seeded synthetic inputs for the rain-rendering hot path (no datasets ship with this repo).

Everything here is *input generation*: RGB + depth frames, a Garg-Nayar-layout streak
database and a particles XML in the schema the reference parses
(reference: common/bad_weather.py:108-146 for the DB layout, :185-239 for the XML
attributes).  The shapes/statistics follow SURVEY.md section 8(d).  numpy only; the
on-disk writers lazily import cv2 for PNG encoding (I/O, not the hot path).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np

# (focal mm, f-number, gain, exposure ms, render_scale) -- reference config/*.py settings()
CAMERAS = {
    "customdb": dict(cam_focal=6.0, cam_f_number=6.0, cam_gain=20.0, cam_exposure=2.0, render_scale=1),
    "kitti": dict(cam_focal=6.0, cam_f_number=6.0, cam_gain=20.0, cam_exposure=2.0, render_scale=1),
    "cityscapes": dict(cam_focal=6.0, cam_f_number=6.0, cam_gain=20.0, cam_exposure=5.0, render_scale=2),
    "nuscenes": dict(cam_focal=5.5, cam_f_number=1.8, cam_gain=1.0, cam_exposure=5.0, render_scale=1),
}

# named workloads of BASELINE.json (W, H, fallrate mm/h, dataset, streaks-in-XML per frame)
WORKLOADS = {
    "C1": dict(W=640, H=480, fallrate=10, dataset="customdb", n_xml=420),
    "C2": dict(W=1242, H=375, fallrate=25, dataset="kitti", n_xml=650),
    "C3": dict(W=1024, H=512, fallrate=50, dataset="cityscapes", n_xml=1300),
    "C4": dict(W=1600, H=900, fallrate=25, dataset="nuscenes", n_xml=650),
    "C5": dict(W=1242, H=375, fallrate=100, dataset="kitti", n_xml=2600),
}


def _bilinear_upsample(small: np.ndarray, H: int, W: int) -> np.ndarray:
    """Plain bilinear upsampling of a (h, w, c) array to (H, W, c) (numpy only)."""
    h, w = small.shape[:2]
    ys = (np.arange(H) + 0.5) * h / H - 0.5
    xs = (np.arange(W) + 0.5) * w / W - 0.5
    y0 = np.clip(np.floor(ys).astype(int), 0, h - 1)
    x0 = np.clip(np.floor(xs).astype(int), 0, w - 1)
    y1 = np.clip(y0 + 1, 0, h - 1)
    x1 = np.clip(x0 + 1, 0, w - 1)
    fy = np.clip(ys - y0, 0, 1)[:, None, None]
    fx = np.clip(xs - x0, 0, 1)[None, :, None]
    a = small[y0][:, x0]
    b = small[y0][:, x1]
    c = small[y1][:, x0]
    d = small[y1][:, x1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def make_frame(W: int, H: int, seed: int):
    """One synthetic frame: uint8 BGR (H, W, 3) and float32 depth in metres (H, W).

    Low-frequency colour noise plus a vertical sky gradient; depth falls off with image
    height like a road scene.  Depth is quantised to 1/256 m so the uint16 PNG round trip
    of the reference (common/generator.py:360-365: ``imread(...)/256.``) is lossless.
    """
    rng = np.random.RandomState((1234 + 7919 * int(seed)) % (2 ** 32))
    small = rng.uniform(0, 255, size=(max(H // 8, 2), max(W // 8, 2), 3))
    img = _bilinear_upsample(small, H, W)
    sky = np.linspace(70.0, -25.0, H)[:, None, None]
    img = np.clip(0.8 * img + sky + rng.uniform(-6, 6, size=(H, W, 3)), 0, 255)
    bgr = np.ascontiguousarray(img.astype(np.uint8))
    y = np.arange(H, dtype=np.float64)[:, None]
    depth = np.clip(80.0 * (1.0 - y / H) + 2.0 + rng.uniform(0, 2, size=(H, W)), 1.0, 200.0)
    depth_u16 = np.round(depth * 256.0).astype(np.uint16)
    return bgr, np.ascontiguousarray(depth_u16.astype(np.float32) / np.float32(256.0))


@dataclass
class StreakDB:
    """Normalised uint8 textures exactly as the reference keeps them after
    DBManager.load_streak_database (common/bad_weather.py:139-146)."""
    textures: list            # list of (h_t, 32) uint8 gray (the reference's BGR has 3 equal channels)
    ratios: np.ndarray        # sorted unique W/H ratios (common/bad_weather.py:143-145)
    raw16: list               # the 16-bit source images (for writing the on-disk DB)
    norm: dict                # cv -> list of 10 coefficients

    def concat(self):
        heights = np.array([t.shape[0] for t in self.textures], dtype=np.int32)
        data = np.concatenate([t.reshape(-1) for t in self.textures]).astype(np.uint8)
        return heights, data


DB_HEIGHTS = (640, 320, 160, 80, 40)


def make_streak_db(seed: int = 0, heights=DB_HEIGHTS, width: int = 32) -> StreakDB:
    """5 ``cv`` x 10 ``osc`` textures, width 32, Gaussian-profile streaks with an
    osc-dependent modulation, 16-bit, plus normalisation coefficients in [0.5, 0.95]."""
    rng = np.random.RandomState((4321 + int(seed)) % (2 ** 32))
    raw16, textures, norm = [], [], {}
    for k, h in enumerate(heights):
        coeffs = [float(np.round(rng.uniform(0.5, 0.95), 6)) for _ in range(10)]
        norm[k] = coeffs
        for j in range(10):
            y = (np.arange(h) + 0.5) / h
            x = (np.arange(width) + 0.5) / width - 0.5
            centre = 0.06 * np.sin(2 * np.pi * (j + 1) * y * 0.5 + k)[:, None]
            sig = 0.09 + 0.03 * np.cos(2 * np.pi * (j + 1) * y)[:, None] * 0.5
            prof = np.exp(-0.5 * ((x[None, :] - centre) / sig) ** 2)
            env = (np.sin(np.pi * y) ** 0.35)[:, None]
            osc = 0.75 + 0.25 * np.cos(2 * np.pi * (j + 2) * y + 0.3 * j)[:, None]
            img = prof * env * osc
            img[img < 0.02] = 0.0
            img16 = np.round(np.clip(img, 0, 1) * 65535.0).astype(np.uint16)
            raw16.append(img16)
            # reference: ((255.0 * norm * img) / 65535.0).astype(uint8)  (bad_weather.py:141)
            textures.append(((255.0 * coeffs[j] * img16) / 65535.0).astype(np.uint8))
    ratios = np.unique(np.array([width / h for h in heights]))
    return StreakDB(textures=textures, ratios=ratios, raw16=raw16, norm=norm)


def write_streak_db(db: StreakDB, root: str):
    """Write the DB in the reference's on-disk layout (main.py:132-137)."""
    import cv2
    tex_dir = os.path.join(root, "env_light_database", "size32")
    txt_dir = os.path.join(root, "env_light_database", "txt")
    os.makedirs(tex_dir, exist_ok=True)
    os.makedirs(txt_dir, exist_ok=True)
    n_cv = len(db.raw16) // 10
    for k in range(n_cv):
        for j in range(10):
            cv2.imwrite(os.path.join(tex_dir, "cv%d_osc%d.png" % (k, j)), db.raw16[k * 10 + j])
    with open(os.path.join(txt_dir, "normalized_env_max.txt"), "w") as f:
        for k in range(n_cv):
            f.write("cv%d\n" % k)
            # NB trailing space before the newline: the reference parser drops the last token
            f.write(" ".join(repr(c) for c in db.norm[k]) + " \n")


def terminal_velocity(d_m):
    """Atlas et al. 1973."""
    return 9.65 - 10.3 * np.exp(-600.0 * d_m)


def make_particles(W: int, H: int, n_frames: int, n_xml: int, exposure_ms: float, seed: int = 0,
                   render_scale: int = 1):
    """Synthetic imaged streaks per simulator frame, in *simulator* conventions
    (image y up, world z negative forward, positions at full sensor resolution).

    Returns a list (per frame) of dicts of arrays: pid, wp1, wp2 (n,3), wd (n,), ip1, ip2 (n,2),
    iw1, iw2 (n,).  Pinhole with f_px = 6e-3/4.65e-6 * (W*render_scale/1242).
    """
    rng = np.random.RandomState((99 + int(seed)) % (2 ** 32))
    Wf, Hf = W * render_scale, H * render_scale
    f_px = 6e-3 / 4.65e-6 * (Wf / 1242.0)
    T = exposure_ms / 1000.0
    frames = []
    for _ in range(n_frames):
        z = rng.uniform(0.25, 3.5, n_xml)
        d = rng.uniform(0.5e-3, 4e-3, n_xml)
        # sample image positions slightly beyond the sensor so some streaks are culled
        u = rng.uniform(-0.05 * Wf, 1.05 * Wf, n_xml)
        v = rng.uniform(-0.05 * Hf, 1.05 * Hf, n_xml)   # y up
        x1 = (u - Wf / 2.0) * z / f_px
        y1 = (v - Hf / 2.0) * z / f_px
        fall = terminal_velocity(d) * T
        wind = rng.normal(0.0, 0.15, n_xml) * fall
        x2 = x1 + wind
        y2 = y1 - fall
        z2 = z + rng.normal(0, 0.002, n_xml)
        ip1 = np.stack([u, v], 1)
        ip2 = np.stack([Wf / 2.0 + f_px * x2 / z2, Hf / 2.0 + f_px * y2 / z2], 1)
        iw1 = d * f_px / z
        iw2 = d * f_px / z2
        frames.append(dict(
            pid=np.arange(n_xml, dtype=np.int64),
            wp1=np.stack([x1, y1, -z], 1), wp2=np.stack([x2, y2, -z2], 1),
            wd=d, ip1=ip1, ip2=ip2, iw1=iw1, iw2=iw2))
    return frames


def write_particles_xml(frames, path: str, exposure_ms: float):
    """Serialise in the AHLSimulation schema."""
    os.makedirs(os.path.dirname(path), exist_ok=True)

    def vec(a):
        return "[" + ";".join(repr(float(t)) for t in a) + "]"

    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<camera statslevel="0">\n')
        for fi, fr in enumerate(frames):
            n = len(fr["pid"])
            f.write('  <i id="%d" t="%d" d="%d" rs="%d">\n' % (fi, int(exposure_ms * 1e6), int(fi * 1e8), n))
            for k in range(n):
                f.write('    <r pid="%d" wp1="%s" wd1="%r" wp2="%s" wd2="%r" ip1="%s" iw1="%r" ip2="%s" iw2="%r" />\n' % (
                    int(fr["pid"][k]), vec(fr["wp1"][k]), float(fr["wd"][k]), vec(fr["wp2"][k]), float(fr["wd"][k]),
                    vec(fr["ip1"][k]), float(fr["iw1"][k]), vec(fr["ip2"][k]), float(fr["iw2"][k])))
            f.write("  </i>\n")
        f.write("</camera>\n")


def write_dataset(root: str, dataset: str, seq: str, W: int, H: int, n_frames: int, fallrate: int,
                  n_xml: int, seed: int = 0, n_sim_frames: int | None = None, render_scale: int | None = None):
    """Lay out a complete reference-style tree under ``root`` (customdb layout,
    config/customdb.py:5-20): images, uint16 depth PNGs, streak DB and particles XML.
    Returns the dict of paths main.check_arg needs."""
    import cv2
    cam = CAMERAS[dataset]
    rs = render_scale or cam["render_scale"]
    src = os.path.join(root, "source", dataset, seq)
    os.makedirs(os.path.join(src, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(src, "depth"), exist_ok=True)
    for i in range(n_frames):
        # the image at sensor resolution (W*rs, H*rs); the depth map at render resolution (depth_scale = rs,
        # the Cityscapes arrangement, config/cityscapes.py:41-42)
        bgr, _ = make_frame(W * rs, H * rs, seed * 1000 + i)
        _, depth = make_frame(W, H, seed * 1000 + i)
        cv2.imwrite(os.path.join(src, "rgb", "%06d.png" % i), bgr)
        cv2.imwrite(os.path.join(src, "depth", "%06d.png" % i), np.round(depth * 256.0).astype(np.uint16))
    db = make_streak_db(seed)
    db_root = os.path.join(root, "rainstreakdb")
    write_streak_db(db, db_root)
    nsf = n_sim_frames or n_frames
    frames = make_particles(W, H, nsf, n_xml, cam["cam_exposure"], seed, rs)
    xml = os.path.join(root, "particles", dataset, seq, "rain", "%dmm" % fallrate, "sim_camera0.xml")
    write_particles_xml(frames, xml, cam["cam_exposure"])
    return dict(dataset_root=os.path.join(root, "source"), particles=os.path.join(root, "particles"),
                streaks_db=db_root, xml=xml, output=os.path.join(root, "output"))
