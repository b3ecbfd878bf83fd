"""Shared scenario builder for the tests: the same synthetic inputs handed to the oracle
(oracle.rain_oracle) and to the product (rain_rendering_b200)."""
from __future__ import annotations

import os
import tempfile

import numpy as np

from oracle import rain_oracle as ro
from rain_rendering_b200 import api, streaks as S, synth


class Scenario:
    def __init__(self, W, H, n_frames, n_xml, fallrate=25, dataset="kitti", noise_scale=0.0, noise_std=0.0,
                 opacity=1.0, seed=0, n_sim_frames=None, render_scale=1):
        cam = synth.CAMERAS[dataset]
        self.W, self.H, self.n_frames, self.render_scale = W, H, n_frames, render_scale
        self.cam = ro.Camera(W=W, H=H, focal_mm=cam["cam_focal"], f_number=cam["cam_f_number"],
                             exposure_ms=cam["cam_exposure"], gain=cam["cam_gain"], fallrate=fallrate,
                             opacity_attenuation=opacity, noise_scale=noise_scale, noise_std=noise_std)
        self.db = synth.make_streak_db(seed)
        rs = render_scale
        self.bgr = np.stack([synth.make_frame(W * rs, H * rs, seed * 1000 + i)[0] for i in range(n_frames)])
        self.depth = np.stack([synth.make_frame(W, H, seed * 1000 + i)[1] for i in range(n_frames)])
        nsf = n_sim_frames or n_frames
        parts = synth.make_particles(W, H, nsf, n_xml, cam["cam_exposure"], seed, rs)
        with tempfile.TemporaryDirectory() as d:
            xml = os.path.join(d, "sim_camera0.xml")
            synth.write_particles_xml(parts, xml, cam["cam_exposure"])
            self.oracle_frames = ro.load_streaks_from_xml(xml, rs, W, H)
            self.sim_frames = S.load_streaks_from_xml(xml, rs, W, H)
        self._tables = None
        self._omega = None

    # ---- oracle side ----
    @property
    def tables(self):
        if self._tables is None:
            self._tables = ro.build_env_tables(self.W, self.H, self.cam.focal_m)
        return self._tables

    @property
    def omega(self):
        if self._omega is None:
            self._omega = ro.solid_angles(self.H, self.tables.W_env)
        return self._omega

    def oracle_frame(self, i, f32_mode="canonical", keep_patches=False):
        return ro.render_frame(self.bgr[i], self.depth[i], self.oracle_frames[i % len(self.oracle_frames)],
                               self.db.textures, self.db.ratios, self.cam, i, self.tables, self.omega,
                               f32_mode=f32_mode, keep_patches=keep_patches, render_scale=getattr(self, "render_scale", 1))

    # ---- product side ----
    def records(self):
        """(concatenated STREAK_DTYPE records, offsets) for all frames, in frame order (stateful
        like the reference: wind-noise write-back persists in the simulator frames)."""
        recs, offs = [], [0]
        for i in range(self.n_frames):
            sim = self.sim_frames[i % len(self.sim_frames)]
            r = api.assemble_frame_records(sim, self.W, self.H, self.db.ratios, i, self.cam.noise_std, self.cam.noise_scale)
            recs.append(r)
            offs.append(offs[-1] + len(r))
        return np.concatenate(recs) if recs else np.zeros(0, S.STREAK_DTYPE), np.array(offs, np.int32)

    def context(self, max_batch=None, device=0):
        ctx = api.RainContext(device)
        ctx.set_streak_db(self.db.textures, self.db.ratios)
        c = self.cam
        ctx.set_camera(self.W, self.H, c.focal_mm, c.f_number, c.exposure_ms, c.gain, c.fallrate, c.opacity_attenuation,
                       max_batch or self.n_frames, render_scale=getattr(self, "render_scale", 1))
        return ctx


def ulp_diff_f32(a, b):
    """|a - b| in units of float32 ULPs (a, b float32 arrays, finite)."""
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_scenario(name: str):
    """Scenario rebuilt from a committed fixture (inputs + outputs of the live reference,
    see oracle/make_golden.py).  Returns (scenario, npz)."""
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    W, H, nf = int(g["W"]), int(g["H"]), int(g["n_frames"])
    sc = Scenario.__new__(Scenario)
    sc.W, sc.H, sc.n_frames, sc.render_scale = W, H, nf, 1
    cam = synth.CAMERAS["customdb"]
    sc.cam = ro.Camera(W=W, H=H, focal_mm=cam["cam_focal"], f_number=cam["cam_f_number"], exposure_ms=cam["cam_exposure"],
                       gain=cam["cam_gain"], fallrate=int(g["fallrate"]), opacity_attenuation=float(g["opacity"]),
                       noise_scale=float(g["noise_scale"]), noise_std=float(g["noise_std"]))
    sc.db = synth.make_streak_db(int(g["db_seed"]))
    sc.bgr = np.ascontiguousarray(g["bgr"])
    sc.depth = np.ascontiguousarray(g["depth_u16"].astype(np.float32) / 256.)
    with tempfile.TemporaryDirectory() as d:
        xml = os.path.join(d, "sim_camera0.xml")
        with open(xml, "wb") as f:
            f.write(g["xml"].tobytes())
        sc.oracle_frames = ro.load_streaks_from_xml(xml, 1, W, H)
        sc.sim_frames = S.load_streaks_from_xml(xml, 1, W, H)
    sc._tables = None
    sc._omega = None
    return sc, g
