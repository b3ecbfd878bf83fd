"""The exact arithmetic the CUDA kernels execute (rain_rendering_b200/csrc/rr_cvmath.h and
rr_streak_geom.h are host+device headers), compiled for the host by tests/hostsim and checked
bit-for-bit against cv2 and the oracle on the CPU."""
import ctypes as C
import os
import subprocess

import cv2
import numpy as np
import pytest

from oracle import rain_oracle as ro
from rain_rendering_b200 import _lib, streaks as S
from util import Scenario

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hs():
    src = os.path.join(HERE, "hostsim", "hostsim.cpp")
    so = os.path.join(HERE, "hostsim", "libhostsim.so")
    deps = [src] + [os.path.join(HERE, "..", "rain_rendering_b200", "csrc", h) for h in ("rr_cvmath.h", "rr_streak_geom.h", "rr_types.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src])
    lib = C.CDLL(so)
    assert lib.hs_sizeof_plan() == _lib.PLAN_DTYPE.itemsize and lib.hs_sizeof_rec() == S.STREAK_DTYPE.itemsize
    return lib


class CamDev(C.Structure):
    _fields_ = [("W", C.c_int), ("H", C.c_int), ("H_env", C.c_int), ("W_env", C.c_int)] + \
               [(n, C.c_double) for n in ("focal_m", "f_number", "focus_plane", "pix_size", "radius", "fov_deg", "opacity_att", "exposure_blend")] + \
               [("db_width", C.c_int), ("n_tex", C.c_int)]


def test_fill_convex_poly_restatement_equals_cv2(hs):
    rng = np.random.RandomState(0)
    for trial in range(1200):
        W, H = rng.randint(20, 200), rng.randint(20, 120)
        n = rng.randint(3, 24)
        kind = trial % 4
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        if kind == 0:
            pts = np.stack([rng.uniform(.2, .8) * W + rng.uniform(.1, .6) * W * np.cos(ang), rng.uniform(.2, .8) * H + rng.uniform(.1, .6) * H * np.sin(ang)], 1)
            pts = np.clip(pts[::-1] if rng.rand() < .5 else pts, 0, [W, H])
        elif kind == 1:
            r = rng.uniform(.1, .5, n)
            pts = np.clip(np.stack([W / 2 + r * W * np.cos(ang), H / 2 + r * H * np.sin(ang)], 1), 0, [W, H])
        elif kind == 2:
            pts = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
        else:
            pts = np.clip(np.stack([W / 2 + .7 * W * np.cos(ang), H / 2 + .7 * H * np.sin(ang)], 1), -5, [W + 5, H + 5])
        pts = pts.astype(np.int64)
        pts = np.vstack([pts, pts[:1]])
        ref = np.zeros((H, W), np.float64)
        cv2.fillConvexPoly(ref, pts, 1)
        vx, vy = np.ascontiguousarray(pts[:, 0], np.int32), np.ascontiguousarray(pts[:, 1], np.int32)
        mine = np.zeros((H, W), np.uint8)
        hs.hs_fill_convex_poly(_lib.ptr(vx), _lib.ptr(vy), len(pts), W, H, _lib.ptr(mine))
        assert np.array_equal(mine.astype(bool), ref.astype(bool)), (trial, pts.tolist())


@pytest.mark.parametrize("W,H,n_xml,exposure_ds,noise", [(1242, 375, 700, "kitti", 3.0), (640, 480, 2500, "kitti", 0.0), (1600, 900, 400, "nuscenes", 0.0),
                                                         (1024, 512, 500, "cityscapes", 8.0)])
def test_patch_and_fov_mask_equal_oracle(hs, W, H, n_xml, exposure_ds, noise):
    sc = Scenario(W, H, 1, n_xml, dataset=exposure_ds, noise_scale=1.0 if noise else 0.0, noise_std=noise, seed=3)
    cam = sc.cam
    cd = CamDev(W, H, H, sc.tables.W_env, cam.focal_m, cam.f_number, cam.focus_plane, cam.pix_size, cam.radius, cam.fov_deg,
                cam.opacity_attenuation, cam.exposure_ms / 1000., 32, 50)
    recs, offs = sc.records()
    todo = ro.filter_in_frame(sc.oracle_frames[0], W, H)
    assert len(todo) == len(recs) > 100
    plan = np.zeros(1, _lib.PLAN_DTYPE)
    modes, nverts = set(), set()
    for s, r in zip(todo, recs):
        rr = np.array([r])
        tex = np.ascontiguousarray(sc.db.textures[r["tex_idx"]])
        patch, minC = ro.make_patch(s, tex, cam, float(r["noise_deg"]))
        out = np.zeros(patch.size + 8)
        n = hs.hs_patch(_lib.ptr(rr), C.byref(cd), _lib.ptr(tex), tex.shape[0], _lib.ptr(plan), _lib.ptr(out), out.size)
        p = plan[0]
        assert (p["ph"], p["pw"]) == patch.shape and (p["minx"], p["miny"]) == tuple(minC)
        assert np.array_equal(out[:n].reshape(patch.shape), patch), (s.pid, s.drop_type)
        modes.add("big" if p["type"] == 0 else int(p["resize_mode"]))
        # field of view
        poly = ro.fov_polygon(s, cam, (H, sc.tables.W_env, 3))
        mref, sref = ro.fov_mask(poly, H, sc.tables.W_env)
        mask = np.zeros((H, sc.tables.W_env), np.uint8)
        pxy, npoly = np.zeros(64), C.c_int(0)
        ivx, ivy = np.zeros(40, np.int32), np.zeros(40, np.int32)
        m = hs.hs_fov_mask(_lib.ptr(rr), C.c_double(cam.radius), C.c_double(cam.fov_deg), H, sc.tables.W_env, _lib.ptr(mask),
                           _lib.ptr(pxy), C.byref(npoly), _lib.ptr(ivx), _lib.ptr(ivy))
        assert npoly.value == len(poly) and np.abs(pxy[:2 * len(poly)].reshape(-1, 2) - poly).max() < 1e-9
        assert m == len(sref) and np.array_equal(mask.astype(bool), mref)
        nverts.add(npoly.value)
    assert "big" in modes and 2 in modes


def test_scipy_gaussian_weights(hs):
    from scipy.ndimage import gaussian_filter1d
    for sigma in [0.13, 0.4, 1.08, 2.37, 4.95, 9.0]:
        r = C.c_int(0)
        w = np.zeros(200)
        hs.hs_gauss_weights(C.c_double(sigma), C.byref(r), _lib.ptr(w))
        assert r.value == int(4 * sigma + 0.5)
        imp = np.zeros(2 * r.value + 1)
        imp[r.value] = 1.0
        ref = gaussian_filter1d(imp, sigma, mode="constant")
        assert np.abs(w[:2 * r.value + 1] - ref).max() < 1e-15


def test_division_free_xyY_is_bit_exact_for_every_colour(hs):
    """k_env_prefix converts BGR -> xyY with reciprocal + FMA corrections instead of five IEEE divisions;
    the two forms must agree bit for bit on the whole input domain (all 2^24 uint8 colours, NaN -> 0 included)."""
    hs.hs_env_xyY_mismatches.restype = C.c_long
    assert hs.hs_env_xyY_mismatches() == 0
    hs.hs_u8_unit_mismatches.restype = C.c_long
    assert hs.hs_u8_unit_mismatches() == 0
    # and the literal form is the reference's arithmetic (common/my_utils.py:55-68)
    rng = np.random.RandomState(5)
    bgr = rng.randint(0, 256, (500, 3)).astype(np.uint8)
    bgr[:3] = [[0, 0, 0], [255, 255, 255], [0, 0, 1]]
    out = np.zeros(3)
    for b, g, r in bgr:
        hs.hs_env_xyY(C.c_double(b / 255.0), C.c_double(g / 255.0), C.c_double(r / 255.0), _lib.ptr(out))
        with np.errstate(all="ignore"):
            ref = np.nan_to_num(ro.rgb_to_xyY(np.array([[[r / 255.0, g / 255.0, b / 255.0]]]))[0, 0])
        assert np.allclose(out, ref, rtol=1e-14, atol=0), (b, g, r)      # np.dot's BLAS order / FMA differs in the last bit


@pytest.mark.parametrize("W,H,n_xml,noise", [(1242, 375, 700, 3.0), (640, 480, 2500, 0.0), (1242, 375, 900, 25.0)])
def test_canvas_row_spans_cover_every_sample_that_touches_the_texture(hs, W, H, n_xml, noise):
    """k_raster samples only the column span rr_canvas_row_span returns for a canvas row and writes zeros elsewhere:
    no pixel outside a span may have a bilinear tap inside the texture."""
    sc = Scenario(W, H, 1, n_xml, noise_scale=1.0 if noise else 0.0, noise_std=noise, seed=11)
    cam = sc.cam
    cd = CamDev(W, H, H, sc.tables.W_env, cam.focal_m, cam.f_number, cam.focus_plane, cam.pix_size, cam.radius, cam.fov_deg,
                cam.opacity_attenuation, cam.exposure_ms / 1000., 32, 50)
    recs, _ = sc.records()
    hs.hs_canvas_span_violations.restype = C.c_long
    plan = np.zeros(1, _lib.PLAN_DTYPE)
    stats = np.zeros(3, np.int64)
    n_checked = 0
    for r in recs:
        rr = np.array([r])
        tex = np.ascontiguousarray(sc.db.textures[r["tex_idx"]])
        hs.hs_patch(_lib.ptr(rr), C.byref(cd), _lib.ptr(tex), tex.shape[0], _lib.ptr(plan), None, 0)
        if not plan[0]["valid"] or plan[0]["type"] == 0:
            continue
        assert hs.hs_canvas_span_violations(_lib.ptr(plan), 32, _lib.ptr(stats)) == 0, int(r["pid"])
        n_checked += 1
    assert n_checked > 100
    assert stats[2] <= stats[1] <= stats[0] and stats[1] < 1.15 * stats[2] + 4 * n_checked * 400     # spans are tight, not just safe
