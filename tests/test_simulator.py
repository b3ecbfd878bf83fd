"""On-the-fly particle simulator (rr_simulate_particles): the closed AHLSimulation binary cannot
run anywhere, so parity is statistical: drop-size statistics, the v_t(D) curve,
streak counts proportional to the fall rate, imaging geometry, field-of-view culling."""
import ctypes as C

import numpy as np
import pytest

from rain_rendering_b200 import _lib, api, streaks as S, synth


def _phys(D):
    lib = _lib.load()
    v, f, m = C.c_double(), C.c_double(), C.c_double()
    lib.rr_host_sim_physics(C.c_double(D), C.byref(v), C.byref(f), C.byref(m))
    return v.value, f.value, m.value


def test_force_model_terminal_velocity_follows_atlas_1973():
    # AHLRaindrop::TerminalVelocityAtlasEtAl1973: v_t = 9.65 - 10.3 exp(-600 D); the drag law's own
    # equilibrium must agree with it over the sizes that matter (0.5 .. 5 mm)
    for D in (0.5e-3, 1e-3, 2e-3, 3e-3, 4e-3, 5e-3):
        v, f, m = _phys(D)
        atlas = 9.65 - 10.3 * np.exp(-600 * D)
        assert abs(f / (m * 9.81) - 1) < 1e-9          # equilibrium of the integrator's force balance
        assert abs(m - 1000 * np.pi / 6 * D ** 3) < 1e-15
        assert abs(v / atlas - 1) < 0.12, (D, v, atlas)
    vs = [_phys(D)[0] for D in np.linspace(0.2e-3, 6e-3, 30)]
    assert all(b > a for a, b in zip(vs, vs[1:]))      # monotone in diameter


@pytest.mark.gpu
def test_simulated_streaks_statistics_and_geometry():
    ctx = api.RainContext(0)
    W, H, T = 1242, 375, 2.0
    f_px = 6e-3 / 4.65e-6
    frames, mean25 = ctx.simulate_particles(0, 60, W, H, 25, exposure_ms=T, seed=11)
    frames50, mean50 = ctx.simulate_particles(0, 60, W, H, 50, exposure_ms=T, seed=11)
    n25 = np.mean([len(f) for f in frames])
    n50 = np.mean([len(f) for f in frames50])
    assert 100 < n25 < 20000
    # counts grow with the fall rate, slightly slower than linearly (Marshall-Palmer slope changes)
    assert 1.3 < n50 / n25 < 2.2
    allr = np.concatenate(frames)
    # imaging geometry: widths, depth, streak length
    z1, z2 = -allr["wp1"][:, 2], -allr["wp2"][:, 2]
    assert (z1 > 0.2).all() and (z1 < 15.01).all()
    assert np.allclose(allr["iw1"], allr["wd1"] * f_px / z1, rtol=1e-12)
    fall = allr["wp1"][:, 1] - allr["wp2"][:, 1]
    vt = np.array([_phys(D)[0] for D in allr["wd1"][:200]])
    assert np.allclose(fall[:200], vt * T / 1000., rtol=2e-3)      # already at terminal velocity: constant-speed branch
    # FOV culling: an end point inside the sensor, and wide enough somewhere
    u1, v1, u2, v2 = allr["ip1"][:, 0], allr["ip1"][:, 1], allr["ip2"][:, 0], allr["ip2"][:, 1]
    in1 = (u1 >= 0) & (u1 < W) & (v1 >= 0) & (v1 < H)
    in2 = (u2 >= 0) & (u2 < W) & (v2 >= 0) & (v2 < H)
    assert (in1 | in2).all()
    assert (np.maximum(allr["iw1"], allr["iw2"]) >= 1.0).all()
    assert np.allclose(u1, W / 2 + f_px * allr["wp1"][:, 0] / z1) and np.allclose(v1, H / 2 + f_px * allr["wp1"][:, 1] / z1)
    # uniform over the sensor horizontally
    hist = np.histogram(u1[in1], bins=6, range=(0, W))[0]
    assert hist.min() > 0.7 * hist.mean()
    # drop sizes: visible drops are biased to large diameters, none outside the limits
    d_mm = allr["wd1"] * 1e3
    assert d_mm.min() >= 0.1 and d_mm.max() <= 10 and 0.6 < np.median(d_mm) < 3.0
    # determinism: a frame depends on (seed, absolute index) only
    again, _ = ctx.simulate_particles(7, 2, W, H, 25, exposure_ms=T, seed=11)
    assert np.array_equal(again[0], frames[7]) and np.array_equal(again[1], frames[8])
    other, _ = ctx.simulate_particles(7, 1, W, H, 25, exposure_ms=T, seed=12)
    assert len(other[0]) != len(frames[7]) or not np.array_equal(other[0], frames[7])
    ctx.close()


@pytest.mark.gpu
def test_simulator_feeds_the_renderer():
    """Config C3 shape: simulate on the fly, build records with the XML loader's arithmetic, render."""
    W, H = 1024, 512
    ctx = api.RainContext(0)
    db = synth.make_streak_db(0)
    ctx.set_streak_db(db.textures, db.ratios)
    ctx.set_camera(W, H, exposure_ms=5.0, fallrate=50, max_batch=2)
    sims, _ = ctx.simulate_particles(0, 2, W, H, 50, pix_size_um=4.65 * 1242 / W, exposure_ms=5.0, seed=3)
    recs, offs = [], [0]
    for i, sim in enumerate(sims):
        r = S.records_from_sim(sim, 1, W, H)
        r = api.assemble_frame_records(r, W, H, db.ratios, i)
        recs.append(r)
        offs.append(offs[-1] + len(r))
    assert offs[-1] > 200
    frames = [synth.make_frame(W, H, i) for i in range(2)]
    out = ctx.render_frames(np.stack([f[0] for f in frames]), np.stack([f[1] for f in frames]), np.concatenate(recs), np.array(offs, np.int32))
    assert np.isfinite(out["bgr"]).all() and (out["mask"] > 0).mean() > 0.02
    ctx.close()
