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


def _device_records(ptr, n):
    """copy n rr_streak_rec back from a raw device pointer (tests only)"""
    import torch
    from rain_rendering_b200.dist import _DevicePtr
    if n == 0:
        return np.zeros(0, S.STREAK_DTYPE)
    t = torch.as_tensor(_DevicePtr(ptr, n * S.STREAK_DTYPE.itemsize), device=torch.device("cuda", 0))
    return t.cpu().numpy().view(S.STREAK_DTYPE).copy()


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,rs,fallrate,exposure", [(1242, 375, 1, 25, 2.0), (2048, 1024, 2, 50, 5.0)])
def test_device_resident_simulation_gives_the_host_paths_records_bit_for_bit(W, H, rs, fallrate, exposure):
    """rr_simulate_records_device (BASELINE C3: on-the-fly simulation feeding the renderer with no host round trip of the
    streaks) against the host route it replaces: rr_simulate_particles -> the XML loader's arithmetic (records_from_sim,
    common/bad_weather.py:200-238) -> in-frame filter + NumPy-RNG mirror (assemble_batch).  Same float64 operations in the
    same order and the same MT19937 stream: the records must be identical, byte for byte."""
    ctx = api.RainContext(0)
    db = synth.make_streak_db(0)
    n, first = 6, 3
    pix = 4.65 * 1242 / W
    sims, mean = ctx.simulate_particles(first, n, W, H, fallrate, pix_size_um=pix, exposure_ms=exposure, seed=5)
    host = [S.records_from_sim(sim, rs, W // rs, H // rs) for sim in sims]
    want, want_offs = api.assemble_batch(host, list(range(first, first + n)), W // rs, H // rs, db.ratios)
    ptr, offs, mean_d = ctx.simulate_records_device(first, n, W, H, fallrate, db.ratios, render_scale=rs, pix_size_um=pix, exposure_ms=exposure, seed=5)
    assert mean_d == mean and np.array_equal(offs, want_offs) and offs[-1] > 50 * n
    got = _device_records(ptr, int(offs[-1]))
    for name in S.STREAK_DTYPE.names:
        assert np.array_equal(got[name], want[name], equal_nan=True) if got[name].dtype.kind == "f" else np.array_equal(got[name], want[name]), name
    assert set(np.unique(got["type"]).tolist()) <= {0, 1, 2} and len(np.unique(got["tex_idx"])) > 5
    # and straight into the renderer, device to device
    import torch
    Wr, Hr = W // rs, H // rs
    ctx.set_streak_db(db.textures, db.ratios)
    ctx.set_camera(Wr, Hr, exposure_ms=exposure, fallrate=fallrate, max_batch=n, render_scale=rs)
    frames = [synth.make_frame(W, H, i) for i in range(n)]
    bgr = np.stack([f[0] for f in frames])
    d16 = np.stack([np.rint(synth.make_frame(Wr, Hr, i)[1] * 256).astype(np.uint16) for i in range(n)])
    dev = torch.device("cuda", 0)
    t_bgr, t_d = torch.from_numpy(bgr).to(dev), torch.from_numpy(d16.view(np.int16)).to(dev)
    t_u8 = torch.empty((n, Hr, Wr, 3), dtype=torch.uint8, device=dev)
    t_mask = torch.empty((n, Hr, Wr), dtype=torch.float32, device=dev)
    ptr, offs, _ = ctx.simulate_records_device(first, n, W, H, fallrate, db.ratios, render_scale=rs, pix_size_um=pix, exposure_ms=exposure, seed=5)
    ctx.render_frames_device(t_bgr.data_ptr(), t_d.data_ptr(), True, ptr, offs, d_out_mask=t_mask.data_ptr(), d_out_u8=t_u8.data_ptr())
    ref = ctx.render_frames(bgr, d16, want, want_offs, want=("mask", "u8"))
    assert np.array_equal(t_u8.cpu().numpy(), ref["u8"]) and np.array_equal(t_mask.cpu().numpy(), ref["mask"])
    ctx.close()


@pytest.mark.gpu
def test_drop_sizes_follow_marshall_palmer_and_counts_follow_the_rate():
    """statistical parity of the simulator stand-in.  (1) The diameters of the imaged drops must follow the
    density the model states -- Marshall-Palmer N(D) = 8000 exp(-4.1 R^-0.21 D) weighted by 1 / v(D) (airborne
    concentration of a flux) and by the volume in which a drop of that size is imaged wider than a pixel: chi-square
    against that density, computed here independently from the force model.  (2) The mean streak count at five rates
    follows the water budget: proportional to R * integral(N V)^-1 ..., i.e. to the expectation the library reports, and
    grows monotonically with R."""
    ctx = api.RainContext(0)
    W, H, T = 1242, 375, 2.0
    f_px = 6e-3 / 4.65e-6
    R = 25.0
    frames, mean = ctx.simulate_particles(0, 150, W, H, R, exposure_ms=T, seed=21)
    allr = np.concatenate(frames)
    d_mm = allr["wd1"] * 1e3
    assert len(d_mm) > 20000
    # expected density of imaged diameters
    edges = np.linspace(0.1, 6.0, 25)
    centres = np.linspace(0.1, 10.0, 4001)
    lam = 4.1 * R ** -0.21
    v = np.array([_phys(D * 1e-3)[0] for D in centres])
    zmax = np.minimum(centres * 1e-3 * f_px / 1.0, 15.0)
    z0 = 0.25
    a = (W + 4) * H / f_px ** 2
    b = (W + 4) / f_px * (v * T / 1000.0)
    vol = np.where(zmax > z0, a * (zmax ** 3 - z0 ** 3) / 3 + b * (zmax ** 2 - z0 ** 2) / 2, 0.0)
    dens = 8000 * np.exp(-lam * centres) / v * vol
    # Candidates are drawn from `dens` and placed uniformly in the region the volume term describes; every candidate in it is
    # imaged (it is wide enough by construction of the region and an end point falls on the sensor) except those in the
    # two-pixel side margins, a diameter-independent fraction: the imaged diameters follow `dens` itself.
    obs = np.histogram(d_mm, bins=edges)[0].astype(float)
    cdf = np.concatenate([[0], np.cumsum((dens[1:] + dens[:-1]) / 2 * np.diff(centres))])
    exp = np.diff(np.interp(edges, centres, cdf))
    exp = exp / exp.sum() * obs.sum()
    keep = exp > 20
    chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    dof = int(keep.sum()) - 1
    assert dof >= 10 and chi2 / dof < 3.0, (chi2, dof, obs[keep][:8], exp[keep][:8])      # statistical scatter gives ~1; a 3 % shape error gives > 5
    # counts vs rate
    rates = [5, 10, 25, 50, 100]
    means, counts = [], []
    for r in rates:
        fr, m = ctx.simulate_particles(0, 40, W, H, r, exposure_ms=T, seed=3)
        means.append(m)
        counts.append(np.mean([len(f) for f in fr]))
    counts, means = np.array(counts), np.array(means)
    assert (np.diff(counts) > 0).all() and (np.diff(means) > 0).all()
    ratio = counts / means                         # imaged / candidates: geometry only, nearly independent of R
    assert ratio.max() / ratio.min() < 1.25
    growth = np.log(counts[-1] / counts[0]) / np.log(rates[-1] / rates[0])
    assert 0.6 < growth < 1.05                      # sub-linear: the Marshall-Palmer slope flattens with R
    ctx.close()
