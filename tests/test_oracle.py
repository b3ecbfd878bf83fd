"""The oracle (oracle/rain_oracle.py) pinned against the reference: committed golden vectors
produced by the untouched reference (oracle/make_golden.py) and, when /root/reference is present,
the live reference itself."""
import hashlib
import os
import shutil
import tempfile

import numpy as np
import pytest

from oracle import clipper_rect, rain_oracle as ro, ref_harness
from oracle.make_golden import host_signature
from util import Scenario, golden_scenario


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _render_all(sc, mode):
    outs = [sc.oracle_frame(i, mode) for i in range(sc.n_frames)]
    rainy = np.stack([np.clip(o.out_bgr[..., ::-1], 0, 1) for o in outs])     # what plt.imsave receives
    mask = np.stack([o.rain_mask for o in outs])
    return outs, rainy, mask


def test_oracle_native_reproduces_reference_golden_small():
    sc, g = golden_scenario("small_256x192")
    outs, rainy, mask = _render_all(sc, "native")
    assert [o.n_streaks for o in outs] == g["n_streaks"].tolist()
    same_host = str(g["host"]) == host_signature()
    # mask and streak geometry never touch the float32 stages: always bit-exact
    assert np.array_equal(mask, g["rain_mask"])
    if same_host:
        assert np.array_equal(outs[0].fog, g["fog0"])
        assert np.array_equal(np.round(outs[0].env * 255).astype(np.uint8), g["env0_u8"])
        assert np.array_equal(rainy, g["rainy_rgb"])
    else:  # numpy's float32 exp / OpenCV's float32 filter differ between CPUs (see oracle docstring)
        assert np.abs(rainy - g["rainy_rgb"]).max() < 5e-7


def test_oracle_native_reproduces_reference_golden_c1():
    sc, g = golden_scenario("c1_640x480")
    outs, rainy, mask = _render_all(sc, "native")
    assert [o.n_streaks for o in outs] == g["n_streaks"].tolist()
    assert _sha(mask) == str(g["mask_sha"])
    if str(g["host"]) == host_signature():
        assert _sha(rainy) == str(g["rainy_sha"])
    assert np.abs(rainy - g["rainy_rgb_f32"]).max() < 5e-7


def test_oracle_native_reproduces_reference_golden_c2_frame():
    """BASELINE C2's frame shape (1242x375, odd height, KITTI optics) at 25 mm/h, 609 streaks: the fixture holds the
    SHA-256 of the reference's float64 outputs and the uint8 image it would save."""
    sc, g = golden_scenario("c2_1242x375")
    outs, rainy, mask = _render_all(sc, "native")
    assert [o.n_streaks for o in outs] == g["n_streaks"].tolist() == [609]
    assert _sha(mask) == str(g["mask_sha"])                      # never touches the float32 stages: always bit-exact
    if str(g["host"]) == host_signature():
        assert _sha(rainy) == str(g["rainy_sha"])
    u8 = (rainy * 255).astype(np.uint8)
    assert np.abs(u8.astype(int) - g["rainy_u8"].astype(int)).max() <= 1 and (u8 != g["rainy_u8"]).mean() < 1e-4


def test_canonical_mode_stays_within_float32_noise_of_native():
    sc, g = golden_scenario("small_256x192")
    _, rainy_c, mask_c = _render_all(sc, "canonical")
    sc2, _ = golden_scenario("small_256x192")
    outs_n, rainy_n, mask_n = _render_all(sc2, "native")
    assert np.array_equal(mask_c, mask_n)
    assert np.abs(rainy_c - rainy_n).max() < 5e-7          # a few float32 ULPs of the extinction
    u8c = (rainy_c * 255).astype(np.uint8).astype(int)
    u8n = (rainy_n * 255).astype(np.uint8).astype(int)
    assert np.abs(u8c - u8n).max() <= 1


@pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not mounted")
def test_oracle_bit_exact_against_live_reference():
    from rain_rendering_b200 import synth
    root = tempfile.mkdtemp(prefix="rr_live_")
    try:
        W, H = 192, 160
        paths = synth.write_dataset(root, "customdb", "seq1", W, H, 2, 10, 900, seed=5, n_sim_frames=1)
        ref = ref_harness.run_reference(paths, "customdb", 10, noise_scale=1.0, noise_std=2.0)
        cam = ro.Camera(W=W, H=H, fallrate=10, noise_scale=1.0, noise_std=2.0)
        tex, ratios = ro.load_streak_database(os.path.join(paths["streaks_db"], "env_light_database", "size32"),
                                              os.path.join(paths["streaks_db"], "env_light_database", "txt", "normalized_env_max.txt"))
        frames = ro.load_streaks_from_xml(paths["xml"], 1, W, H)
        for i, name in enumerate(sorted(ref)):
            bg, depth = ro.read_frame(os.path.join(root, "source/customdb/seq1/rgb", name + ".png"),
                                      os.path.join(root, "source/customdb/seq1/depth", name + ".png"))
            r = ro.render_frame(bg, depth, frames[i % len(frames)], tex, ratios, cam, i)
            assert np.array_equal(r.fog, ref[name]["fog"])
            assert np.array_equal(r.env, ref[name]["env"])
            assert np.array_equal(np.clip(r.out_bgr[..., ::-1], 0, 1), ref[name]["rainy_rgb"])
            assert np.array_equal(r.rain_mask, ref[name]["rain_mask"])
            assert r.n_streaks == len(ref[name]["streaks"]) > 5
    finally:
        shutil.rmtree(root, ignore_errors=True)


@pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not mounted")
def test_oracle_bit_exact_against_live_reference_render_scale_2():
    """Cityscapes arrangement (config/cityscapes.py:41-42): frames at twice the render size, depth at
    render size, simulator positions at sensor resolution."""
    from rain_rendering_b200 import synth
    root = tempfile.mkdtemp(prefix="rr_live2_")
    try:
        W, H = 192, 128
        paths = synth.write_dataset(root, "customdb", "seq1", W, H, 1, 50, 900, seed=6, render_scale=2)
        ref = ref_harness.run_reference(paths, "customdb", 50, settings_override={"render_scale": 2, "depth_scale": 2})
        cam = ro.Camera(W=W, H=H, fallrate=50)
        tex, ratios = ro.load_streak_database(os.path.join(paths["streaks_db"], "env_light_database", "size32"),
                                              os.path.join(paths["streaks_db"], "env_light_database", "txt", "normalized_env_max.txt"))
        frames = ro.load_streaks_from_xml(paths["xml"], 2, W, H)
        name = sorted(ref)[0]
        bg, depth = ro.read_frame(os.path.join(root, "source/customdb/seq1/rgb", name + ".png"),
                                  os.path.join(root, "source/customdb/seq1/depth", name + ".png"))
        assert bg.shape[:2] == (2 * H, 2 * W) and depth.shape == (H, W)
        r = ro.render_frame(bg, depth, frames[0], tex, ratios, cam, 0, render_scale=2)
        assert np.array_equal(r.fog, ref[name]["fog"])
        assert np.array_equal(np.clip(r.out_bgr[..., ::-1], 0, 1), ref[name]["rainy_rgb"])
        assert np.array_equal(r.rain_mask, ref[name]["rain_mask"]) and r.n_streaks == len(ref[name]["streaks"]) > 5
    finally:
        shutil.rmtree(root, ignore_errors=True)


def test_env_tables_match_reference_formula_for_all_baseline_sizes():
    # shapes quoted in SURVEY.md section 8: C1 480x985, C2 375x1909, C3 512x1573, C4 900x2373
    for (W, H, f_mm, want) in [(640, 480, 6.0, 985), (1242, 375, 6.0, 1909), (1024, 512, 6.0, 1573), (1600, 900, 5.5, 2373)]:
        t = ro.build_env_tables(W, H, f_mm / 1000.)
        assert t.W_env == want and t.src.shape == (H, want)
        assert t.src.max() < W * H


def test_clipper_restatement_basics():
    sq = [(1.9, 1.2), (10.7, 1.9), (10.2, 8.8), (1.1, 8.1)]
    out = clipper_rect.intersect_with_rect(sq, 20, 20)
    assert len(out) == 1 and sorted(map(tuple, out[0])) == sorted([(1, 1), (10, 1), (10, 8), (1, 8)])
    assert clipper_rect.area2(out[0]) > 0
    # collinear and duplicate vertices disappear
    out = clipper_rect.intersect_with_rect([(0, 0), (5, 0), (10, 0), (10, 10), (10, 10), (0, 10)], 20, 20)
    assert len(out[0]) == 4
    # clipped against the rectangle
    out = clipper_rect.intersect_with_rect([(-5, 2), (30, 2), (30, 8), (-5, 8)], 20, 20)
    xs = [p[0] for p in out[0]]
    assert min(xs) == 0 and max(xs) == 20
    # degenerate -> nothing
    assert clipper_rect.intersect_with_rect([(1, 1), (2, 2), (3, 3)], 20, 20) == []


def test_solid_angles_sum_to_sphere():
    om = ro.solid_angles(48, 97)
    assert abs(om.sum() - 4 * np.pi) < 1e-9 and (om > 0).all()
