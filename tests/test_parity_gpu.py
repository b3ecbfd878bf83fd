"""Parity of the CUDA path (through the C ABI) against the oracle and the reference goldens.

Tolerances (BASELINE.json north_star):
  * rain mask: support (mask > 0) identical; values within 1 float32 ULP of float32(oracle)
  * rainy image float32: within 1 float32 ULP of float32(oracle float64) *in the value range
    of an image* -- measured as |diff| <= 1 ULP at magnitude 1 (2^-23) for |v| < 1, which is the
    ULP bound wherever it is defined without blowing up around the mean-shift zero crossing
  * rainy image uint8: within 1 LSB
The oracle runs in its "canonical" float32 mode (platform independent, see oracle/rain_oracle.py);
against the reference's own goldens (numpy / OpenCV float32 kernels of the generating host) the
bound is the documented 5e-7.
"""
import numpy as np
import pytest

from util import Scenario, golden_scenario, ulp_diff_f32

pytestmark = pytest.mark.gpu

ULP1 = 2.0 ** -23


def _check_image_ulp(bgr, ref64):
    """north_star: the float32 image within 1 ULP of float32(reference).  True float32 ULP distance wherever the value is
    at least 2^-6 in magnitude; below that (the mean shift moves values through zero, where an ULP shrinks without
    bound while the float64 chain's own absolute error does not) the bound is 1 ULP *at 2^-6*, 2^-29."""
    ref32 = ref64.astype(np.float32)
    big = np.abs(ref32) >= 2.0 ** -6
    assert ulp_diff_f32(bgr[big], ref32[big]).max() <= 1
    if (~big).any():
        assert np.abs(bgr[~big].astype(np.float64) - ref32[~big].astype(np.float64)).max() <= 2.0 ** -29


def _check_frame(out, i, o):
    mask, bgr, u8 = out["mask"][i], out["bgr"][i], out["u8"][i]
    assert np.array_equal(mask > 0, o.rain_mask > 0), "rain-mask support differs"
    assert ulp_diff_f32(mask, o.rain_mask.astype(np.float32)).max() <= 1
    _check_image_ulp(bgr, o.out_bgr)
    assert np.abs(u8.astype(int) - o.out_u8.astype(int)).max() <= 1


@pytest.mark.parametrize("W,H,n_xml,dataset,fallrate,noise", [
    (640, 480, 1500, "kitti", 10, 0.0),        # BASELINE C1 shape
    (1242, 375, 650, "kitti", 25, 0.0),        # BASELINE C2 shape (odd height)
    (512, 256, 1800, "cityscapes", 50, 2.0),   # 5 ms exposure, wind noise
    (800, 450, 700, "nuscenes", 100, 0.0),     # nuScenes optics (f/1.8: large circles of confusion)
])
def test_full_frames_match_oracle(W, H, n_xml, dataset, fallrate, noise):
    sc = Scenario(W, H, 2, n_xml, fallrate=fallrate, dataset=dataset, noise_scale=1.0 if noise else 0.0, noise_std=noise,
                  n_sim_frames=1 if noise else None)
    ctx = sc.context()
    recs, offs = sc.records()
    out = ctx.render_frames(sc.bgr, sc.depth, recs, offs)
    assert ctx.kernel_launches() > 10
    for i in range(sc.n_frames):
        o = sc.oracle_frame(i, "canonical")
        assert o.n_streaks == offs[i + 1] - offs[i] > 20
        _check_frame(out, i, o)
    ctx.close()


def test_many_streaks_per_frame_and_a_wide_environment_map():
    """More than 512 streaks in a frame (the compositor's hit lists then take several rounds) and an environment
    map wider than 2048 pixels (BASELINE C4 shape: three prefix-sum tiles per row)."""
    sc = Scenario(400, 300, 1, 9000, fallrate=100)
    ctx = sc.context()
    recs, offs = sc.records()
    assert offs[-1] > 1100
    out = ctx.render_frames(sc.bgr, sc.depth, recs, offs)
    _check_frame(out, 0, sc.oracle_frame(0, "canonical"))
    ctx.close()
    sc = Scenario(1600, 900, 1, 260, fallrate=5, dataset="nuscenes")
    ctx = sc.context()
    assert ctx.W_env > 2048
    recs, offs = sc.records()
    assert offs[-1] > 20
    out = ctx.render_frames(sc.bgr, sc.depth, recs, offs)
    _check_frame(out, 0, sc.oracle_frame(0, "canonical"))
    ctx.close()


def test_bright_frames_take_the_per_channel_in_scatter_path():
    """With KITTI optics beta_hg * E_c exceeds 1 once the mean image level passes ~0.52: the clip of
    add_attenuation.py:72 then acts on distant pixels and the three in-scatter images are no longer one image times a
    scalar, so k_fog blurs each channel on its own.  Every other scenario of this suite stays below 1."""
    sc = Scenario(320, 240, 1, 2500, fallrate=25)
    sc.bgr = np.clip(sc.bgr.astype(np.float32) * 1.35 + 10, 0, 255).astype(np.uint8)
    sc.depth = (sc.depth * 30).astype(np.float32)          # up to ~2.4 km: 1 - f_ext reaches 1, A (1 - f_ext) exceeds 1
    c = sc.cam
    g = 0.97
    beta_hg = (1 - g * g) / (4 * np.pi * (1 + g * g) ** 1.5)
    A = beta_hg * 4 * c.f_number ** 2 * (sc.bgr[0] / 255.0).reshape(-1, 3).mean(0) / (c.exposure_ms * 1e-3 * c.gain * np.pi)
    assert (A > 1.05).all(), A
    ctx = sc.context()
    o = sc.oracle_frame(0, "canonical")
    fog = ctx.fog_only(sc.bgr, sc.depth)[0]
    assert np.abs(fog - np.moveaxis(o.fog, -1, 0)).max() < 1e-13
    f_ext = np.exp(-0.312 * 25 ** 0.67 * sc.depth[0] / 1000.0)
    assert (A.min() * (1 - f_ext)).max() > 1.0                                           # the clip at :72 really acts
    recs, offs = sc.records()
    out = ctx.render_frames(sc.bgr, sc.depth, recs, offs)
    _check_frame(out, 0, o)
    ctx.close()


def test_render_scale_2_cityscapes_arrangement():
    """BASELINE C3 shape: frames arrive at twice the render size and are reduced on the device."""
    sc = Scenario(512, 256, 2, 1800, fallrate=50, dataset="cityscapes", render_scale=2)
    assert sc.bgr.shape == (2, 512, 1024, 3)
    ctx = sc.context()
    recs, offs = sc.records()
    out = ctx.render_frames(sc.bgr, sc.depth, recs, offs)
    for i in range(sc.n_frames):
        o = sc.oracle_frame(i, "canonical")
        assert o.n_streaks == offs[i + 1] - offs[i] > 20
        _check_frame(out, i, o)
    ctx.close()


def test_stage_parity_tables_fog_env_photometry():
    sc = Scenario(640, 480, 1, 1500, fallrate=25)
    ctx = sc.context()
    assert (ctx.H_env, ctx.W_env) == sc.tables.src.shape
    assert np.array_equal(ctx.debug_read("env_src"), sc.tables.src)
    om = ctx.debug_read("omega")
    assert np.abs(om / sc.omega - 1).max() < 1e-7 and abs(om.sum() - 4 * np.pi) < 1e-9
    o = sc.oracle_frame(0, "canonical")
    fog = ctx.fog_only(sc.bgr, sc.depth)[0]
    assert np.abs(fog - np.moveaxis(o.fog, -1, 0)).max() < 1e-13
    oenv = np.round(o.env * 255).astype(np.uint8)
    env = ctx.envmap_only(np.moveaxis(o.fog, -1, 0)[None])[0]
    assert np.array_equal(env, oenv)
    recs, offs = sc.records()
    ph = ctx.streak_photometry_only(oenv, recs)
    ref = np.array([[p["fov_xy_avg"][0], p["fov_xy_avg"][1], p["drop_Y"]] for p in o.per_streak])
    assert ph.shape == ref.shape and np.abs(ph / ref - 1).max() < 1e-11
    ctx.close()


def test_edge_cases_empty_ragged_and_out_of_frame():
    sc = Scenario(320, 240, 4, 2500, fallrate=25)
    ctx = sc.context()
    recs, offs = sc.records()
    # ragged batch: frame 1 has no streaks at all, frame 2 keeps only three
    keep = np.ones(len(recs), bool)
    keep[offs[1]:offs[2]] = False
    keep[offs[2] + 3:offs[3]] = False
    recs2 = recs[keep]
    offs2 = np.array([0, offs[1], offs[1], offs[1] + 3, offs[1] + 3 + (offs[4] - offs[3])], np.int32)
    out = ctx.render_frames(sc.bgr, sc.depth, recs2, offs2)
    assert (out["mask"][1] == 0).all()
    o1 = sc.oracle_frame(1, "canonical")      # oracle with streaks ...
    sc.oracle_frames[1] = []
    o1e = sc.oracle_frame(1, "canonical")     # ... and without
    _check_frame(out, 1, o1e)
    assert np.abs(o1.rain_mask).max() > 0
    # a batch with zero streaks anywhere
    out0 = ctx.render_frames(sc.bgr[:1], sc.depth[:1], recs[:0], np.array([0, 0], np.int32))
    sc.oracle_frames[0] = []
    _check_frame(out0, 0, sc.oracle_frame(0, "canonical"))
    ctx.close()


def test_degenerate_streaks_are_skipped_like_the_reference():
    """The reference wraps every streak in try/except and skips the ones that raise
    (common/generator.py:180-189): a streak at the camera centre (no view direction) and one at zero
    depth (infinite circle of confusion) must vanish on both sides, the frame must not."""
    sc = Scenario(512, 256, 1, 900, fallrate=25)
    ctx = sc.context()
    bad_a, bad_b = sc.oracle_frames[0][3], sc.oracle_frames[0][7]
    bad_a.wp1[:] = 0; bad_a.wp2[:] = 0                     # |P| = 0 -> NaN direction -> "Drop skipped"
    bad_b.wp1[2] = 0.0                                     # o = 0 -> c = inf -> int(10 c) raises
    sc.sim_frames[0]["wp1"][3] = 0; sc.sim_frames[0]["wp2"][3] = 0
    sc.sim_frames[0]["wp1"][7, 2] = 0.0
    recs, offs = sc.records()
    pids = {int(bad_a.pid), int(bad_b.pid)}
    assert pids <= set(recs["pid"].tolist()), "the degenerate streaks must be inside the frame for this test"
    out = ctx.render_frames(sc.bgr, sc.depth, recs, offs)
    o = sc.oracle_frame(0, "canonical")
    assert {p for p, _ in o.skipped} == pids
    _check_frame(out, 0, o)
    plans = ctx.debug_read("plans", 0, len(recs))
    assert sorted(recs["pid"][plans["valid"] == 0].tolist()) == sorted(pids)
    ctx.close()


def test_two_contexts_on_two_devices_in_one_process():
    """The opt-in to > 48 KB of dynamic shared memory (k_fog, k_raster) is a per-DEVICE function attribute: a process-wide
    "set once" flag left the second GPU of a process without it (round 1).  Contexts on devices 0 and 1, rendered
    alternately, must both match the single-device result.  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sc = Scenario(384, 256, 2, 900, fallrate=25)
    recs, offs = sc.records()
    ctx0 = sc.context(device=0)
    ctx1 = sc.context(device=1)
    a0 = ctx0.render_frames(sc.bgr, sc.depth, recs, offs)
    a1 = ctx1.render_frames(sc.bgr, sc.depth, recs, offs)
    b0 = ctx0.render_frames(sc.bgr, sc.depth, recs, offs)
    for k in ("bgr", "mask", "u8"):
        assert np.array_equal(a0[k], a1[k]) and np.array_equal(a0[k], b0[k])
    _check_frame(a1, 0, sc.oracle_frame(0, "canonical"))
    ctx1.close(); ctx0.close()


def test_determinism_and_batch_independence():
    sc = Scenario(512, 256, 3, 900, fallrate=25)
    ctx = sc.context()
    recs, offs = sc.records()
    a = ctx.render_frames(sc.bgr, sc.depth, recs, offs)
    b = ctx.render_frames(sc.bgr, sc.depth, recs, offs)
    for k in ("bgr", "mask", "u8"):
        assert np.array_equal(a[k], b[k])
    # frame 2 rendered alone equals frame 2 rendered inside the batch (frames are independent)
    one = ctx.render_frames(sc.bgr[2:3], sc.depth[2:3], recs[offs[2]:offs[3]], np.array([0, offs[3] - offs[2]], np.int32))
    for k in ("bgr", "mask", "u8"):
        assert np.array_equal(one[k][0], a[k][2])
    ctx.close()


def test_pipelined_submissions_equal_synchronous_renders():
    """rr_submit_frames / rr_wait_frames with two host buffer sets: every batch of a stream of
    different batches must come out exactly as the synchronous call renders it."""
    from rain_rendering_b200 import api
    scs = [Scenario(384, 256, 8, 900 + 150 * k, fallrate=25, seed=k) for k in range(3)]
    ctx = scs[0].context()
    want, sets = [], []
    for sc in scs:
        recs, offs = sc.records()
        want.append(ctx.render_frames(sc.bgr, sc.depth, recs, offs))
        hb = dict(bgr=api.PinnedBuffer(sc.bgr.shape, np.uint8), depth=api.PinnedBuffer(sc.depth.shape, np.float32),
                  recs=api.PinnedBuffer(recs.shape, recs.dtype), offs=offs,
                  out=api.PinnedBuffer(sc.bgr.shape, np.float32), mask=api.PinnedBuffer(sc.depth.shape, np.float32),
                  u8=api.PinnedBuffer(sc.bgr.shape, np.uint8))
        hb["bgr"].array[...] = sc.bgr; hb["depth"].array[...] = sc.depth; hb["recs"].array[...] = recs
        sets.append(hb)
    order = [0, 1, 2, 0, 2, 1, 1, 0]
    def submit(i):
        hb = sets[i]
        hb["out"].array[...] = 0; hb["mask"].array[...] = -1; hb["u8"].array[...] = 0
        ctx.submit_frames(hb["bgr"].array, hb["depth"].array, hb["recs"].array, hb["offs"], hb["out"].array, hb["mask"].array, hb["u8"].array)
    def check(i):
        hb = sets[i]
        assert np.array_equal(hb["out"].array, want[i]["bgr"]) and np.array_equal(hb["mask"].array, want[i]["mask"])
        assert np.array_equal(hb["u8"].array, want[i]["u8"])
    # the same host set is never in flight twice: wait before re-submitting it
    inflight = []
    for i in order:
        if i in inflight or len(inflight) == 2:
            ctx.wait_frames()
            check(inflight.pop(0))
        if i in inflight:
            ctx.wait_frames()
            check(inflight.pop(0))
        submit(i)
        inflight.append(i)
    while inflight:
        ctx.wait_frames()
        check(inflight.pop(0))
    ctx.close()


@pytest.mark.gpu
def test_lanes_round_robin_over_contexts_equals_one_context():
    """api.RainLanes: batches submitted in turn to three contexts on one GPU (up to six in flight, their kernels
    overlapping) come out bit for bit as one context renders them synchronously, in submission order."""
    from rain_rendering_b200 import api
    scs = [Scenario(384, 256, 8, 900 + 150 * k, fallrate=25, seed=k) for k in range(3)]
    one = scs[0].context()
    c = scs[0].cam
    lanes = api.RainLanes(0, 3)
    lanes.set_streak_db(scs[0].db.textures, scs[0].db.ratios)
    lanes.set_camera(scs[0].W, scs[0].H, c.focal_mm, c.f_number, c.exposure_ms, c.gain, c.fallrate, c.opacity_attenuation, scs[0].n_frames)
    assert lanes.capacity == 6
    want, inputs = [], []
    for sc in scs:
        recs, offs = sc.records()
        want.append(one.render_frames(sc.bgr, sc.depth, recs, offs))
        inputs.append((sc, recs, offs))
    order = [0, 1, 2, 2, 1, 0, 1, 2, 0, 0, 2]
    sets, inflight = [], []
    def host_set(i):
        sc, recs, offs = inputs[i]
        hb = dict(i=i, bgr=api.PinnedBuffer(sc.bgr.shape, np.uint8), depth=api.PinnedBuffer(sc.depth.shape, np.float32),
                  recs=api.PinnedBuffer(recs.shape, recs.dtype), offs=offs, out=api.PinnedBuffer(sc.bgr.shape, np.float32),
                  mask=api.PinnedBuffer(sc.depth.shape, np.float32), u8=api.PinnedBuffer(sc.bgr.shape, np.uint8))
        hb["bgr"].array[...] = sc.bgr; hb["depth"].array[...] = sc.depth; hb["recs"].array[...] = recs
        hb["out"].array[...] = 0; hb["mask"].array[...] = -1; hb["u8"].array[...] = 0
        return hb
    def retire():
        lanes.wait_frames()
        hb = inflight.pop(0)
        w = want[hb["i"]]
        assert np.array_equal(hb["out"].array, w["bgr"]) and np.array_equal(hb["mask"].array, w["mask"]) and np.array_equal(hb["u8"].array, w["u8"])
    for i in order:
        if lanes.inflight == lanes.capacity:
            retire()
        hb = host_set(i)                      # a fresh host set per submission: none is in flight twice
        lanes.submit_frames(hb["bgr"].array, hb["depth"].array, hb["recs"].array, hb["offs"], hb["out"].array, hb["mask"].array, hb["u8"].array)
        inflight.append(hb)
    with pytest.raises(api._lib.RainError):
        while True:                           # the queue refuses a seventh submission
            hb = host_set(0)
            sets.append(hb)
            lanes.submit_frames(hb["bgr"].array, hb["depth"].array, hb["recs"].array, hb["offs"], hb["out"].array, hb["mask"].array, hb["u8"].array)
            inflight.append(hb)
    while inflight:
        retire()
    with pytest.raises(api._lib.RainError):
        lanes.wait_frames()
    lanes.close(); one.close()


def test_compact_boundary_formats_uint16_depth_and_saved_mask_forms():
    """rr_frame_io: the depth PNG's uint16 samples in (divided by 256 on the device, generator.py:365) and the rain
    mask out in the forms that are saved (generator.py:467): plt.imsave's colormap index, the 16-bit normalised
    mask and the (min, max) it was normalised with -- 9 instead of 14 staged bytes per pixel."""
    from oracle import rain_oracle as ro
    sc = Scenario(384, 256, 3, 1400, fallrate=25)
    d16 = np.clip(np.rint(sc.depth * 256.0), 0, 65535).astype(np.uint16)
    sc.depth = (d16.astype(np.float32) / 256.).astype(np.float32)        # what the reference decodes from that PNG
    ctx = sc.context()
    recs, offs = sc.records()
    ref = ctx.render_frames(sc.bgr, sc.depth, recs, offs, want=("bgr", "mask", "u8", "idx8", "u16", "range"))
    got = ctx.render_frames(sc.bgr, d16, recs, offs, want=("u8", "idx8", "range"))       # only what a writer needs
    assert got["bgr"] is None and got["mask"] is None
    assert np.array_equal(got["u8"], ref["u8"]) and np.array_equal(got["idx8"], ref["idx8"]) and np.array_equal(got["range"], ref["range"])
    for i in range(sc.n_frames):
        o = sc.oracle_frame(i, "canonical")
        _check_frame(ref, i, o)
        idx, (lo, hi) = ro.imsave_mask_index(o.rain_mask)
        assert ref["range"][i, 0] == lo == 0.0 and abs(ref["range"][i, 1] / hi - 1) < 1e-12
        # the device mask equals the oracle's to ~1e-16 relative, so an index can only differ where t * 256 sits on an integer
        d = np.abs(ref["idx8"][i].astype(int) - idx.astype(int))
        assert d.max() <= 1 and (d != 0).mean() < 1e-4
        assert np.array_equal(ref["idx8"][i] > 0, idx > 0) or (d != 0).sum() < 8
        d16m = np.abs(ref["u16"][i].astype(int) - ro.mask_u16(o.rain_mask).astype(int))
        assert d16m.max() <= 1 and (d16m != 0).mean() < 1e-3
    # a frame without rain: flat mask -> all zeros, range (0, 0)
    none = ctx.render_frames(sc.bgr[:1], d16[:1], recs[:0], np.array([0, 0], np.int32), want=("idx8", "u16", "range"))
    assert (none["idx8"] == 0).all() and (none["u16"] == 0).all() and (none["range"] == 0).all()
    ctx.close()


def test_png_image_data_made_on_the_gpu_decodes_to_the_rendered_pixels(tmp_path):
    """rr_frame_io.out_png_*: the zlib streams the device emits (Sub filter + one literal-only dynamic Huffman block + Adler-32)
    must inflate -- with Python's zlib, and as framed PNG files with OpenCV's libpng -- to exactly the uint8 image and the
    viridis-coloured mask index of the same render.  Odd width (1 + 4 W bytes per scanline: rows straddle word boundaries),
    a frame without rain (a mask of one colour: two used symbols) and batches through the asynchronous path."""
    import zlib
    import cv2
    from rain_rendering_b200 import api, pngio
    W, H = 333, 200
    sc = Scenario(W, H, 5, 1400, fallrate=25)
    ctx = sc.context()
    recs, offs = sc.records()
    keep = np.ones(len(recs), bool)
    keep[offs[3]:offs[4]] = False                      # frame 3 gets no streaks
    recs = recs[keep]
    offs = np.concatenate([offs[:4], offs[4:] - (offs[4] - offs[3])]).astype(np.int32)
    ref = ctx.render_frames(sc.bgr, sc.depth, recs, offs, want=("u8", "idx8"))
    stride = ctx.png_stream_bound()
    n = sc.n_frames
    lut = pngio.viridis_rgb()

    def check(png):
        for kind in ("image", "mask"):
            for i in range(n):
                size = int(png[kind + "_sizes"][i])
                assert 0 < size <= stride
                raw = np.frombuffer(zlib.decompress(png[kind][i, :size].tobytes()), np.uint8).reshape(H, 4 * W + 1)
                assert (raw[:, 0] == 1).all()                                              # Sub filter on every scanline
                px = np.cumsum(raw[:, 1:].reshape(H, W, 4).astype(np.uint32), axis=1).astype(np.uint8)   # undo Sub
                want = ref["u8"][i][..., ::-1] if kind == "image" else lut[ref["idx8"][i]]
                assert np.array_equal(px[..., :3], want) and (px[..., 3] == 255).all(), (kind, i)
    png = dict(image=np.zeros((n, stride), np.uint8), mask=np.zeros((n, stride), np.uint8), image_sizes=np.zeros(n, np.uint32), mask_sizes=np.zeros(n, np.uint32))
    ctx.render_frames(sc.bgr, sc.depth, recs, offs, want=(), png=png)
    check(png)
    assert png["mask_sizes"][3] < png["mask_sizes"][0] and png["image_sizes"].min() > 1000
    # framed as files: what OpenCV reads back
    paths = [str(tmp_path / ("r%d.png" % i)) for i in range(n)]
    assert pngio.write_streams(paths, png["image"], png["image_sizes"], W, H, 3) == 0
    for i in range(n):
        a = cv2.imread(paths[i], cv2.IMREAD_UNCHANGED)
        assert a.shape == (H, W, 4) and np.array_equal(a[..., :3], ref["u8"][i]) and (a[..., 3] == 255).all()
    # asynchronous submissions with page-locked buffers: same streams, byte for byte
    sets = []
    for _ in range(2):
        hb = dict(bgr=api.PinnedBuffer(sc.bgr.shape, np.uint8), depth=api.PinnedBuffer(sc.depth.shape, np.float32), recs=api.PinnedBuffer(recs.shape, recs.dtype),
                  image=api.PinnedBuffer((n, stride), np.uint8), mask=api.PinnedBuffer((n, stride), np.uint8),
                  image_sizes=api.PinnedBuffer((n,), np.uint32), mask_sizes=api.PinnedBuffer((n,), np.uint32))
        hb["bgr"].array[...] = sc.bgr; hb["depth"].array[...] = sc.depth; hb["recs"].array[...] = recs
        sets.append(hb)
    for k in range(4):
        hb = sets[k & 1]
        if k >= 2:
            ctx.wait_frames()
        ctx.submit_frames(hb["bgr"].array, hb["depth"].array, hb["recs"].array, offs,
                          png=dict(image=hb["image"].array, mask=hb["mask"].array, image_sizes=hb["image_sizes"].array, mask_sizes=hb["mask_sizes"].array))
    ctx.wait_frames(); ctx.wait_frames()
    for hb in sets:
        got = dict(image=hb["image"].array, mask=hb["mask"].array, image_sizes=hb["image_sizes"].array, mask_sizes=hb["mask_sizes"].array)
        check(got)
        for kind in ("image", "mask"):
            assert np.array_equal(got[kind + "_sizes"], png[kind + "_sizes"])
            for i in range(n):
                assert np.array_equal(got[kind][i, :got[kind + "_sizes"][i]], png[kind][i, :png[kind + "_sizes"][i]])
    ctx.close()


@pytest.mark.parametrize("name", ["small_256x192", "c1_640x480"])
def test_against_reference_goldens(name):
    sc, g = golden_scenario(name)
    ctx = sc.context()
    recs, offs = sc.records()
    assert (np.diff(offs) == g["n_streaks"]).all()
    out = ctx.render_frames(sc.bgr, sc.depth, recs, offs)
    ref_rgb = g["rainy_rgb"] if "rainy_rgb" in g else g["rainy_rgb_f32"].astype(np.float64)
    ref_mask = g["rain_mask"] if "rain_mask" in g else g["rain_mask_f32"].astype(np.float64)
    got = np.clip(out["bgr"][..., ::-1].astype(np.float64), 0, 1)
    assert np.abs(got - ref_rgb).max() < 5e-7
    assert np.array_equal(out["mask"] > 0, ref_mask > 0)
    assert ulp_diff_f32(out["mask"], ref_mask.astype(np.float32)).max() <= 1
    ref_u8 = (ref_rgb * 255).astype(np.uint8)[..., ::-1].astype(int)
    assert np.abs(out["u8"].astype(int) - ref_u8).max() <= 1
    # the reference authors' own acceptance metric (scripts/check_difference.py:17-49)
    diff = np.abs(out["u8"].astype(int) - ref_u8)
    assert diff.mean() < 0.01
    ctx.close()


def test_against_reference_golden_c2_frame():
    """BASELINE C2's frame (1242x375, 25 mm/h, 609 streaks) against what the live reference produced for it: the uint8
    image it would save (rainy_u8, RGB) and -- through the float32 mask's SHA being pinned by the oracle on the CPU
    (tests/test_oracle.py) -- the mask, compared here with the oracle's float64 mask of the same frame."""
    sc, g = golden_scenario("c2_1242x375")
    ctx = sc.context()
    recs, offs = sc.records()
    assert np.diff(offs).tolist() == g["n_streaks"].tolist() == [609]
    d16 = np.ascontiguousarray(g["depth_u16"])
    out = ctx.render_frames(sc.bgr, d16, recs, offs, want=("bgr", "mask", "u8", "idx8"))
    ref_u8 = g["rainy_u8"][..., ::-1].astype(int)
    diff = np.abs(out["u8"].astype(int) - ref_u8)
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3
    assert diff.mean() < 0.01                                   # scripts/check_difference.py:17-49
    o = sc.oracle_frame(0, "native")
    assert np.array_equal(out["mask"][0] > 0, o.rain_mask > 0)
    assert ulp_diff_f32(out["mask"][0], o.rain_mask.astype(np.float32)).max() <= 1
    ctx.close()


@pytest.mark.parametrize("name,W,H,n_xml,dataset,fallrate,rs", [
    ("C3", 1024, 512, 1300, "cityscapes", 50, 2),        # 2048x1024 frames reduced on the device, 5 ms exposure
    ("C4-100", 1600, 900, 2600, "nuscenes", 100, 1),     # W_env 2373: three prefix tiles per row; f/1.8 defocus
    ("C4-200", 1600, 900, 4800, "nuscenes", 200, 1),
    ("C5", 1242, 375, 2600, "kitti", 100, 1),            # KITTI at 100 mm/h: ~1600 streaks in the frame
])
def test_baseline_configs_at_full_size_one_frame_each(name, W, H, n_xml, dataset, fallrate, rs):
    """BASELINE.json configs C3, C4 (the two heavy rates of its sweep) and C5 at their real sizes and streak counts
    (synth.WORKLOADS), one frame each against the oracle: the oracle needs tens of seconds per such frame."""
    sc = Scenario(W, H, 1, n_xml, fallrate=fallrate, dataset=dataset, render_scale=rs)
    ctx = sc.context()
    recs, offs = sc.records()
    lo = {"C3": 600, "C4-100": 1200, "C4-200": 2200, "C5": 1200}[name]
    assert offs[-1] > lo, (name, offs[-1])
    out = ctx.render_frames(sc.bgr, sc.depth, recs, offs)
    o = sc.oracle_frame(0, "canonical")
    assert o.n_streaks == offs[1]
    _check_frame(out, 0, o)
    ctx.close()


def test_full_size_properties_c2_batch():
    """BASELINE C2 at full size: properties that need no oracle."""
    sc = Scenario(1242, 375, 8, 650, fallrate=25)
    ctx = sc.context()
    recs, offs = sc.records()
    out = ctx.render_frames(sc.bgr, sc.depth, recs, offs)
    assert np.isfinite(out["bgr"]).all() and (out["mask"] >= 0).all()
    # mean shift: mean(out) == mean(bg) (generator.py:461-464)
    for i in range(sc.n_frames):
        assert abs(out["bgr"][i].astype(np.float64).mean() - (sc.bgr[i] / 255.0).mean()) < 1e-6
    # no streaks -> mask all zero and the image is the fogged frame mean-shifted: linear in nothing else
    none = ctx.render_frames(sc.bgr, sc.depth, recs[:0], np.zeros(sc.n_frames + 1, np.int32))
    assert (none["mask"] == 0).all()
    touched = out["mask"] > 0
    assert touched.mean() > 0.05
    # pixels never touched by a streak differ from the no-rain render only by the mean shift
    d = (out["bgr"].astype(np.float64) - none["bgr"].astype(np.float64))
    for i in range(sc.n_frames):
        un = ~touched[i]
        spread = d[i][un].max() - d[i][un].min()
        assert spread < 1e-6
    ctx.close()
