"""The drop-in ``common`` package: import resolution next to the reference (CPU) and a full
Generator.run() against the oracle (GPU)."""
import os
import shutil
import subprocess
import sys
import tempfile
import types

import cv2
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "rain_rendering_b200", "dropin")
REF = os.environ.get("RAIN_REFERENCE_ROOT", "/root/reference")


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "main.py")), reason="reference tree not mounted")
def test_reference_main_resolves_the_dropin_generator_and_its_own_db():
    code = r"""
import sys, os
sys.path[:0] = [%r, %r, %r, %r]
os.chdir(%r)
import main, common.generator, common.db, common.bad_weather
assert common.generator.__file__.startswith(%r), common.generator.__file__
assert common.bad_weather.__file__.startswith(%r)
assert common.db.__file__.startswith(%r), common.db.__file__
assert main.Generator is common.generator.Generator
import tools.particles_simulation as tps          # what main.py:206-208 imports when a particles XML is missing
assert tps.__file__.startswith(%r), tps.__file__
assert callable(tps.process)
print("OK")
""" % (DROPIN, ROOT, os.path.join(ROOT, "oracle", "ref_shims"), REF, REF, DROPIN, DROPIN, REF, DROPIN)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def _args(paths, dataset, fallrate, seq="seq1"):
    from rain_rendering_b200 import synth
    cam = synth.CAMERAS[dataset]
    a = types.SimpleNamespace()
    a.conflict_strategy, a.rendering_strategy = "overwrite", None
    a.output, a.dataset, a.dataset_root = paths["output"], dataset, os.path.join(paths["dataset_root"], dataset)
    a.sequences = [seq]
    a.images = {seq: os.path.join(a.dataset_root, seq, "rgb")}
    a.depth = {seq: os.path.join(a.dataset_root, seq, "depth")}
    a.calib = {seq: None}
    a.particles = {seq: [paths["xml"]]}
    a.weather = [{"weather": "rain", "fallrate": fallrate}]
    a.texture = os.path.join(paths["streaks_db"], "env_light_database", "size32")
    a.norm_coeff = os.path.join(paths["streaks_db"], "env_light_database", "txt", "normalized_env_max.txt")
    a.save_envmap = False
    a.settings = dict(cam_exposure=cam["cam_exposure"], cam_gain=cam["cam_gain"], cam_focal=cam["cam_focal"], cam_f_number=cam["cam_f_number"],
                      cam_focus_plane=6.0, render_scale=1, depth_scale=1)
    a.noise_scale, a.noise_std, a.opacity_attenuation = 1.0, 2.0, 0.9
    a.frame_start, a.frame_end, a.frame_step, a.frames, a.verbose = 0, None, 1, [], False
    return a


@pytest.mark.gpu
def test_generator_run_writes_the_oracle_images():
    for k in [k for k in sys.modules if k == "common" or k.startswith("common.")]:
        del sys.modules[k]
    sys.path.insert(0, DROPIN)
    try:
        from rain_rendering_b200 import synth
        from oracle import rain_oracle as ro
        import common.generator as gen
        assert gen.__file__.startswith(DROPIN)
        root = tempfile.mkdtemp(prefix="rr_dropin_")
        W, H, nf = 384, 256, 5
        paths = synth.write_dataset(root, "customdb", "seq1", W, H, nf, 25, 1200, seed=3, n_sim_frames=2)
        a = _args(paths, "customdb", 25)
        os.environ["RAIN_B200_BATCH"] = "2"          # 5 frames -> batches of 2, 2, 1
        g = gen.Generator(a)
        g.run()
        out_dir = os.path.join(paths["output"], "customdb", "seq1", "rain", "25mm")
        tex, ratios = ro.load_streak_database(a.texture, a.norm_coeff)
        frames = ro.load_streaks_from_xml(paths["xml"], 1, W, H)
        cam = ro.Camera(W=W, H=H, fallrate=25, noise_scale=1.0, noise_std=2.0, opacity_attenuation=0.9)
        for i in range(nf):
            name = "%06d" % i
            got = cv2.imread(os.path.join(out_dir, "rainy_image", name + ".png"))
            assert got is not None and os.path.exists(os.path.join(out_dir, "rain_mask", name + ".png"))
            bg, depth = ro.read_frame(os.path.join(a.images["seq1"], name + ".png"), os.path.join(a.depth["seq1"], name + ".png"))
            o = ro.render_frame(bg, depth, frames[i % 2], tex, ratios, cam, i, f32_mode="canonical")
            assert np.abs(got.astype(int) - o.out_u8.astype(int)).max() <= 1
        # skip strategy: nothing is re-rendered
        a.conflict_strategy = "skip"
        t = os.path.getmtime(os.path.join(out_dir, "rainy_image", "000000.png"))
        gen.Generator(a).run()
        assert os.path.getmtime(os.path.join(out_dir, "rainy_image", "000000.png")) == t
        shutil.rmtree(root, ignore_errors=True)
    finally:
        sys.path.remove(DROPIN)
        for k in [k for k in sys.modules if k == "common" or k.startswith("common.")]:
            del sys.modules[k]


@pytest.mark.gpu
def test_dropin_stage_classes_match_oracle():
    for k in [k for k in sys.modules if k == "common" or k.startswith("common.")]:
        del sys.modules[k]
    sys.path.insert(0, DROPIN)
    try:
        from oracle import rain_oracle as ro
        from rain_rendering_b200 import synth
        import common.add_attenuation as att
        import common.bad_weather as bw
        import common.solid_angle as sa
        W, H = 320, 200
        bgr, depth = synth.make_frame(W, H, 11)
        cam = ro.Camera(W=W, H=H, fallrate=50)
        fog = att.FogRain(rain_intensity=50, focal=0.006, f_number=6.0, angle=90, exposure=2, camera_gain=20).fog_rain_layer(bgr / 255.0, depth)
        ofog = ro.fog_rain_layer(bgr / 255.0, depth, cam, "canonical")
        assert np.abs(fog - ofog).max() < 1e-13
        env = bw.EnvironmentMapGenerator(0.006, W, H).generate_map(ofog)
        oenv = ro.generate_map(ofog, ro.build_env_tables(W, H, 0.006))
        assert np.array_equal(env, oenv)
        om = sa.get_solid_angles(env)
        assert np.abs(om / ro.solid_angles(env.shape[0], env.shape[1]) - 1).max() < 1e-7
        with pytest.raises(NotImplementedError):
            bw.RainRenderer(0.006, 6.0, 6, 10, 165).add_drop_to_image()
        assert bw.DBManager.classify_drop(4) == bw.DropType.Big and bw.DBManager.classify_drop(1) == bw.DropType.Small
    finally:
        sys.path.remove(DROPIN)
        for k in [k for k in sys.modules if k == "common" or k.startswith("common.")]:
            del sys.modules[k]


class _FakeBuf:
    def __init__(self, shape, dtype):
        self.array = np.zeros(shape, dtype)

    def free(self):
        self.array = None


class _FakeCtx:
    """Stands in for RainContext: 'renders' u8 = bgr + 1, mask = depth * 2 when a batch is waited for, and can be told
    to report a patch-arena overflow on the n-th wait (the asynchronous API does not grow the arena itself)."""
    W, H, render_scale = 24, 16, 1

    def __init__(self, overflow_on_wait=None):
        self.q, self.log, self.waits, self.overflow_on_wait = [], [], 0, overflow_on_wait

    def _render(self, bgr, depth, mask, u8):
        u8[...] = bgr + 1
        mask[...] = depth * 2

    def render_frames(self, bgr, depth, recs, offs, out_bgr, out_mask, out_u8, want=()):
        self.log.append(("sync", len(bgr), int(offs[-1])))
        self._render(bgr, depth, out_mask, out_u8)

    def submit_frames(self, bgr, depth, recs, offs, out_bgr, out_mask, out_u8):
        assert len(self.q) < 2
        self.log.append(("submit", len(bgr), int(offs[-1])))
        self.q.append((bgr, depth, out_mask, out_u8))

    def wait_frames(self):
        from rain_rendering_b200._lib import RainError
        self.waits += 1
        item = self.q.pop(0)
        if self.overflow_on_wait == self.waits:
            raise RainError("rr_wait_frames failed (-4): patch arena overflow: need 10 float64 elements, have 5")
        self._render(*item)

    def synchronize(self):
        self.q = []


@pytest.mark.parametrize("overflow_on_wait", [None, 2])
def test_frame_pipeline_orders_batches_and_recovers_from_arena_overflow(tmp_path, overflow_on_wait):
    for k in [k for k in sys.modules if k == "common" or k.startswith("common.")]:
        del sys.modules[k]
    sys.path.insert(0, DROPIN)
    try:
        import common.generator as gen
        from PIL import Image
        from rain_rendering_b200.streaks import STREAK_DTYPE
        H, W = 16, 24
        src = tmp_path / "src"
        src.mkdir()
        frames = list(range(11))
        for i in frames:
            cv2.imwrite(str(src / ("i%03d.png" % i)), np.full((H, W, 3), i, np.uint8))
            cv2.imwrite(str(src / ("d%03d.png" % i)), np.full((H, W), i * 256, np.uint16))
        (src / "d004.png").write_bytes(b"corrupt")                                  # the frame is skipped (generator.py:361-363)
        Image.fromarray(np.full((H, W, 3), 7, np.uint8)).convert("P").save(str(src / "i007.png"))     # palette: OpenCV fallback
        fallbacks = []

        def fallback(image_file, depth_file):
            fallbacks.append(os.path.basename(image_file))
            d = cv2.imread(depth_file, cv2.IMREAD_UNCHANGED)
            if d is None:
                return None, None
            return cv2.imread(image_file), d.astype(np.float32) / 256.

        ctx = _FakeCtx(overflow_on_wait)
        pipe = gen._FramePipeline(ctx, batch=3, io_threads=4, fallback_decode=fallback, alloc=_FakeBuf)
        order = []

        def assemble(i):
            order.append(i)
            return np.zeros(i % 3 + 1, STREAK_DTYPE)

        for b0 in range(0, 11, 3):
            q = [(str(src / ("i%03d.png" % i)), str(src / ("d%03d.png" % i)), i, str(tmp_path / "rainy_image" / ("%03d.png" % i)),
                  str(tmp_path / "rain_mask" / ("%03d.png" % i))) for i in frames[b0:b0 + 3]]
            pipe.push(q, assemble)
        pipe.finish(assemble)
        pipe.close()
        assert sorted(fallbacks) == ["i004.png", "i007.png"]
        assert order == [i for i in frames if i != 4]               # records are assembled in frame order
        assert ctx.log[0][0] == "sync" and ctx.log[0][1] == 3       # the first batch sizes the arena
        assert [e[1] for e in ctx.log if e[0] == "submit"] == [2, 3, 2]
        if overflow_on_wait:
            assert [e[0] for e in ctx.log].count("sync") >= 2       # the batches in flight were re-rendered synchronously
        for i in frames:
            p = tmp_path / "rainy_image" / ("%03d.png" % i)
            if i == 4:
                assert not p.exists()
                continue
            img = cv2.imread(str(p))
            assert img is not None and np.array_equal(img, cv2.imread(str(src / ("i%03d.png" % i))) + 1), i      # the fake context renders u8 = bgr + 1
            assert (tmp_path / "rain_mask" / ("%03d.png" % i)).exists()
        assert pipe.frames_done == 10
    finally:
        sys.path.remove(DROPIN)
        for k in [k for k in sys.modules if k == "common" or k.startswith("common.")]:
            del sys.modules[k]


class _FakeSimCtx:
    """Stands in for RainContext.simulate_particles: deterministic streaks, records the calls."""

    def __init__(self):
        self.calls = []

    def simulate_particles(self, first_frame, n_frames, W, H, fallrate, **kw):
        from rain_rendering_b200.streaks import SIM_STREAK_DTYPE
        self.calls.append((first_frame, n_frames, W, H, fallrate, kw))
        frames = []
        for f in range(first_frame, first_frame + n_frames):
            rng = np.random.RandomState(1000 + f)
            n = 40 + f
            r = np.zeros(n, SIM_STREAK_DTYPE)
            r["pid"] = np.arange(n) + 7
            z = rng.uniform(0.3, 4, n)
            r["wp1"] = np.stack([rng.uniform(-1, 1, n), rng.uniform(-0.5, 0.5, n), -z], 1)
            r["wp2"] = r["wp1"] + np.stack([np.zeros(n), -rng.uniform(0.005, 0.02, n), np.zeros(n)], 1)
            r["wd1"] = r["wd2"] = rng.uniform(5e-4, 4e-3, n)
            r["ip1"] = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
            r["ip2"] = r["ip1"] + np.stack([rng.normal(0, 1, n), -rng.uniform(3, 60, n)], 1)
            r["iw1"] = r["iw2"] = rng.uniform(0.3, 6, n)
            frames.append(r)
        return frames, 123.0


def test_dropin_particles_simulation_writes_the_xml_main_py_expects(tmp_path):
    for k in [k for k in sys.modules if k == "tools" or k.startswith("tools.")]:
        del sys.modules[k]
    sys.path.insert(0, DROPIN)
    try:
        import tools.particles_simulation as ps
        from rain_rendering_b200 import streaks as S
        assert ps.__file__.startswith(DROPIN)
        opts = dict(cam_hz=10, cam_WH=[640, 480], cam_CCD_WH=[640, 480], cam_CCD_pixsize=4.65, cam_focal=6, cam_exposure=2, sim_hz=2000,
                    sim_mode="normal", sim_duration=0.6, sim_steps={}, sequences={"a": 1})
        ctx = _FakeSimCtx()
        root = str(tmp_path / "particles" / "customdb" / "seq1")
        weather = {"weather": "rain", "fallrate": 25}
        path = ps.simulate_to_xml(root, opts, weather, ctx=ctx)
        # the file main.py globs for (common/my_utils.py:172-173), one device call for the six frames of 0.6 s at 10 Hz
        import glob
        assert glob.glob(os.path.join(root, "rain", "25mm", "*_camera0.xml")) == [path]
        assert [(c[0], c[1], c[2], c[3], c[4]) for c in ctx.calls] == [(0, 6, 640, 480, 25.0)]
        assert ctx.calls[0][5]["focal_mm"] == 6.0 and ctx.calls[0][5]["sim_hz"] == 2000.0 and ctx.calls[0][5]["seed"] == 0
        assert os.path.exists(os.path.join(root, "rain", "25mm", "sim_options.json"))
        # the loader reads back exactly the records the in-memory path would build from the device output
        frames, ids = S.load_streaks_from_xml(path, 1, 640, 480, with_ids=True)
        want, _ = _FakeSimCtx().simulate_particles(0, 6, 640, 480, 25)
        assert ids == list(range(6))
        for got, w in zip(frames, want):
            ref = S.records_from_sim(w, 1, 640, 480)
            assert got.tobytes() == ref.tobytes() or all(np.array_equal(got[n], ref[n], equal_nan=True) if got[n].dtype.kind == "f" else
                                                         np.array_equal(got[n], ref[n]) for n in S.STREAK_DTYPE.names)
        # an existing simulation is kept unless forced (tools/simulation.py:262-269)
        assert ps.simulate_to_xml(root, opts, weather, ctx=ctx) is None and len(ctx.calls) == 1
        assert ps.simulate_to_xml(root, opts, weather, redo=True, ctx=ctx) == path and len(ctx.calls) == 2
        # steps mode: one camera frame per step, a parameter stays applied until changed (common/db.py:44-58)
        opts2 = dict(opts, sim_mode="steps", sim_steps={"cam_motion": [0, 0, 30, 30, 50], "rain_fallrate": [5, 5, 5, 5, 5]})
        ctx2 = _FakeSimCtx()
        ps.simulate_to_xml(str(tmp_path / "p2"), opts2, weather, ctx=ctx2)
        assert [(c[0], c[1], c[4], c[5]["cam_speed_kmh"]) for c in ctx2.calls] == [(0, 2, 5.0, 0.0), (2, 2, 5.0, 30.0), (4, 1, 5.0, 50.0)]
        # process(): the reference's signature and loops (tools/particles_simulation.py:23-42)
        ps._ctx = _FakeSimCtx()
        out = ps.process({"path": [str(tmp_path / "p3"), str(tmp_path / "p4")], "options": [opts, opts],
                          "weather": [weather, {"weather": "rain", "fallrate": 50}]}, force_recompute=True)
        assert len(out) == 4 and len(ps._ctx.calls) == 4
        ps._ctx = None
    finally:
        sys.path.remove(DROPIN)
        for k in [k for k in sys.modules if k == "tools" or k.startswith("tools.")]:
            del sys.modules[k]
