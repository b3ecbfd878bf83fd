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


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "main.py")), reason="reference tree not mounted")
def test_reference_main_check_arg_runs_for_nuscenes_through_the_repaired_plugin(tmp_path):
    """BASELINE C4 through the reference's own main.check_arg.  Upstream, ``import config.nuscenes`` already
    fails (config/nuscenes.py:4), then :28 and :56; with the drop-in first on the path the repaired plug-in resolves the
    frames (file lists that answer os.path.exists like a folder, main.py:152-158), and the untouched rest of check_arg
    finds the particle files for every rate of the sweep."""
    from rain_rendering_b200 import synth
    root = str(tmp_path)
    W, H, nf = 320, 180, 4
    paths = synth.write_dataset(root, "nuscenes", "scene-0001", W, H, nf, 25, 200, seed=2, n_sim_frames=2)
    ds = os.path.join(paths["dataset_root"], "nuscenes")
    cam = os.path.join(ds, "samples", "CAM_FRONT")
    os.makedirs(cam)
    depth_root = os.path.join(root, "depth", "nuscenes")
    os.makedirs(depth_root)
    rel = []
    for i in range(nf):
        img = cv2.imread(os.path.join(ds, "scene-0001", "rgb", "%06d.png" % i))
        name = "n015-2018-cam_front__%04d" % i
        cv2.imwrite(os.path.join(cam, name + ".jpg"), img)
        np.save(os.path.join(depth_root, name + ".npy"), np.full((H, W), 10.0, np.float32))
        rel.append(os.path.join("samples", "CAM_FRONT", name + ".jpg"))
    with open(os.path.join(ds, "rain_b200_index.json"), "w") as f:
        import json
        json.dump({"scenes": {"scene-0001": rel[:3], "scene-0002": rel[3:]}}, f)
    # particle files for two rates of the sweep, where main.py globs for them (main.py:176-209, my_utils.particles_path)
    for seq in ("scene-0001", "scene-0002"):
        for rate in (5, 100):
            d = os.path.join(paths["particles"], "nuscenes", seq, "rain", "%dmm" % rate)
            os.makedirs(d, exist_ok=True)
            shutil.copyfile(paths["xml"], os.path.join(d, "sim_camera0.xml"))
    code = r"""
import sys, os
sys.path[:0] = [%r, %r, %r, %r]
os.chdir(%r)
import numpy as np
np.int = int; np.float = float
import main, config.nuscenes as cn
assert cn.__file__.startswith(%r), cn.__file__
import config.kitti as ck
assert ck.__file__.startswith(%r), ck.__file__
a = main.check_arg(["--dataset", "nuscenes", "--dataset_root", %r, "--depth", %r, "--particles", %r, "--streaks_db", %r,
                    "--intensity", "5,100", "--output", %r, "--noverbose"])
assert list(a.sequences) == ["scene-0001", "scene-0002"], a.sequences
assert len(a.images["scene-0001"]) == 3 and len(a.images["scene-0002"]) == 1 and a.images["scene-0001"][0].endswith(".jpg")
assert a.depth["scene-0001"][0].endswith(".npy") and os.path.isfile(a.depth["scene-0001"][0])
assert [w["fallrate"] for w in a.weather] == [5, 100]
assert all(len(a.particles[s]) == 2 and all(p.endswith("_camera0.xml") for p in a.particles[s]) for s in a.sequences)
assert a.settings["cam_focal"] == 5.5 and a.settings["cam_f_number"] == 1.8 and a.settings["render_scale"] == 1
print("OK")
""" % (DROPIN, ROOT, os.path.join(ROOT, "oracle", "ref_shims"), REF, REF, DROPIN, REF, paths["dataset_root"], os.path.join(root, "depth"),
       paths["particles"], paths["streaks_db"], os.path.join(root, "out"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


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
        a.save_envmap = True
        g = gen.Generator(a)
        g.run()
        assert g.last_stats["frames"] == nf
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
            # the reference's file formats (plt.imsave): RGBA, the mask coloured through viridis from its normalised index
            rgba = cv2.imread(os.path.join(out_dir, "rainy_image", name + ".png"), cv2.IMREAD_UNCHANGED)
            assert rgba.shape == (H, W, 4) and (rgba[..., 3] == 255).all()
            m = cv2.imread(os.path.join(out_dir, "rain_mask", name + ".png"), cv2.IMREAD_UNCHANGED)
            idx, _ = ro.imsave_mask_index(o.rain_mask)
            from rain_rendering_b200 import pngio
            lut = pngio.viridis_rgb()
            same = (m[..., 2::-1] == lut[idx]).all(-1)
            near = (m[..., 2::-1] == lut[np.minimum(idx.astype(int) + 1, 255)]).all(-1) | (m[..., 2::-1] == lut[np.maximum(idx.astype(int) - 1, 0)]).all(-1)
            assert m.shape == (H, W, 4) and (same | near).all() and (~same).mean() < 1e-4
            env_file = os.path.join(paths["output"], "customdb", "seq1", "envmap", name + ".png")
            env = cv2.imread(env_file, cv2.IMREAD_UNCHANGED)
            want_env = ((np.round(o.env * 255).astype(np.uint8) / 255.0) * 255).astype(np.uint8)       # plt.imsave of BGR_env_map[..., ::-1]
            assert env is not None and env.shape[2] == 4 and np.array_equal(env[..., :3], want_env)
        # the compact pair of files: RGB image + the normalised mask as 16-bit gray
        a.save_envmap = False
        a.output = os.path.join(root, "out_compact")
        os.environ["RAIN_B200_OUTPUT_FORMAT"] = "compact"
        try:
            gen.Generator(a).run()
        finally:
            del os.environ["RAIN_B200_OUTPUT_FORMAT"]
        cdir = os.path.join(a.output, "customdb", "seq1", "rain", "25mm")
        for i in range(nf):
            name = "%06d" % i
            c = cv2.imread(os.path.join(cdir, "rainy_image", name + ".png"), cv2.IMREAD_UNCHANGED)
            assert c.shape == (H, W, 3) and np.array_equal(c, cv2.imread(os.path.join(out_dir, "rainy_image", name + ".png")))
            m16 = cv2.imread(os.path.join(cdir, "rain_mask", name + ".png"), cv2.IMREAD_UNCHANGED)
            assert m16.dtype == np.uint16 and m16.shape == (H, W) and m16.max() == 65535 and m16.min() == 0
        a.output = paths["output"]
        with pytest.raises(NotImplementedError):                   # 'white' / 'naive_db' are not silently rendered as the full model
            a.rendering_strategy = "white"
            gen.Generator(a)
        a.rendering_strategy = None
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
def test_generator_run_nuscenes_arrangement_jpg_images_and_npy_depth():
    """The nuScenes branch of Generator.run (generator.py:235-246,306-311): images and depth maps arrive as per-sequence FILE
    LISTS, CAM_FRONT frames are .jpg (decoded by OpenCV: the native codec is PNG only), depth is float32 .npy, the size is
    probed with cv2.imread and the simulator frames are spread over the files with np.linspace."""
    for k in [k for k in sys.modules if k == "common" or k.startswith("common.")]:
        del sys.modules[k]
    sys.path.insert(0, DROPIN)
    try:
        from rain_rendering_b200 import synth
        from oracle import rain_oracle as ro
        import common.generator as gen
        root = tempfile.mkdtemp(prefix="rr_nusc_")
        W, H, nf = 400, 224, 5
        paths = synth.write_dataset(root, "nuscenes", "scene-0001", W, H, nf, 25, 600, seed=4, n_sim_frames=3)
        src = os.path.join(paths["dataset_root"], "nuscenes", "scene-0001")
        imgs, deps = [], []
        for i in range(nf):
            png = os.path.join(src, "rgb", "%06d.png" % i)
            jpg = os.path.join(src, "rgb", "n015-cam_front-%06d.jpg" % i)
            cv2.imwrite(jpg, cv2.imread(png), [cv2.IMWRITE_JPEG_QUALITY, 95])
            os.remove(png)
            d = cv2.imread(os.path.join(src, "depth", "%06d.png" % i), cv2.IMREAD_UNCHANGED).astype(np.float32) / 256.
            npy = os.path.join(src, "depth", "n015-cam_front-%06d.npy" % i)
            np.save(npy, d)
            imgs.append(jpg); deps.append(npy)
        a = _args(paths, "nuscenes", 25, seq="scene-0001")
        a.images, a.depth = {"scene-0001": imgs}, {"scene-0001": deps}
        a.noise_scale, a.noise_std = 0.0, 0.0
        os.environ["RAIN_B200_BATCH"] = "4"
        gen.Generator(a).run()
        out_dir = os.path.join(paths["output"], "nuscenes", "scene-0001", "rain", "25mm")
        tex, ratios = ro.load_streak_database(a.texture, a.norm_coeff)
        frames = ro.load_streaks_from_xml(paths["xml"], 1, W, H)
        c = synth.CAMERAS["nuscenes"]
        cam = ro.Camera(W=W, H=H, focal_mm=c["cam_focal"], f_number=c["cam_f_number"], exposure_ms=c["cam_exposure"], gain=c["cam_gain"],
                        fallrate=25, opacity_attenuation=0.9)
        render_ix = np.linspace(0, len(frames), nf, endpoint=False, dtype=int)
        for i in range(nf):
            got = cv2.imread(os.path.join(out_dir, "rainy_image", "n015-cam_front-%06d.png" % i))
            assert got is not None
            o = ro.render_frame(cv2.imread(imgs[i]), np.load(deps[i]), frames[int(render_ix[i]) % len(frames)], tex, ratios, cam, int(render_ix[i]),
                                f32_mode="canonical")
            assert np.abs(got.astype(int) - o.out_u8.astype(int)).max() <= 1
        # a float64 depth map would change the reference's arithmetic (generator.py:367): refused, never downcast
        np.save(deps[0], np.load(deps[0]).astype(np.float64))
        with pytest.raises(NotImplementedError):
            gen.Generator(a).run()
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
    """Stands in for RainContext: 'renders' u8 = bgr + 1, mask index = low byte of the depth sample when a batch is waited
    for, and can be told to report a patch-arena overflow on the n-th wait (the asynchronous API does not grow the arena
    itself)."""
    W, H, render_scale = 24, 16, 1

    def __init__(self, overflow_on_wait=None):
        self.q, self.log, self.waits, self.overflow_on_wait = [], [], 0, overflow_on_wait

    def _render(self, bgr, depth, out):
        out["out_u8"][...] = bgr + 1
        out["out_idx8"][...] = (depth >> 8).astype(np.uint8)
        out["out_range"][...] = 0

    def render_frames(self, bgr, depth, recs, offs, want=(), **out):
        assert depth.dtype == np.uint16 and len(recs) == offs[-1]
        self.log.append(("sync", len(bgr), int(offs[-1])))
        self._render(bgr, depth, out)

    def submit_frames(self, bgr, depth, recs, offs, **out):
        assert len(self.q) < 2 and len(recs) == offs[-1]
        self.log.append(("submit", len(bgr), int(offs[-1])))
        self.q.append((bgr, depth, out))

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

        def fallback(image_file, depth_file, depth_u16):
            assert depth_u16
            fallbacks.append(os.path.basename(image_file))
            d = cv2.imread(depth_file, cv2.IMREAD_UNCHANGED)
            if d is None:
                return None, None
            return cv2.imread(image_file), d

        ctx = _FakeCtx(overflow_on_wait)
        pipe = gen._FramePipeline(ctx, batch=3, io_threads=4, fallback_decode=fallback, alloc=_FakeBuf)
        order = []

        def assemble(indices, record_buffer):
            order.extend(indices)
            counts = [i % 3 + 1 for i in indices]
            buf = record_buffer(sum(counts))
            assert buf.dtype == STREAK_DTYPE and len(buf) >= sum(counts)
            return buf[:sum(counts)], np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)

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
            m = cv2.imread(str(tmp_path / "rain_mask" / ("%03d.png" % i)), cv2.IMREAD_UNCHANGED)     # RGBA: viridis of the index the fake "rendered"
            from rain_rendering_b200 import pngio
            assert m.shape == (H, W, 4) and (m[..., 2::-1] == pngio.viridis_rgb()[i]).all() and (m[..., 3] == 255).all()
        assert pipe.frames_done == 10 and set(pipe.stats) >= {"decode_wait", "gpu_wait", "write_wait", "assemble"}
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


def _import_dropin_bad_weather():
    for k in [k for k in sys.modules if k == "common" or k.startswith("common.")]:
        del sys.modules[k]
    sys.path.insert(0, DROPIN)
    try:
        import common.bad_weather as bw
    finally:
        sys.path.remove(DROPIN)
    return bw


def test_dropin_geometry_calls_answer_like_the_oracle():
    """compute_circle, warping_points and FovComputation.compute_fov_plane_points of the drop-in return the reference's
    values (oracle restatement, pinned to the live reference by tests/test_oracle.py); the polygon goes through
    rr_host_fov_polygon, the header code the device path compiles."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import Scenario
    from oracle import rain_oracle as ro
    try:
        bw = _import_dropin_bad_weather()
        sc = Scenario(1242, 375, 1, 600, dataset="kitti", seed=11)
        cam = sc.cam
        env_shape = (cam.H, sc.tables.W_env, 3)
        rr = bw.RainRenderer(cam.focal_m, cam.f_number, cam.focus_plane, cam.radius, cam.fov_deg)
        fc = bw.FovComputation(camera=np.array([0, 0, 0]))
        n_big = n24 = 0
        streaks = sc.oracle_frames[0]
        assert len(streaks) > 300
        for s in streaks:
            d = bw.Streak()
            d.world_position_start, d.world_position_end = s.wp1.copy(), s.wp2.copy()
            d.image_position_start, d.image_position_end = s.ip1.copy(), s.ip2.copy()
            d.image_diameter_start, d.image_diameter_end, d.max_width = s.iw1, s.iw2, s.max_width
            pts, pts3d, pos, direction = fc.compute_fov_plane_points(d, cam.radius, cam.fov_deg, 20, env_shape)
            ref = ro.fov_polygon(s, cam, env_shape)
            assert pts.shape == ref.shape and (len(ref) == 0 or np.abs(pts - ref).max() < 1e-9)
            n24 += len(ref) == 24
            P = (s.wp1 + s.wp2) / 2
            P[1], P[2] = P[2], P[1].copy()
            assert np.array_equal(pos, P) and np.allclose(direction, P / np.linalg.norm(P), rtol=0, atol=1e-15)
            o = abs(P[1])
            if cam.pix_size == 4.65e-06:
                assert rr.compute_circle(o) == ro.circle_of_confusion_px(o, cam)
            if s.drop_type == ro.BIG:
                n_big += 1
                tex = sc.db.textures[0]
                p1, p2, maxC, minC = bw.RainRenderer.warping_points(d, tex, cam.W, cam.H)
                q1, q2, qmax, qmin = ro.warping_points(s, tex.shape[1], tex.shape[0], cam.W, cam.H)
                assert np.array_equal(p1, q1) and np.array_equal(p2, q2) and np.array_equal(maxC, qmax) and np.array_equal(minC, qmin)
        assert n_big > 5
        assert rr.compute_circle(3.0, is_infinity=True) == cam.focal_m ** 2 / (cam.f_number * 3.0)
        with pytest.raises(NotImplementedError):
            fc.compute_fov_plane_points(d, cam.radius, cam.fov_deg, 16, env_shape)
        with pytest.raises(NotImplementedError):
            bw.FovComputation(camera=np.array([0, 1, 0])).compute_fov_plane_points(d, cam.radius, cam.fov_deg, 20, env_shape)
        with pytest.raises(NotImplementedError):
            rr.add_drop_to_image()
    finally:
        for k in [k for k in sys.modules if k == "common" or k.startswith("common.")]:
            del sys.modules[k]


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "main.py")), reason="reference tree not mounted")
def test_dropin_geometry_calls_answer_like_the_live_reference():
    """The same three calls against the UNTOUCHED reference classes, in a subprocess each side (both packages are called
    ``common``)."""
    code = r"""
import sys, os, pickle
import numpy as np
side, ROOT, DROPIN, REF = sys.argv[1:5]
sys.path.insert(0, ROOT)
if side == "ref":
    from oracle import ref_harness
    bw = ref_harness.load_reference()["bw"]
else:
    sys.path.insert(0, DROPIN)
    import common.bad_weather as bw
rng = np.random.RandomState(3)
out = []
rr = bw.RainRenderer(0.006, 6.0, 6.0, 10.0, 165.0)
fc = bw.FovComputation(camera=np.array([0, 0, 0]))
for i in range(200):
    d = bw.Streak()
    c = np.array([rng.uniform(-6, 6), rng.uniform(-2, 3), rng.uniform(0.5, 25)])
    d.world_position_start = c + rng.uniform(-0.05, 0.05, 3)
    d.world_position_end = c + rng.uniform(-0.05, 0.05, 3)
    d.image_position_start = rng.randint(-20, 1300, 2)
    d.image_position_end = d.image_position_start + rng.randint(-15, 60, 2)
    d.image_diameter_start, d.image_diameter_end = rng.uniform(4, 12), rng.uniform(4, 12)
    pts, _, pos, direction = fc.compute_fov_plane_points(d, 10.0, 165.0, 20, (375, 1909, 3))
    p1, p2, maxC, minC = bw.RainRenderer.warping_points(d, np.zeros((200 + i, 32)), 1242, 375)
    out.append((np.asarray(pts), np.asarray(pos), np.asarray(direction), rr.compute_circle(abs(c[2])), rr.compute_circle(abs(c[2]), True),
                np.asarray(p1), np.asarray(p2), np.asarray(maxC, float), np.asarray(minC, float)))
pickle.dump(out, open(sys.argv[5], "wb"))
"""
    res = {}
    for side in ("ref", "dropin"):
        path = os.path.join(tempfile.gettempdir(), "rr_geom_%s_%d.pkl" % (side, os.getpid()))
        r = subprocess.run([sys.executable, "-c", code, side, ROOT, DROPIN, REF, path], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        import pickle
        res[side] = pickle.load(open(path, "rb"))
        os.remove(path)
    n24 = 0
    for a, b in zip(res["ref"], res["dropin"]):
        assert a[0].shape == b[0].shape and (a[0].size == 0 or np.abs(a[0] - b[0]).max() < 1e-9)
        n24 += len(a[0]) == 24
        assert np.array_equal(a[1], b[1]) and np.abs(a[2] - b[2]).max() < 1e-15
        assert a[3] == b[3] and a[4] == b[4]
        for k in (5, 6, 7, 8):
            assert np.array_equal(a[k], b[k])
    assert n24 > 0
