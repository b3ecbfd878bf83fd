"""End-to-end throughput of the drop-in Generator (PNG decode -> rr_submit_frames_io -> PNG encode) on a synthetic
KITTI-shaped sequence laid out like the reference's customdb tree (what ``python main.py --dataset customdb`` drives,
main.py:230-231 -> common/generator.py:193).  Prints one JSON line per configuration; ``bench.py`` imports ``measure``
for its ``dropin_png_e2e`` key.

  python tools/dropin_e2e.py [n_frames] [batch] [io_threads,...]
"""
import json
import os
import shutil
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "rain_rendering_b200", "dropin")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from rain_rendering_b200 import synth  # noqa: E402


def make_args(paths, dataset, fallrate, seq="seq1"):
    cam = synth.CAMERAS["kitti"]
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
    a.noise_scale, a.noise_std, a.opacity_attenuation = 0.0, 0.0, 1.0
    a.frame_start, a.frame_end, a.frame_step, a.frames, a.verbose = 0, None, 1, [], False
    return a


def make_dataset(n_frames, workload="C2", n_distinct=8):
    """n_frames image + depth PNGs (n_distinct synthesised, the rest hard links to them) under a RAM-backed directory when
    there is one.  -> (root, paths)"""
    wl = synth.WORKLOADS[workload]
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    root = tempfile.mkdtemp(prefix="rr_e2e_", dir=base)
    paths = synth.write_dataset(root, "customdb", "seq1", wl["W"], wl["H"], n_distinct, wl["fallrate"], wl["n_xml"], seed=5, n_sim_frames=8)
    src = os.path.join(paths["dataset_root"], "customdb", "seq1")
    for i in range(n_distinct, n_frames):
        for sub in ("rgb", "depth"):
            a, b = os.path.join(src, sub, "%06d.png" % (i % n_distinct)), os.path.join(src, sub, "%06d.png" % i)
            try:
                os.link(a, b)
            except OSError:
                shutil.copyfile(a, b)
    return root, paths


def measure(n_frames=1024, batch=64, io_threads=None, workload="C2", paths=None, warm=True):
    """Two passes of Generator(args).run() over the sequence (the first warms the CUDA context, the page cache and the patch
    arena); the second is timed whole -- constructor, streak DB, particles XML, camera tables, every frame decoded, rendered
    and written.  ``steady`` excludes what happens once per (sequence, weather): set-up and the synchronous first batch."""
    own = paths is None
    root = None
    if own:
        root, paths = make_dataset(n_frames, workload)
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    import common.generator as gen
    wl = synth.WORKLOADS[workload]
    os.environ["RAIN_B200_BATCH"] = str(batch)
    if io_threads:
        os.environ["RAIN_B200_IO_THREADS"] = str(io_threads)
    a = make_args(paths, "customdb", wl["fallrate"])
    try:
        if warm:
            gen.Generator(a).run()
        t0 = time.time()
        g = gen.Generator(a)
        g.run()
        dt = time.time() - t0
        st = g.last_stats
        steady_frames = st["frames"] - min(batch, st["frames"])
        t_steady = st["seconds"] - st["first_batch"]
        out = {"metric": "drop-in Generator.run frames/s incl. PNG decode + encode", "value": n_frames / dt, "unit": "frames/s", "frames": n_frames,
               "seconds": dt, "steady_frames_per_s": steady_frames / t_steady if t_steady > 0 and steady_frames > 0 else None,
               "batch": batch, "io_threads": g.io_threads, "host_cores": os.cpu_count(), "size": [wl["W"], wl["H"]], "fallrate": wl["fallrate"],
               "output_format": g.output_format, "gpu_png": bool(g.gpu_png and g.output_format == "reference"),
               "input": "8-bit RGB + 16-bit depth PNG files", "setup_s": round(dt - st["seconds"], 4),
               "value_note": "whole second pass of Generator(args).run(): constructor, streak DB, particles XML, camera tables, every frame decoded, rendered, written; "
                             "page-locked buffers and the CUDA context are reused from the first pass (kept per process)",
               "waits_s": {k: round(v, 4) for k, v in st.items() if k not in ("frames",)},
               "tmp": os.path.dirname(paths["output"])}
    finally:
        if own and root:
            shutil.rmtree(root, ignore_errors=True)
    return out


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    threads = [int(t) or None for t in sys.argv[3].split(",")] if len(sys.argv) > 3 else [None]
    root, paths = make_dataset(n_frames)
    try:
        for k, t in enumerate(threads):
            print(json.dumps(measure(n_frames, batch, t, paths=paths, warm=(k == 0))))
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
