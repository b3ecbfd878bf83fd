"""End-to-end throughput of the drop-in Generator (PNG decode -> rr_submit_frames -> PNG encode) on a synthetic
KITTI-shaped sequence laid out like the reference's customdb tree.  Prints one JSON line per configuration.

  python tools/dropin_e2e.py [n_frames] [batch]
"""
import json
import os
import shutil
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "rain_rendering_b200", "dropin"), ROOT]

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


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 192
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    wl = synth.WORKLOADS["C2"]
    root = tempfile.mkdtemp(prefix="rr_e2e_")
    t0 = time.time()
    n_distinct = 8                                            # synthesising a frame costs 0.3 s: write a few, copy the files
    paths = synth.write_dataset(root, "customdb", "seq1", wl["W"], wl["H"], n_distinct, wl["fallrate"], wl["n_xml"], seed=5, n_sim_frames=8)
    src = os.path.join(paths["dataset_root"], "customdb", "seq1")
    for i in range(n_distinct, n_frames):
        for sub in ("rgb", "depth"):
            shutil.copyfile(os.path.join(src, sub, "%06d.png" % (i % n_distinct)), os.path.join(src, sub, "%06d.png" % i))
    t_data = time.time() - t0
    import common.generator as gen
    os.environ["RAIN_B200_BATCH"] = str(batch)
    for io_threads in (64, 16):
        os.environ["RAIN_B200_IO_THREADS"] = str(io_threads)
        a = make_args(paths, "customdb", wl["fallrate"])
        g = gen.Generator(a)
        g.run() if io_threads == 64 else None             # first pass warms the context, the page cache and the arena
        t0 = time.time()
        gen.Generator(a).run()
        dt = time.time() - t0
        print(json.dumps({"metric": "drop-in Generator.run frames/s incl. PNG decode + encode", "value": n_frames / dt, "frames": n_frames,
                          "batch": batch, "io_threads": io_threads, "size": [wl["W"], wl["H"]], "fallrate": wl["fallrate"],
                          "seconds": dt, "dataset_write_s": t_data}))
    shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
