"""Run the BASELINE configs C3-C5 at full size through the library (no oracle: properties only) and
print frames/s; catches capacity problems (arena growth, record slots) that the small parity tests miss."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rain_rendering_b200 import api, synth  # noqa: E402


def run(name, batch):
    wl = synth.WORKLOADS[name]
    cam = synth.CAMERAS[wl["dataset"]]
    db, bgr, depth, recs, offs = bench.build_batch(wl, 0, batch)
    ctx = api.RainContext(0)
    ctx.set_streak_db(db.textures, db.ratios)
    ctx.set_camera(wl["W"], wl["H"], cam["cam_focal"], cam["cam_f_number"], cam["cam_exposure"], cam["cam_gain"], wl["fallrate"], 1.0, batch)
    out = ctx.render_frames(bgr, depth, recs, offs)          # may grow the arena and re-run
    t0 = time.time()
    out2 = ctx.render_frames(bgr, depth, recs, offs)
    dt = time.time() - t0
    assert np.isfinite(out["bgr"]).all() and (out["mask"] >= 0).all()
    for k in ("bgr", "mask", "u8"):
        assert np.array_equal(out[k], out2[k])
    means = np.abs(out["bgr"].astype(np.float64).mean(axis=(1, 2, 3)) - (bgr / 255.0).mean(axis=(1, 2, 3))).max()
    print("%s %dx%d %d mm/h: %d frames, %.0f streaks/frame, %.1f frames/s (pageable host buffers), mask coverage %.2f, mean-shift err %.1e, timings %s"
          % (name, wl["W"], wl["H"], wl["fallrate"], batch, offs[-1] / batch, batch / dt, (out["mask"] > 0).mean(), means,
             {k: round(v, 2) for k, v in ctx.timings().items()}))
    ctx.close()


if __name__ == "__main__":
    for name, batch in (("C3", 16), ("C4", 8), ("C5", 16)):
        run(name, batch)
