"""Experiment: end-to-end throughput (host buffers, copies inside) when batches alternate over several independent contexts on one
GPU, each with its own streams, device buffers and page-locked host sets, against one context with two submissions in flight.
  python tools/exp_lanes_e2e.py [lanes] [in flight per lane] [steps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rain_rendering_b200 import api, synth  # noqa: E402
from rain_rendering_b200.streaks import STREAK_DTYPE  # noqa: E402


def main():
    lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    batch = 64
    wl = synth.WORKLOADS["C2"]
    W, H = wl["W"], wl["H"]
    cam = synth.CAMERAS[wl["dataset"]]
    db, bgr, depth_f, d16, sim, recs, offs = bench.build_batch(wl, 0, batch)
    seeds = list(range(batch))
    L = []
    for k in range(lanes):
        ctx = api.RainContext(0)
        ctx.set_streak_db(db.textures, db.ratios)
        ctx.set_camera(W, H, cam["cam_focal"], cam["cam_f_number"], cam["cam_exposure"], cam["cam_gain"], wl["fallrate"], 1.0, batch)
        sets = []
        for _ in range(depth + 1):
            hb = dict(bgr=api.PinnedBuffer(bgr.shape, np.uint8), depth=api.PinnedBuffer(d16.shape, np.uint16),
                      recs=api.PinnedBuffer((len(recs) + 1024,), STREAK_DTYPE), idx8=api.PinnedBuffer((batch, H, W), np.uint8),
                      u8=api.PinnedBuffer((batch, H, W, 3), np.uint8), rng=api.PinnedBuffer((batch, 2), np.float64))
            hb["bgr"].array[...] = bgr; hb["depth"].array[...] = d16
            sets.append(hb)
        ctx.render_frames(bgr, d16, recs, np.ascontiguousarray(offs), want=("u8", "idx8"))      # sizes the arena
        L.append(dict(ctx=ctx, sets=sets, turn=0, inflight=0))

    def run(n):
        for i in range(n):
            ln = L[i % lanes]
            hb = ln["sets"][ln["turn"]]
            ln["turn"] = (ln["turn"] + 1) % len(ln["sets"])
            r, o = api.assemble_batch(sim, seeds, W, H, db.ratios, out=hb["recs"].array)
            if ln["inflight"] == depth:
                ln["ctx"].wait_frames(); ln["inflight"] -= 1
            ln["ctx"].submit_frames(hb["bgr"].array, hb["depth"].array, r, o, None, None, hb["u8"].array, hb["idx8"].array, None, hb["rng"].array)
            ln["inflight"] += 1
        for ln in L:
            while ln["inflight"]:
                ln["ctx"].wait_frames(); ln["inflight"] -= 1

    run(2 * lanes * (depth + 1))
    t0 = time.perf_counter()
    run(steps)
    dt = time.perf_counter() - t0
    ref = L[0]["sets"][0]["idx8"].array
    same = all(np.array_equal(ref, hb["idx8"].array) for ln in L for hb in ln["sets"])
    print("lanes %d, %d in flight each, sub-batches %s: %.3f ms per 64 frames, %.0f frames/s, outputs equal %s"
          % (lanes, depth, os.environ.get("RR_SUB_BATCHES_ASYNC", "2"), 1000 * dt / steps, batch * steps / dt, same))


if __name__ == "__main__":
    main()
