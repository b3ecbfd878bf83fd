"""Experiment: does running two half-batches through two independent contexts (own streams, own buffers) on one GPU beat one
64-frame batch through one context?  (How much the kernels of different frames' stages fill each other's idle issue slots.)
  python tools/exp_two_contexts.py [lanes] [frames per lane]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rain_rendering_b200 import _lib, api, synth  # noqa: E402


def main():
    import torch
    lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    wl = synth.WORKLOADS["C2"]
    W, H = wl["W"], wl["H"]
    cam = synth.CAMERAS[wl["dataset"]]
    dev = torch.device("cuda", 0)
    ctxs, ios, keep = [], [], []
    for k in range(lanes):
        db, bgr, depth, d16, sim, recs, offs = bench.build_batch(wl, k, per)
        ctx = api.RainContext(0)
        ctx.set_streak_db(db.textures, db.ratios)
        ctx.set_camera(W, H, cam["cam_focal"], cam["cam_f_number"], cam["cam_exposure"], cam["cam_gain"], wl["fallrate"], 1.0, per)
        t = dict(bgr=torch.from_numpy(bgr).to(dev), d=torch.from_numpy(d16.view(np.int16)).to(dev), r=torch.from_numpy(recs.view(np.uint8).reshape(-1)).to(dev),
                 o=torch.empty((per, H, W, 3), dtype=torch.float32, device=dev), m=torch.empty((per, H, W), dtype=torch.float32, device=dev),
                 u=torch.empty((per, H, W, 3), dtype=torch.uint8, device=dev), i=torch.empty((per, H, W), dtype=torch.uint8, device=dev))
        offs_c = np.ascontiguousarray(offs)
        io = _lib.FrameIO(t["bgr"].data_ptr(), t["d"].data_ptr(), _lib.DEPTH_U16_256, 0, t["r"].data_ptr(), _lib.ptr(offs_c).value,
                          t["o"].data_ptr(), t["m"].data_ptr(), t["u"].data_ptr(), t["i"].data_ptr(), None, None, None, None, None, None, 0)
        ctxs.append(ctx); ios.append(io); keep.append((t, offs_c))
        _lib.check(ctx.lib.rr_render_frames_device_io(ctx.h, per, C.byref(io), 1), "warm")      # sizes the arena
    lib = ctxs[0].lib

    def step():
        for ctx, io in zip(ctxs, ios):
            _lib.check(lib.rr_render_frames_device_io(ctx.h, per, C.byref(io), 0), "render")
        for ctx in ctxs:
            ctx.synchronize()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("lanes %d x %d frames: %.3f ms per %d frames, %.0f frames/s" % (lanes, per, 1000 * dt / n, lanes * per, lanes * per * n / dt))


if __name__ == "__main__":
    main()
