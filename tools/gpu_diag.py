"""Stage-by-stage GPU-vs-oracle diagnostics (run under gpurun; prints, asserts nothing)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import Scenario, ulp_diff_f32  # noqa


def main():
    W, H, nf, nxml = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (640, 480, 2, 1500))]
    ns = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
    sc = Scenario(W, H, nf, nxml, fallrate=25, noise_scale=ns, noise_std=3.0 if ns else 0.0)
    ctx = sc.context()
    print("env size", ctx.H_env, ctx.W_env, "oracle", sc.tables.src.shape)
    src = ctx.debug_read("env_src")
    print("env_src mismatches:", int((src != sc.tables.src).sum()))
    om = ctx.debug_read("omega")
    print("omega rel err max:", float(np.abs(om / sc.omega - 1).max()), "sum", om.sum(), sc.omega.sum())
    fog = ctx.fog_only(sc.bgr, sc.depth)
    recs, offs = sc.records()
    t0 = time.time()
    out = ctx.render_frames(sc.bgr, sc.depth, recs, offs)
    print("render wall s", time.time() - t0, "timings", ctx.timings(), "launches", ctx.kernel_launches())
    plans = ctx.debug_read("plans", 0, len(recs))
    for i in range(nf):
        t0 = time.time()
        o = sc.oracle_frame(i, "canonical", keep_patches=True)
        onat = sc.oracle_frame(i, "native") if i == 0 else None
        print("frame", i, "oracle s %.1f" % (time.time() - t0), "streaks", o.n_streaks, offs[i + 1] - offs[i], "skipped", len(o.skipped))
        ofog = np.moveaxis(o.fog, -1, 0)
        d = np.abs(fog[i] - ofog)
        print("  fog maxabs", d.max(), "n>1e-12", int((d > 1e-12).sum()), "f32ulp max", int(ulp_diff_f32(fog[i].astype(np.float32), ofog.astype(np.float32)).max()))
        env = ctx.envmap_only(np.moveaxis(o.fog, -1, 0)[None])[0]
        oenv = np.round(o.env * 255).astype(np.uint8)
        print("  env mismatches (given oracle fog):", int((env != oenv).sum()))
        ph = ctx.streak_photometry_only(oenv, recs[offs[i]:offs[i + 1]])
        ref = np.array([[p["fov_xy_avg"][0], p["fov_xy_avg"][1], p["drop_Y"]] for p in o.per_streak])
        if len(ref) == len(ph):
            print("  photometry rel err max", float(np.nanmax(np.abs(ph / ref - 1))))
        else:
            print("  photometry count mismatch", len(ref), len(ph))
        # tex idx / noise
        r = recs[offs[i]:offs[i + 1]]
        print("  tex_idx equal", np.array_equal(r["tex_idx"], np.array(o.tex_idx)), "noise equal", np.array_equal(r["noise_deg"], np.array(o.noise)))
        pl = plans[offs[i]:offs[i + 1]]
        # (plan placement is the tight block of DESIGN.md 6.4, not the reference's padded block: not compared)
        dm = np.abs(out["mask"][i].astype(np.float64) - o.rain_mask)
        print("  mask support equal", np.array_equal(out["mask"][i] > 0, o.rain_mask > 0), "mask f32 ulp max",
              int(ulp_diff_f32(out["mask"][i], o.rain_mask.astype(np.float32)).max()), "maxabs", dm.max())
        u = ulp_diff_f32(out["bgr"][i], o.out_bgr.astype(np.float32))
        print("  bgr f32 ulp max", int(u.max()), "hist", np.bincount(np.minimum(u.ravel(), 5)).tolist(), "maxabs",
              np.abs(out["bgr"][i].astype(np.float64) - o.out_bgr).max())
        du = np.abs(out["u8"][i].astype(int) - o.out_u8.astype(int))
        print("  u8 max diff", int(du.max()), "n diff", int((du > 0).sum()))
        if onat is not None:
            u2 = ulp_diff_f32(out["bgr"][i], onat.out_bgr.astype(np.float32))
            print("  vs NATIVE oracle: bgr f32 ulp max", int(u2.max()), "maxabs", np.abs(out["bgr"][i].astype(np.float64) - onat.out_bgr).max(),
                  "u8 max", int(np.abs(out["u8"][i].astype(int) - onat.out_u8.astype(int)).max()))


if __name__ == "__main__":
    main()
