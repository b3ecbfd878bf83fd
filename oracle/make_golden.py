"""Generate the golden fixtures under tests/golden/ by running the UNTOUCHED reference
(/root/reference) through oracle/ref_harness.py on seeded synthetic inputs.

    python -m oracle.make_golden            # from the repo root, in the build container

The reference cannot travel to the GPU box, so its outputs are committed here together with the
exact inputs.  Two fixtures:
  small_256x192.npz  2 frames, wind noise on, full float64 outputs + stage intermediates
  c1_640x480.npz     BASELINE config C1 (640x480, 10 mm/h, pre-computed XML): outputs as
                     float32 + SHA-256 of the float64 arrays (bit-exact pin of the oracle)
  c2_1242x375.npz    BASELINE config C2's frame (1242x375, 25 mm/h): SHA-256 of the float64 outputs +
                     the uint8 image the reference would save
"""
from __future__ import annotations

import hashlib
import os
import platform
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402
from rain_rendering_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def host_signature() -> str:
    import cv2
    ids = dict(SSE4_1=6, AVX=10, AVX2=11, FMA3=12, AVX512_SKX=256)   # cv::CpuFeatures
    feats = [f for f, i in ids.items() if cv2.checkHardwareSupport(i)]
    return "numpy %s; cv2 %s [%s]; %s" % (np.__version__, cv2.__version__, ",".join(feats), platform.machine())


def run(name, W, H, n_frames, fallrate, n_xml, seed, noise_scale, noise_std, opacity, n_sim_frames, full64, compact=False):
    import cv2
    root = tempfile.mkdtemp(prefix="rr_golden_")
    try:
        paths = synth.write_dataset(root, "customdb", "seq1", W, H, n_frames, fallrate, n_xml, seed=seed,
                                    n_sim_frames=n_sim_frames)
        ref = ref_harness.run_reference(paths, "customdb", fallrate, noise_scale=noise_scale, noise_std=noise_std,
                                        opacity_attenuation=opacity)
        names = sorted(ref)
        src = os.path.join(root, "source", "customdb", "seq1")
        bgr = np.stack([cv2.imread(os.path.join(src, "rgb", n + ".png")) for n in names])
        depth_u16 = np.stack([cv2.imread(os.path.join(src, "depth", n + ".png"), cv2.IMREAD_UNCHANGED) for n in names])
        xml = open(paths["xml"]).read()
        out = dict(W=W, H=H, n_frames=n_frames, fallrate=fallrate, seed=seed, noise_scale=noise_scale, noise_std=noise_std,
                   opacity=opacity, bgr=bgr, depth_u16=depth_u16, xml=np.frombuffer(xml.encode(), np.uint8),
                   db_seed=seed, host=host_signature(), n_streaks=np.array([len(ref[n]["streaks"]) for n in names]))
        rainy = np.stack([ref[n]["rainy_rgb"] for n in names])      # RGB, clipped, float64 (what imsave receives)
        mask = np.stack([ref[n]["rain_mask"] for n in names])
        out["rainy_sha"] = sha(rainy)
        out["mask_sha"] = sha(mask)
        if full64:
            out["rainy_rgb"] = rainy
            out["rain_mask"] = mask
            out["fog0"] = ref[names[0]]["fog"]
            out["env0_u8"] = np.round(ref[names[0]]["env"] * 255).astype(np.uint8)
        elif compact:      # hashes of the float64 arrays + what plt.imsave would quantise to (small fixture for a large frame)
            out["rainy_u8"] = (rainy * 255).astype(np.uint8)
        else:
            out["rainy_rgb_f32"] = rainy.astype(np.float32)
            out["rain_mask_f32"] = mask.astype(np.float32)
        os.makedirs(GOLD, exist_ok=True)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(name, "frames", n_frames, "streaks", out["n_streaks"].tolist(), "bytes", os.path.getsize(os.path.join(GOLD, name + ".npz")))
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    assert ref_harness.reference_available(), "needs /root/reference (run in the build container)"
    only = sys.argv[1:]
    if not only or "c2_1242x375" in only:
        # BASELINE config C2's frame shape and rain rate (KITTI optics, odd height): hashes + the quantised image
        run("c2_1242x375", 1242, 375, 1, 25, 1000, seed=4, noise_scale=0.0, noise_std=0.0, opacity=1.0, n_sim_frames=1, full64=False, compact=True)
    if only:
        sys.exit(0)
    run("small_256x192", 256, 192, 3, 25, 600, seed=1, noise_scale=1.5, noise_std=3.0, opacity=0.8, n_sim_frames=2, full64=True)
    run("c1_640x480", 640, 480, 1, 10, 420, seed=2, noise_scale=0.0, noise_std=0.0, opacity=1.0, n_sim_frames=1, full64=False)
