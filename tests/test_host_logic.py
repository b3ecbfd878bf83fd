"""CPU checks of the product's host side: the C-ABI library loads and exports every symbol the
header declares (no compute calls without a GPU), baked tables, the NumPy RNG mirror, the
particles-XML loader and the per-frame record assembly (against the oracle's own parser)."""
import ctypes
import os
import re

import cv2
import numpy as np
import pytest

from rain_rendering_b200 import _lib, api, streaks as S
from oracle import rain_oracle as ro
from util import Scenario

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "rain_b200.h")).read()
    names = set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    for n in sorted(names):
        assert hasattr(lib, n), "librain_b200.so does not export %s" % n
    assert lib.rr_version() >= 100


def test_struct_sizes_match_header():
    assert S.STREAK_DTYPE.itemsize == 128
    assert ctypes.sizeof(_lib.Camera) == 8 + 10 * 8 + 8      # W, H, ten doubles, render_scale + reserved
    assert _lib.PLAN_DTYPE.itemsize == 280


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.RainError, match="no usable CUDA device|CUDA"):
        api.RainContext(0)


def test_baked_gaussian_tables_equal_opencv():
    lib = _lib.load()
    k64, k32, k15 = np.zeros(25), np.zeros(25, np.float32), np.zeros(15, np.int32)
    lib.rr_host_tables(_lib.ptr(k64), _lib.ptr(k32), _lib.ptr(k15))
    assert np.array_equal(k64, cv2.getGaussianKernel(25, 25, cv2.CV_64F).ravel())
    assert np.array_equal(k32, cv2.getGaussianKernel(25, 25, cv2.CV_32F).ravel())
    assert k15.sum() == 256
    # the fixed-point kernel reproduces cv2.GaussianBlur((15,15), 0) on uint8 exactly
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (40, 50, 3)).astype(np.uint8)

    def r101(n, r):
        i = np.arange(-r, n + r)
        i = np.where(i < 0, -i, i)
        return np.where(i >= n, 2 * (n - 1) - i, i)

    p = img[:, r101(50, 7)].astype(np.int64)
    h = sum(int(k15[t]) * p[:, t:t + 50] for t in range(15))
    q = h[r101(40, 7)]
    v = sum(int(k15[t]) * q[t:t + 40] for t in range(15))
    assert np.array_equal(((v + 32768) >> 16).astype(np.uint8), cv2.GaussianBlur(img, (15, 15), 0))


@pytest.mark.parametrize("seed", [0, 3, 2 ** 31 + 5])
@pytest.mark.parametrize("std,scale", [(0.0, 0.0), (4.0, 1.5)])
def test_rng_mirror_equals_numpy_legacy_stream(seed, std, scale):
    rs = np.random.RandomState(1)
    n = 700
    types = rs.randint(0, 3, n).astype(np.uint8)
    buckets = rs.randint(0, 5, n).astype(np.int32)
    np.random.seed(seed)
    tex, noise = [], []
    for i in range(n):
        tex.append(np.random.randint(10 * buckets[i], 10 * buckets[i] + 10))
        noise.append(np.random.normal(0.0, std) * scale if types[i] != 0 else 0.0)
    t2, n2 = api.draw_randoms(seed, types, buckets, std, scale)
    assert np.array_equal(t2, np.array(tex))
    assert np.array_equal(n2, np.array(noise))


def test_xml_loader_and_record_assembly_match_the_oracle_parser():
    sc = Scenario(320, 200, 3, 2500, noise_scale=2.0, noise_std=3.0, seed=4, n_sim_frames=2)
    for of, pf in zip(sc.oracle_frames, sc.sim_frames):
        assert len(of) == len(pf) > 50
        for a, b in zip(of, pf):
            assert a.pid == b["pid"] and a.max_width == b["max_width"] and a.length == b["length"] and a.drop_type == b["type"]
            assert np.array_equal(a.ip1, b["ip1"]) and np.array_equal(a.ip2, b["ip2"])
            assert np.array_equal(a.wp1, b["wp1"]) and np.array_equal(a.wp2, b["wp2"]) and a.iw1 == b["iw1"] and a.iw2 == b["iw2"]
            assert np.isclose(a.ratio, b["ratio"], rtol=1e-13, equal_nan=True)
    recs, offs = sc.records()
    # replay the oracle's per-frame driver bookkeeping (filter, RNG, wind write-back) without rendering
    for i in range(sc.n_frames):
        np.random.seed(i)
        todo = ro.filter_in_frame(sc.oracle_frames[i % 2], sc.W, sc.H)
        r = recs[offs[i]:offs[i + 1]]
        assert len(todo) == len(r) > 10
        for s, q in zip(todo, r):
            ti = np.random.randint(10 * ro.texture_bucket(s.ratio, sc.db.ratios), 10 * ro.texture_bucket(s.ratio, sc.db.ratios) + 10)
            noise = 0.0 if s.drop_type == ro.BIG else np.random.normal(0.0, sc.cam.noise_std) * sc.cam.noise_scale
            assert ti == q["tex_idx"] and noise == q["noise_deg"] and s.pid == q["pid"]
            assert np.array_equal(s.ip1, q["ip1"]) and np.array_equal(s.ip2, q["ip2"])
            if s.drop_type != ro.BIG:
                ro.make_patch(s, sc.db.textures[ti], sc.cam, noise)      # mutates s.ip1/ip2 like the reference
            assert np.array_equal(s.ip1, q["ip1m"]) and np.array_equal(s.ip2, q["ip2m"])


def test_public_header_is_plain_c(tmp_path):
    """include/rain_b200.h is the drop-in boundary: it must compile as C (no C++ types in the signatures) and a C
    caller must be able to link the host-side entry points."""
    import subprocess
    from rain_rendering_b200 import build as B
    src = tmp_path / "use.c"
    src.write_text(
        '#include "rain_b200.h"\n'
        '#include <stdio.h>\n'
        'int main(void) {\n'
        '    rr_streak_rec r; rr_camera c; rr_sim_params p; rr_sim_streak s; rr_xml_frame f;\n'
        '    double k64[25]; float k32[25]; int k15[15];\n'
        '    (void)r; (void)c; (void)p; (void)s; (void)f;\n'
        '    rr_host_tables(k64, k32, k15);\n'
        '    printf("%d %d %d %d %.6f\\n", rr_version(), (int)sizeof(rr_streak_rec), (int)sizeof(rr_sim_streak), (int)sizeof(rr_xml_frame), k64[12]);\n'
        '    return 0;\n'
        '}\n')
    exe = str(tmp_path / "use")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe,
                           B.LIB, "-Wl,-rpath," + os.path.dirname(B.LIB)])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    ver, rec, sim, xf, k = out.stdout.split()
    assert (int(rec), int(sim), int(xf)) == (128, 120, 32) and int(ver) >= 100 and abs(float(k) - 0.041670) < 1e-5


def test_integration_md_camera_stub_matches_the_header():
    """The ctypes stub a maintainer would copy out of INTEGRATION.md must describe rr_camera exactly (a stub eight bytes short
    makes rr_set_camera read past the caller's struct)."""
    import ctypes as C
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    m = re.search(r"class Camera\(C\.Structure\):.*?\n(?=assert C\.sizeof)", md, re.S)
    assert m, "INTEGRATION.md no longer shows the Camera stub"
    ns = {"C": C}
    exec(m.group(0), ns)
    stub = ns["Camera"]
    assert C.sizeof(stub) == C.sizeof(_lib.Camera) == 96
    assert [(n, t) for n, t in stub._fields_] == [(n, t) for n, t in _lib.Camera._fields_]
    # and the header says the same: 2 + 2 int32 around ten doubles
    hdr = open(os.path.join(root, "include", "rain_b200.h")).read()
    body = hdr[hdr.index("typedef struct rr_camera {"):hdr.index("} rr_camera;")]
    assert len(re.findall(r"\bdouble\s+\w+;", body)) == 10 and "int32_t W, H;" in body and "int32_t render_scale;" in body and "int32_t reserved;" in body


def test_lanes_queue_is_fifo_over_the_contexts_round_robin():
    """api.RainLanes without a GPU (fake contexts): submissions go to the lanes in turn, wait_frames retires the oldest,
    the queue holds two per lane, and configuration calls reach every lane."""
    log = []

    class Fake:
        def __init__(self, device):
            self.k = len(log_ctx); log_ctx.append(self); self.W = self.H = self.max_batch = 0; self.db_ratios = None; self.n = 0
        def set_streak_db(self, t, r=None): self.db_ratios = r
        def set_camera(self, W, H, *a, **k): self.W, self.H, self.max_batch = W, H, 4
        def submit_frames(self, tag): log.append(("submit", self.k, tag)); self.n += 1
        def wait_frames(self): log.append(("wait", self.k)); self.n -= 1
        def synchronize(self): log.append(("sync", self.k))
        def kernel_launches(self): return 10 + self.k
        def close(self): pass

    log_ctx = []
    lanes = api.RainLanes(0, 3, context_factory=Fake)
    lanes.set_streak_db([], [1.0]); lanes.set_camera(64, 32)
    assert (lanes.W, lanes.H, lanes.max_batch) == (64, 32, 4) and all(c.db_ratios == [1.0] for c in log_ctx)
    assert lanes.capacity == 6 and lanes.kernel_launches() == 33
    for t in range(6):
        lanes.submit_frames(t)
    assert [e[1] for e in log] == [0, 1, 2, 0, 1, 2]
    with pytest.raises(_lib.RainError):
        lanes.submit_frames(6)
    lanes.wait_frames(); lanes.wait_frames()
    assert log[-2:] == [("wait", 0), ("wait", 1)] and lanes.inflight == 4
    lanes.submit_frames(6)                       # the turn continues where it stopped: lane 0
    assert log[-1] == ("submit", 0, 6)
    lanes.synchronize()
    assert lanes.inflight == 0 and all(c.n == 0 for c in log_ctx)
    with pytest.raises(_lib.RainError):
        lanes.wait_frames()
