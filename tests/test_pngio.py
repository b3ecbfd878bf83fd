"""The native PNG codec (csrc/rr_host_png.cpp) against OpenCV, which is what the reference decodes with
(common/generator.py:352,360) and what the drop-in used to encode with.  Host logic only."""
import os
import time

import cv2
import numpy as np
import pytest

from rain_rendering_b200 import pngio, synth


def _rand(shape, dtype, seed):
    rng = np.random.RandomState(seed)
    info = np.iinfo(dtype)
    smooth = cv2.resize(rng.randint(0, info.max, (8, 8) + shape[2:]).astype(np.float32), (shape[1], shape[0]), interpolation=cv2.INTER_CUBIC)
    return np.clip(smooth + rng.randint(-3, 4, shape), 0, info.max).astype(dtype)


def test_decode_equals_cv2_imread_for_every_supported_layout(tmp_path):
    H, W = 37, 53
    files = {
        "rgb8.png": _rand((H, W, 3), np.uint8, 1), "gray8.png": _rand((H, W), np.uint8, 2), "rgba8.png": _rand((H, W, 4), np.uint8, 3),
        "rgb16.png": _rand((H, W, 3), np.uint16, 4), "gray16.png": _rand((H, W), np.uint16, 5),
    }
    for name, arr in files.items():
        for level in (0, 3, 9):                      # libpng picks different row filters per level / content
            p = str(tmp_path / ("L%d_%s" % (level, name)))
            assert cv2.imwrite(p, arr, [cv2.IMWRITE_PNG_COMPRESSION, level])
            w, h, ch, depth = pngio.info(p)
            assert (w, h) == (W, H) and depth == arr.dtype.itemsize * 8 and ch == (1 if arr.ndim == 2 else arr.shape[2])
            out = np.zeros((1, H, W, 3), np.uint8)
            st = pngio.read_batch([p], None, out, None, 2)
            assert st.tolist() == [0], (name, st)
            assert np.array_equal(out[0], cv2.imread(p)), name
            if arr.ndim == 2:
                d = np.zeros((1, H, W), np.float32)
                assert pngio.read_batch(None, [p], None, d, 1).tolist() == [0]
                assert np.array_equal(d[0], cv2.imread(p, cv2.IMREAD_UNCHANGED).astype(np.float32) / 256.)


def test_batches_wrong_sizes_and_unsupported_files(tmp_path):
    H, W, n = 40, 64, 7
    imgs = [_rand((H, W, 3), np.uint8, 10 + i) for i in range(n)]
    deps = [_rand((H, W), np.uint16, 30 + i) for i in range(n)]
    ip, dp = [], []
    for i in range(n):
        ip.append(str(tmp_path / ("i%d.png" % i))); dp.append(str(tmp_path / ("d%d.png" % i)))
        cv2.imwrite(ip[-1], imgs[i]); cv2.imwrite(dp[-1], deps[i])
    cv2.imwrite(ip[2], _rand((H + 1, W, 3), np.uint8, 99))                       # wrong size
    from PIL import Image
    Image.fromarray(imgs[4][..., ::-1]).convert("P").save(ip[4])                 # palette PNG: not decoded natively
    Image.fromarray(imgs[5][..., ::-1]).save(ip[5], interlace=True) if False else None
    with open(dp[6], "wb") as f:
        f.write(b"not a png")
    bgr = np.zeros((n, H, W, 3), np.uint8); dep = np.zeros((n, H, W), np.float32)
    st = pngio.read_batch(ip, dp, bgr, dep, 4)
    assert st.tolist() == [0, 0, pngio.SIZE, 0, pngio.UNSUPPORTED, 0, -2]
    for i in (0, 1, 3, 5):
        assert np.array_equal(bgr[i], imgs[i]) and np.array_equal(dep[i], deps[i].astype(np.float32) / 256.)
    with pytest.raises(Exception):
        pngio.info(str(tmp_path / "missing.png"))


@pytest.mark.parametrize("level", [0, 1, 6])
def test_encode_round_trips_through_cv2(tmp_path, level):
    H, W, n = 48, 80, 5
    bgr = np.stack([_rand((H, W, 3), np.uint8, 50 + i) for i in range(n)])
    mask = np.stack([(_rand((H, W), np.uint16, 70 + i).astype(np.float32) / 3000.0) for i in range(n)])
    mask[3] = 0.25                                                                # flat mask -> all zeros, like the drop-in's fallback
    ip = [str(tmp_path / "out" / ("r%d.png" % i)) for i in range(n)]
    mp = [str(tmp_path / "out" / ("m%d.png" % i)) for i in range(n)]
    os.makedirs(str(tmp_path / "out"))
    assert pngio.write_batch(ip, bgr, mp, mask, level, 3) == 0
    for i in range(n):
        assert np.array_equal(cv2.imread(ip[i]), bgr[i])                          # cv2.imwrite(path, bgr) would read back the same
        got = cv2.imread(mp[i], cv2.IMREAD_UNCHANGED)
        lo, hi = float(mask[i].min()), float(mask[i].max())
        norm = (mask[i] - lo) / (hi - lo) if hi > lo else np.zeros_like(mask[i])
        want = (norm.astype(np.float64) * 65535.0 + 0.5).astype(np.uint16)
        assert got.dtype == np.uint16 and np.abs(got.astype(int) - want.astype(int)).max() <= 1
        assert not os.path.exists(ip[i] + ".part")
    assert pngio.write_batch([str(tmp_path / "no_such_dir" / "x.png")], bgr[:1], None, None, level, 1) == 1


def test_codec_speed_on_a_kitti_sized_frame(tmp_path):
    bgr, depth = synth.make_frame(1242, 375, 3)
    p, q = str(tmp_path / "a.png"), str(tmp_path / "d.png")
    cv2.imwrite(p, bgr); cv2.imwrite(q, np.round(depth * 256).astype(np.uint16))
    out, dep = np.zeros((1, 375, 1242, 3), np.uint8), np.zeros((1, 375, 1242), np.float32)
    t0 = time.perf_counter(); [pngio.read_batch([p], [q], out, dep, 1) for _ in range(3)]; t_rd = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter(); [(cv2.imread(p), cv2.imread(q, cv2.IMREAD_UNCHANGED)) for _ in range(3)]; t_cv = (time.perf_counter() - t0) / 3
    m = np.random.RandomState(0).rand(1, 375, 1242).astype(np.float32)
    t0 = time.perf_counter(); [pngio.write_batch([p + "w.png"], out, [p + "m.png"], m, 1, 1) for _ in range(3)]; t_wr = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter(); [(cv2.imwrite(p + "c.png", out[0]), cv2.imwrite(p + "cm.png", (m[0] * 65535).astype(np.uint16))) for _ in range(3)]; t_cw = (time.perf_counter() - t0) / 3
    print("decode native %.1f ms vs cv2 %.1f ms; encode native %.1f ms vs cv2 %.1f ms (one thread, image + depth/mask)" % (t_rd * 1e3, t_cv * 1e3, t_wr * 1e3, t_cw * 1e3))
    assert np.array_equal(out[0], bgr) and t_wr < 1.2 * t_cw and t_rd < 1.5 * t_cv


def test_corrupted_pixel_data_is_rejected_not_decoded(tmp_path):
    """A flipped byte inside the compressed stream must fail zlib's Adler-32 (or the stream structure), never yield a
    silently different image: the frame then takes the OpenCV fallback, which reports the file the way the reference
    sees it."""
    arr = _rand((40, 64, 3), np.uint8, 3)
    p = str(tmp_path / "a.png")
    cv2.imwrite(p, arr)
    raw = bytearray(open(p, "rb").read())
    i = raw.index(b"IDAT") + 4 + 40
    raw[i] ^= 0x55
    q = str(tmp_path / "bad.png")
    open(q, "wb").write(bytes(raw))
    out = np.zeros((2, 40, 64, 3), np.uint8)
    st = pngio.read_batch([p, q], None, out, None, 2)
    assert st[0] == 0 and st[1] != 0 and np.array_equal(out[0], arr)


def test_native_parsers_survive_mutated_files(tmp_path):
    """Truncated / mutated PNG and particles-XML files must be rejected (or read), never crash the process: run in a child."""
    import subprocess
    import sys
    code = r'''
import os, random, sys
import numpy as np, cv2
sys.path.insert(0, %r)
from rain_rendering_b200 import pngio, streaks as S, synth, _lib
random.seed(7); rng = np.random.RandomState(7); d = %r
H, W = 21, 35
base = []
for k, arr in enumerate([rng.randint(0, 256, (H, W, 3)).astype(np.uint8), rng.randint(0, 65536, (H, W)).astype(np.uint16)]):
    p = os.path.join(d, "b%%d.png" %% k); cv2.imwrite(p, arr); base.append(open(p, "rb").read())
out = np.zeros((1, H, W, 3), np.uint8); dep = np.zeros((1, H, W), np.float32)
def mutate(b, alphabet=None):
    b = bytearray(b); m = random.random()
    if m < 0.3: return bytes(b[:random.randrange(len(b))])
    if m < 0.8:
        for _ in range(random.randrange(1, 6)): b[random.randrange(len(b))] = random.choice(alphabet) if alphabet else random.randrange(256)
        return bytes(b)
    i = random.randrange(len(b)); b[i:i] = bytes((random.choice(alphabet) if alphabet else random.randrange(256)) for _ in range(random.randrange(1, 30)))
    return bytes(b)
for it in range(400):
    p = os.path.join(d, "f.png"); open(p, "wb").write(mutate(random.choice(base)))
    pngio.read_batch([p], [p], out, dep, 2)
    try: pngio.info(p)
    except Exception: pass
parts = synth.make_particles(320, 240, 2, 40, 2.0, seed=3)
x = os.path.join(d, "s_camera0.xml"); synth.write_particles_xml(parts, x, 2.0); xb = open(x, "rb").read()
for it in range(400):
    p = os.path.join(d, "f.xml"); open(p, "wb").write(mutate(xb, b"<>/\"'=[]; \n0123456789.e-x&"))
    try: S.load_streaks_from_xml(p, 1, 320, 240)
    except _lib.RainError: pass
print("SURVIVED")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "SURVIVED" in out.stdout, (out.returncode, out.stderr[-1500:])
