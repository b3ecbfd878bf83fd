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


def _write_png_with_row_filters(path, arr, filters):
    """A PNG whose scanline y uses filter type filters[y] (0 None, 1 Sub, 2 Up, 3 Average, 4 Paeth): what libpng's adaptive
    filtering produces and OpenCV's writer (Sub only) never does."""
    import struct
    import zlib
    H, W = arr.shape[:2]
    ch = 1 if arr.ndim == 2 else arr.shape[2]
    depth = arr.dtype.itemsize * 8
    raw = arr.astype(">u2").tobytes() if depth == 16 else arr.tobytes()
    bpp, stride = ch * depth // 8, W * ch * depth // 8
    rows = [np.frombuffer(raw[y * stride:(y + 1) * stride], np.uint8).astype(np.int32) for y in range(H)]
    out = bytearray()
    zero = np.zeros(stride, np.int32)
    for y in range(H):
        cur, up = rows[y], rows[y - 1] if y else zero
        left = np.concatenate([zero[:bpp], cur[:-bpp]])
        upleft = np.concatenate([zero[:bpp], up[:-bpp]])
        ft = int(filters[y])
        if ft == 0:
            pred = zero
        elif ft == 1:
            pred = left
        elif ft == 2:
            pred = up
        elif ft == 3:
            pred = (left + up) >> 1
        else:
            pa, pb, pc = np.abs(up - upleft), np.abs(left - upleft), np.abs(left + up - 2 * upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, upleft))
        out.append(ft)
        out += ((cur - pred) & 255).astype(np.uint8).tobytes()
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    color = {1: 0, 2: 4, 3: 2, 4: 6}[ch]
    z = zlib.compress(bytes(out), 6)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, depth, color, 0, 0, 0)) +
                chunk(b"IDAT", z[:len(z) // 2]) + chunk(b"IDAT", z[len(z) // 2:]) + chunk(b"IEND", b""))


def test_decode_with_every_row_filter_and_their_mixtures(tmp_path):
    """Scanlines filtered with None / Sub / Up / Average / Paeth in every order (the decoder leaves Sub and None lines
    filtered when it can hand them to the batch buffer in one pass -- only if the NEXT line does not need them), two IDAT
    chunks, all layouts: equal to cv2.imread."""
    rng = np.random.RandomState(9)
    H, W = 41, 37
    smooth = lambda shape, dt: (np.cumsum(rng.randint(-3, 4, shape), axis=1) % (256 if dt == np.uint8 else 65536)).astype(dt)
    layouts = {"rgb8": smooth((H, W, 3), np.uint8), "gray16": smooth((H, W), np.uint16), "gray8": smooth((H, W), np.uint8),
               "rgba8": smooth((H, W, 4), np.uint8), "rgb16": smooth((H, W, 3), np.uint16)}
    patterns = [np.full(H, t) for t in range(5)] + [rng.randint(0, 5, H) for _ in range(6)] + [np.arange(H) % 2, 1 + np.arange(H) % 2,
                np.where(np.arange(H) % 3 == 0, 4, 1), np.where(np.arange(H) == H - 1, 2, 1), np.where(np.arange(H) == 0, 1, 3)]
    for name, arr in layouts.items():
        for k, filt in enumerate(patterns):
            p = str(tmp_path / ("%s_%d.png" % (name, k)))
            _write_png_with_row_filters(p, arr, filt)
            ref = cv2.imread(p)
            assert ref is not None
            out = np.zeros((1, H, W, 3), np.uint8)
            assert pngio.read_batch([p], None, out, None, 1).tolist() == [0], (name, k)
            assert np.array_equal(out[0], ref), (name, filt.tolist())
            if arr.ndim == 2:
                unch = cv2.imread(p, cv2.IMREAD_UNCHANGED)
                d16 = np.zeros((1, H, W), np.uint16)
                assert pngio.read_batch(None, [p], None, d16, 1).tolist() == [0]
                assert np.array_equal(d16[0], unch.astype(np.uint16)), (name, filt.tolist())
                d32 = np.zeros((1, H, W), np.float32)
                assert pngio.read_batch(None, [p], None, d32, 1).tolist() == [0]
                assert np.array_equal(d32[0], unch.astype(np.float32) / 256.)


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


def test_deflate_encoder_streams_are_valid_zlib_for_edge_cases():
    """csrc/rr_host_deflate.h: run + dynamic-Huffman blocks.  Python's zlib must reproduce the input (it also verifies the
    Adler-32): empty and tiny inputs, runs around the 258-byte match limit, blocks longer than 65536 tokens, a symbol
    distribution whose optimal code is deeper than 15 bits (length-limiting repair), incompressible noise."""
    import zlib
    rng = np.random.RandomState(0)
    fib = [1, 1]
    while len(fib) < 30:
        fib.append(fib[-1] + fib[-2])
    skew = np.concatenate([np.full(min(f, 200000), i, np.uint8) for i, f in enumerate(fib)])
    rng.shuffle(skew)
    cases = [b"", b"a", b"ab", b"aaa", b"aaaa", b"a" * 258, b"a" * 259, b"a" * 260, b"ab" + b"c" * 261 + b"d", bytes(100000),
             b"abc" * 1000, rng.randint(0, 256, 70000).astype(np.uint8).tobytes(), rng.randint(0, 2, 200000).astype(np.uint8).tobytes(),
             (rng.rand(300000) < 0.01).astype(np.uint8).tobytes(), bytes(range(256)) * 300,
             np.repeat(rng.randint(0, 256, 5000).astype(np.uint8), rng.randint(1, 600, 5000)).tobytes(), skew.tobytes()]
    for c in cases:
        z = pngio.zlib_compress_fast(c)
        assert zlib.decompress(z) == c, len(c)
    assert len(pngio.zlib_compress_fast(bytes(100000))) < 200                  # runs really are matches
    noise = cases[11]
    assert len(pngio.zlib_compress_fast(noise)) < len(noise) + 400              # and noise costs only the block headers


def test_inflate_decoder_against_zlib_streams_of_every_block_type_and_mutations():
    """csrc/rr_host_inflate.h: what zlib produces at every level / strategy (stored, fixed and dynamic Huffman blocks, long
    matches, codes longer than the 12-bit first-level table) must decode to the input; truncated, corrupted or
    wrong-size streams must be refused, never crash or overrun."""
    import zlib
    rng = np.random.RandomState(1)
    fib = [1, 1]
    while len(fib) < 24:
        fib.append(fib[-1] + fib[-2])
    skew = np.concatenate([np.full(f, i, np.uint8) for i, f in enumerate(fib)])
    rng.shuffle(skew)
    smooth = np.cumsum(rng.randint(-2, 3, 200000)).astype(np.uint8).tobytes()
    datas = [b"", b"a", b"hello hello hello hello", bytes(70000), rng.randint(0, 256, 100000).astype(np.uint8).tobytes(), smooth, skew.tobytes(),
             (b"0123456789abcdef" * 5000) + rng.randint(0, 256, 3000).astype(np.uint8).tobytes() + bytes(range(256)) * 40,
             np.repeat(rng.randint(0, 256, 3000).astype(np.uint8), rng.randint(1, 400, 3000)).tobytes(),
             # long matches with periods 2 .. 7 (copied in words from a multiple of the period back)
             b"".join(bytes(rng.randint(0, 256, per).astype(np.uint8)) * int(rng.randint(3, 200)) for per in list(range(2, 8)) * 40)]
    n_streams = 0
    for d in datas:
        for level in (0, 1, 6, 9):
            for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
                c = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strat)
                z = c.compress(d[:len(d) // 2]) + c.flush(zlib.Z_FULL_FLUSH) + c.compress(d[len(d) // 2:]) + c.flush()
                assert pngio.zlib_decompress_fast(z, len(d)) == d, (len(d), level, strat)
                n_streams += 1
                if len(d) > 100:
                    assert pngio.zlib_decompress_fast(z, len(d) - 1) is None and pngio.zlib_decompress_fast(z, len(d) + 1) is None
                    assert pngio.zlib_decompress_fast(z[:len(z) // 2], len(d)) is None
        z = pngio.zlib_compress_fast(d)                                      # and the library's own encoder
        assert pngio.zlib_decompress_fast(z, len(d)) == d
    assert n_streams == len(datas) * 20
    # mutations: whatever comes back is either None or exactly what zlib itself would accept
    z0 = zlib.compress(smooth, 6)
    accepted = 0
    for it in range(1500):
        b = bytearray(z0)
        for _ in range(rng.randint(1, 4)):
            b[rng.randint(2, len(b))] = rng.randint(0, 256)
        got = pngio.zlib_decompress_fast(bytes(b), len(smooth))
        if got is not None:
            accepted += 1
            assert zlib.decompress(bytes(b)) == got
    assert accepted < 20                                                     # the Adler-32 catches nearly everything that still parses


def test_inflate_literal_pairs_tail_path_and_adler_blocks():
    """Literal-pair entries of the first-level table (two short codes in one look-up): literal-heavy data with short codes,
    every output length around the 320-byte head-room where the decoder switches to the checked path, lengths around the
    5552-byte Adler-32 blocks, and a wrong trailer."""
    import zlib
    rng = np.random.RandomState(7)
    # skewed byte distribution: Huffman codes of 2 .. 12 bits, no matches worth taking at level 1 / HUFFMAN_ONLY
    p = np.array([0.30, 0.20, 0.12, 0.10, 0.08] + [0.20 / 251] * 251)
    src = rng.choice(256, size=70000, p=p / p.sum()).astype(np.uint8).tobytes()
    lengths = list(range(0, 12)) + list(range(300, 345)) + [640, 641, 5551, 5552, 5553, 11103, 11104, 11105, 70000]
    for n in lengths:
        d = src[:n]
        for strat in (zlib.Z_HUFFMAN_ONLY, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED):
            c = zlib.compressobj(6, zlib.DEFLATED, 15, 8, strat)
            z = c.compress(d) + c.flush()
            assert pngio.zlib_decompress_fast(z, n) == d, (n, strat)
            if n:
                assert pngio.zlib_decompress_fast(z, n - 1) is None, (n, strat)
                bad = bytearray(z); bad[-1] ^= 1                         # Adler-32 trailer off by one bit
                assert pngio.zlib_decompress_fast(bytes(bad), n) is None
    # all 0xff: the largest sums the Adler-32 blocks have to hold
    d = b"\xff" * 200001
    assert pngio.zlib_decompress_fast(zlib.compress(d, 1), len(d)) == d


def test_reference_format_files_rgba_image_and_viridis_mask(tmp_path):
    """plt.imsave's formats (common/generator.py:466-467): RGBA with alpha 255; the mask through the colormap."""
    H, W, n = 45, 83, 4
    bgr = np.stack([_rand((H, W, 3), np.uint8, 150 + i) for i in range(n)])
    idx = np.stack([_rand((H, W), np.uint8, 170 + i) for i in range(n)])
    idx[2] = 0                                                                  # a frame without rain: one colour
    idx[3, :, :40] = 255
    lut = pngio.viridis_rgb()
    assert lut[0].tolist() == [68, 1, 84] and lut[255].tolist() == [253, 231, 36] and lut[128].tolist() == [32, 144, 140]
    for level in (0, 1, 6):
        ip = [str(tmp_path / ("r%d_%d.png" % (level, i))) for i in range(n)]
        mp = [str(tmp_path / ("m%d_%d.png" % (level, i))) for i in range(n)]
        assert pngio.write_batch_rgba(ip, bgr, mp, idx, level, 3) == 0
        for i in range(n):
            a = cv2.imread(ip[i], cv2.IMREAD_UNCHANGED)
            assert a.shape == (H, W, 4) and np.array_equal(a[..., :3], bgr[i]) and (a[..., 3] == 255).all()
            m = cv2.imread(mp[i], cv2.IMREAD_UNCHANGED)
            assert m.shape == (H, W, 4) and np.array_equal(m[..., 2::-1], lut[idx[i]]) and (m[..., 3] == 255).all()
    # the compact pair: RGB + 16-bit gray
    u16 = np.stack([_rand((H, W), np.uint16, 190 + i) for i in range(n)])
    ip = [str(tmp_path / ("c%d.png" % i)) for i in range(n)]
    mp = [str(tmp_path / ("cm%d.png" % i)) for i in range(n)]
    assert pngio.write_batch_u16(ip, bgr, mp, u16, 1, 2) == 0
    for i in range(n):
        assert np.array_equal(cv2.imread(ip[i]), bgr[i]) and np.array_equal(cv2.imread(mp[i], cv2.IMREAD_UNCHANGED), u16[i])


def test_depth_as_uint16_samples_and_header_bombs(tmp_path):
    H, W = 33, 47
    d16 = _rand((H, W), np.uint16, 7); d8 = _rand((H, W), np.uint8, 8)
    p16, p8 = str(tmp_path / "d16.png"), str(tmp_path / "d8.png")
    cv2.imwrite(p16, d16); cv2.imwrite(p8, d8)
    out = np.zeros((2, H, W), np.uint16)
    assert pngio.read_batch(None, [p16, p8], None, out, 2).tolist() == [0, 0]
    assert np.array_equal(out[0], d16) and np.array_equal(out[1], d8.astype(np.uint16))
    # a tiny file whose header claims 65536 x 65536 RGBA16 (34 GB): refused on the size before anything is allocated
    import struct
    import zlib
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)
    bomb = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 65536, 65536, 16, 6, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b"")
    pb = str(tmp_path / "bomb.png")
    open(pb, "wb").write(bomb)
    img = np.zeros((1, H, W, 3), np.uint8)
    assert pngio.read_batch([pb], None, img, None, 1).tolist() == [pngio.SIZE]
    assert pngio.read_batch([pb], None, img, None, 4).tolist() == [pngio.SIZE]
    big = np.zeros((1, 8, 8, 3), np.uint8)                                     # even when the caller's size matches the lie
    assert pngio.info(pb)[:2] == (65536, 65536)


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
