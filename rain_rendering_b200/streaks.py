"""Host-side streak records: the particles-XML loader, the in-frame filter and the per-frame
record assembly that feeds ``rr_render_frames`` (include/rain_b200.h: ``rr_streak_rec``).

Mirrors, on the host, the parts of the reference that are bookkeeping rather than rendering:
  * DBManager.load_streaks_from_xml      common/bad_weather.py:148-248
  * DBManager.classify_drop              common/bad_weather.py:99-106
  * the in-frame streak filter           common/generator.py:413-420
  * DBManager.take_drop_texture buckets  common/bad_weather.py:250-265
  * the wind-noise write-back into the shared Streak objects   common/generator.py:149-161
The two NumPy RNG draws per streak are reproduced by the library's own MT19937 mirror
(``rr_host_draw_randoms`` in csrc/rr_host.cpp) so that no Python-level RNG call sits on the
per-frame path.
"""
from __future__ import annotations

import os
from xml.etree.ElementTree import parse

import numpy as np

BIG, MEDIUM, SMALL = 0, 1, 2

STREAK_DTYPE = np.dtype([
    ("wp1", "<f8", 3), ("wp2", "<f8", 3), ("iw1", "<f8"), ("iw2", "<f8"), ("noise_deg", "<f8"), ("ratio", "<f8"),
    ("ip1", "<i4", 2), ("ip2", "<i4", 2), ("ip1m", "<i4", 2), ("ip2m", "<i4", 2),
    ("max_width", "<i4"), ("length", "<i4"), ("pid", "<i4"), ("type", "u1"), ("tex_idx", "u1"), ("pad", "u1", 2),
], align=False)
assert STREAK_DTYPE.itemsize == 128


def _vec(attr: str):
    return [float(t) for t in attr[1:-1].split(";")]


def _norm2(x, y):
    """np.linalg.norm of 2-vectors as NumPy's BLAS evaluates it (sqrt(fma(y, y, x*x)), rr_host_norm2)."""
    from . import _lib
    x, y = np.ascontiguousarray(x, np.float64), np.ascontiguousarray(y, np.float64)
    out = np.empty_like(x)
    _lib.load().rr_host_norm2(len(x), _lib.ptr(x), _lib.ptr(y), _lib.ptr(out))
    return out


def records_from_raw(wp1, wp2, ip1, ip2, iw1, iw2, pid, render_scale: int, W: int, H: int, return_mask: bool = False):
    """Simulator-convention arrays (image y up, world z negative forward, full sensor resolution) ->
    STREAK_DTYPE records, restricted to ``max_width >= 1 and length >= 1``: the arithmetic of
    DBManager.load_streaks_from_xml (bad_weather.py:208-238), vectorised."""
    n = len(pid)
    rec = np.zeros(n, dtype=STREAK_DTYPE)
    if n == 0:
        return (rec, np.zeros(0, bool)) if return_mask else rec
    wp1 = np.array(wp1, dtype=np.float64)
    wp2 = np.array(wp2, dtype=np.float64)
    ip1 = np.array(ip1, dtype=np.float64) / render_scale        # :208-209
    ip2 = np.array(ip2, dtype=np.float64) / render_scale
    iw1 = np.array(iw1, dtype=np.float64) / render_scale        # :210-211
    iw2 = np.array(iw2, dtype=np.float64) / render_scale
    ip1[:, 1] = H - ip1[:, 1]                                    # :221-222
    ip2[:, 1] = H - ip2[:, 1]
    wp1[:, 2] *= -1                                              # :223-224
    wp2[:, 2] *= -1
    diff = np.abs(ip1 - ip2)
    max_width = np.maximum(iw1, iw2).astype(np.int64)            # :226 int() truncation
    with np.errstate(divide="ignore", invalid="ignore"):
        nrm = _norm2(diff[:, 0], diff[:, 1])                                     # np.linalg.norm(diff), :229
        cos_theta = 0 * (diff[:, 0] / nrm) + -1 * (-(diff[:, 1] / nrm))          # :228-231
        ratio = max_width / (diff[:, 1] / cos_theta)                             # :232-233
    ip1r = np.round(ip1).astype(np.int64)                        # :234-235 (half-even)
    ip2r = np.round(ip2).astype(np.int64)
    d = (ip1r - ip2r).astype(np.float64)
    length = np.ceil(np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])).astype(np.int64)   # :236
    rec["wp1"], rec["wp2"], rec["iw1"], rec["iw2"] = wp1, wp2, iw1, iw2
    rec["ratio"] = ratio
    rec["ip1"], rec["ip2"], rec["ip1m"], rec["ip2m"] = ip1r, ip2r, ip1r, ip2r
    rec["max_width"], rec["length"] = max_width, length
    rec["pid"] = np.asarray(pid)
    rec["type"] = np.where(max_width >= 4, BIG, np.where(max_width > 1, MEDIUM, SMALL))   # :99-106
    keep = (max_width >= 1) & (length >= 1)
    return (rec[keep], keep) if return_mask else rec[keep]


def load_streaks_from_xml_py(path: str, render_scale: int, W: int, H: int, with_ids: bool = False):
    """Pure-Python statement of the loader (xml.etree + the vectorised arithmetic above): the comparand of
    the native loader in the CPU test-suite.  -> list (one entry per simulator frame, dict order of the
    frame ids) of STREAK_DTYPE arrays in dict order of the pids, restricted to ``max_width >= 1 and
    length >= 1``; a streak that fails that test never enters the dict, a later one with the same pid
    replaces the earlier entry in place (bad_weather.py:238), frames likewise by id (:241)."""
    frames = {}
    for frame in parse(path).getroot():
        fid = int(frame.attrib["id"])
        int(frame.attrib["t"]), int(frame.attrib["d"]), int(frame.attrib["rs"])      # read (and required) by the reference
        vals = [(_vec(a["wp1"]), _vec(a["wp2"]), _vec(a["ip1"]), _vec(a["ip2"]), float(a["iw1"]), float(a["iw2"]), int(a["pid"]),
                 float(a["wd1"]), float(a["wd2"])) for a in (drop.attrib for drop in frame)]
        if vals:
            n_all = len(vals)
            rec, kept = records_from_raw([v[0] for v in vals], [v[1] for v in vals], [v[2] for v in vals], [v[3] for v in vals],
                                         [v[4] for v in vals], [v[5] for v in vals], [v[6] for v in vals], render_scale, W, H,
                                         return_mask=True)
            assert len(kept) == n_all
            rows = {}
            for r in rec:
                rows[int(r["pid"])] = r
            rec = np.array(list(rows.values()), dtype=STREAK_DTYPE) if rows else np.zeros(0, dtype=STREAK_DTYPE)
        else:
            rec = np.zeros(0, dtype=STREAK_DTYPE)
        frames[fid] = rec
    if with_ids:
        return list(frames.values()), list(frames.keys())
    return list(frames.values())


CACHE_VERSION = "rr-b200-1"
CACHE_SUFFIX = ".rrcache.npz"


def _cache_key(path: str, render_scale: int, W: int, H: int):
    import hashlib
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return np.array([CACHE_VERSION, h.hexdigest(), "%d,%d,%d" % (int(W), int(H), int(render_scale))])


def load_streaks_from_xml(path: str, render_scale: int, W: int, H: int, with_ids: bool = False, use_cache: bool = False):
    """DBManager.load_streaks_from_xml (bad_weather.py:148-248) through the library's native loader
    (``rr_host_load_particles_xml``, csrc/rr_host_xml.cpp).  -> list, one entry per simulator frame in
    the order the reference's dict holds them, of STREAK_DTYPE arrays (tex_idx / noise_deg still zero: they
    are drawn per rendered frame).  ``with_ids`` also returns the frame ids.

    ``use_cache``: the reference's pickle cache (bad_weather.py:155-178) as a binary file of packed records beside the XML
    (``<xml>.rrcache.npz``), valid while the XML's md5, the image size and the render scale are unchanged (the same three
    conditions the reference checks: version, sim_hash, image_shapeWH); a stale or unreadable cache is rebuilt."""
    import ctypes as C
    from . import _lib
    cache, key = path + CACHE_SUFFIX, None
    if use_cache:
        key = _cache_key(path, render_scale, W, H)
        if os.path.exists(cache):
            try:
                with np.load(cache, allow_pickle=False) as z:
                    if np.array_equal(z["key"], key) and z["rec"].dtype == STREAK_DTYPE and z["hdr"].dtype == _lib.XML_FRAME_DTYPE:
                        hdr, rec = z["hdr"], z["rec"]
                        frames = [rec[int(f["first"]):int(f["first"] + f["count"])].copy() for f in hdr]
                        return (frames, [int(f["id"]) for f in hdr]) if with_ids else frames
                print("Particles cache out-dated. Regenerate.")
            except Exception:
                print("Particles cache unreadable. Regenerate.")
    lib = _lib.load()
    h = C.c_void_p()
    _lib.check(lib.rr_host_load_particles_xml(os.fsencode(path), int(render_scale), int(W), int(H), C.byref(h)), "rr_host_load_particles_xml")
    try:
        nf, nr = C.c_int32(0), C.c_int64(0)
        _lib.check(lib.rr_host_particles_info(h, C.byref(nf), C.byref(nr)), "rr_host_particles_info")
        hdr = np.zeros(nf.value, dtype=_lib.XML_FRAME_DTYPE)
        rec = np.zeros(nr.value, dtype=STREAK_DTYPE)
        _lib.check(lib.rr_host_particles_copy(h, _lib.ptr(hdr), _lib.ptr(rec)), "rr_host_particles_copy")
    finally:
        lib.rr_host_free_particles(h)
    if use_cache:
        try:
            tmp = cache + ".part.%d.npz" % os.getpid()
            np.savez(tmp, key=key, hdr=hdr, rec=rec)
            os.replace(tmp, cache)
        except OSError:
            pass                        # a read-only dataset tree: no cache, no error (the reference would raise here)
    frames = [rec[int(f["first"]):int(f["first"] + f["count"])].copy() for f in hdr]
    if with_ids:
        return frames, [int(f["id"]) for f in hdr]
    return frames


SIM_STREAK_DTYPE = np.dtype([("wp1", "<f8", 3), ("wp2", "<f8", 3), ("wd1", "<f8"), ("wd2", "<f8"), ("ip1", "<f8", 2), ("ip2", "<f8", 2),
                             ("iw1", "<f8"), ("iw2", "<f8"), ("pid", "<i8")])
assert SIM_STREAK_DTYPE.itemsize == 120


def records_from_sim(sim: np.ndarray, render_scale: int, W: int, H: int) -> np.ndarray:
    """rr_simulate_particles output of one frame (SIM_STREAK_DTYPE, any order) -> records, in pid
    order (the device appends with an atomic counter; sorting makes the frame deterministic)."""
    sim = sim[np.argsort(sim["pid"], kind="stable")]
    return records_from_raw(sim["wp1"], sim["wp2"], sim["ip1"], sim["ip2"], sim["iw1"], sim["iw2"], sim["pid"], render_scale, W, H)


def in_frame(rec: np.ndarray, W: int, H: int) -> np.ndarray:
    """Boolean mask of generator.py:413-420 (uses the *current*, possibly wind-mutated, positions)."""
    m = max(H, W)
    s, e = rec["ip1m"], rec["ip2m"]
    ok_w = (1 <= rec["max_width"]) & (rec["max_width"] < m)
    ok_l = (1 <= rec["length"]) & (rec["length"] < m)
    ins = (0 <= s[:, 0]) & (s[:, 0] < W) & (0 <= s[:, 1]) & (s[:, 1] < H)
    ine = (0 <= e[:, 0]) & (e[:, 0] < W) & (0 <= e[:, 1]) & (e[:, 1] < H)
    return ok_w & ok_l & (ins | ine)


def texture_buckets(ratio: np.ndarray, db_ratios: np.ndarray) -> np.ndarray:
    """First i with ratio < db_ratios[i], else 4 (bad_weather.py:251-265)."""
    return np.searchsorted(np.asarray(db_ratios[:4], dtype=np.float64), ratio, side="right").astype(np.int32)


def apply_wind_noise(rec: np.ndarray, noise_deg: np.ndarray):
    """generator.py:149-161 for the non-Big streaks of one frame: rotates the (already rounded)
    end points about their mid point by ``noise`` degrees and writes them back into integer
    storage (truncation).  ``ip1``/``ip2`` keep the pre-rotation values (the streak angle is taken
    from them, generator.py:138-144); ``ip1m``/``ip2m`` receive the result.  Returns rec."""
    rec["ip1"] = rec["ip1m"]
    rec["ip2"] = rec["ip2m"]
    sel = rec["type"] != BIG
    if not sel.any():
        return rec
    s = rec["ip1m"][sel].astype(np.float64)
    e = rec["ip2m"][sel].astype(np.float64)
    nz = noise_deg[sel]
    nx, ny = np.cos(np.deg2rad(nz)), np.sin(np.deg2rad(nz))
    mx = (e[:, 0] + s[:, 0]) / 2
    my = (e[:, 1] + s[:, 1]) / 2
    s_new = np.stack([(s[:, 0] - mx) * nx - (s[:, 1] - my) * ny + mx, (s[:, 0] - mx) * ny + (s[:, 1] - my) * nx + my], 1)
    e_new = np.stack([(e[:, 0] - mx) * nx - (e[:, 1] - my) * ny + mx, (e[:, 0] - mx) * ny + (e[:, 1] - my) * nx + my], 1)
    rec["ip1m"][sel] = s_new.astype(np.int64)       # assignment into an int array truncates toward zero
    rec["ip2m"][sel] = e_new.astype(np.int64)
    return rec
