"""The native particles-XML loader (rr_host_load_particles_xml, csrc/rr_host_xml.cpp) against the
reference's own loader semantics (common/bad_weather.py:148-248), stated in Python by
rain_rendering_b200.streaks.load_streaks_from_xml_py and by the oracle.  Host logic only: no GPU."""
import os

import numpy as np
import pytest

from oracle import rain_oracle as ro
from rain_rendering_b200 import _lib, streaks as S, synth


def _same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x.dtype == y.dtype == S.STREAK_DTYPE and len(x) == len(y)
        assert x.tobytes() == y.tobytes() or all(
            np.array_equal(x[n], y[n], equal_nan=True) if x[n].dtype.kind == "f" else np.array_equal(x[n], y[n]) for n in S.STREAK_DTYPE.names)


@pytest.mark.parametrize("W,H,rs,n_xml", [(640, 480, 1, 1500), (1242, 375, 1, 700), (512, 256, 2, 1800)])
def test_native_loader_equals_python_loader_and_oracle(tmp_path, W, H, rs, n_xml):
    parts = synth.make_particles(W, H, 3, n_xml, 2.0, seed=4, render_scale=rs)
    xml = str(tmp_path / "sim_camera0.xml")
    synth.write_particles_xml(parts, xml, 2.0)
    nat, ids = S.load_streaks_from_xml(xml, rs, W, H, with_ids=True)
    py, ids_py = S.load_streaks_from_xml_py(xml, rs, W, H, with_ids=True)
    assert ids == ids_py == [0, 1, 2]
    _same(nat, py)
    assert sum(len(f) for f in nat) > 300
    # the oracle's Streak objects (bit-exact against the live reference, tests/test_oracle.py)
    for fr, ofr in zip(nat, ro.load_streaks_from_xml(xml, rs, W, H)):
        assert len(fr) == len(ofr)
        for r, s in zip(fr, ofr):
            assert r["pid"] == s.pid and r["max_width"] == s.max_width and r["length"] == s.length
            assert tuple(r["ip1"]) == tuple(s.ip1) and tuple(r["ip2"]) == tuple(s.ip2)
            assert np.array_equal(r["wp1"], s.wp1) and np.array_equal(r["wp2"], s.wp2)
            assert r["iw1"] == s.iw1 and r["iw2"] == s.iw2
            assert r["ratio"] == s.ratio or (np.isnan(r["ratio"]) and np.isnan(s.ratio))


HEAD = '<?xml version="1.0" encoding="utf-8"?>\n<!-- AHL output -->\n<camera statslevel="0">\n'


def _r(pid, ip1, ip2, iw1=2.5, iw2=2.0, extra=""):
    return ('<r pid="%d" wp1="[0.1;0.2;-1.5]" wd1="0.001" wp2="[0.1;0.15;-1.5]" wd2="0.001" ip1="[%s;%s]" iw1="%s" ip2="[%s;%s]" iw2="%s"%s/>'
            % (pid, ip1[0], ip1[1], iw1, ip2[0], ip2[1], iw2, extra))


def _write(tmp_path, body, name="a_camera0.xml"):
    p = str(tmp_path / name)
    with open(p, "w") as f:
        f.write(body)
    return p


def test_dict_semantics_order_quotes_and_filter(tmp_path):
    body = HEAD + "\n".join([
        '<i id="0" t="2000000" d="0" rs="5">',
        _r(7, (100.25, 300.5), (101.75, 260.5)),
        _r(3, (10, 20), (10.2, 20.1), iw1=0.4, iw2=0.3),                 # max_width 0: never enters the dict
        _r(9, (50.5, 51.5), (52.5, 11.5), iw1="4.75e0", iw2=" 1 "),      # Big; exponent and blanks like float()
        _r(7, (200, 300), (201, 250), extra=' label="dup"'),             # same pid: replaces the first entry IN PLACE
        _r(9, (5, 5), (5, 5)),                                           # same pid but length 0: filtered, the earlier 9 stays
        "<r pid='11' wp1='[1;2;-3]' wd1='1e-3' wp2='[1;2;-3.5]' wd2='1e-3' ip1='[7.5;8.5]' iw1='1.5' ip2='[7.5;30.5]' iw2='1.5'></r>",
        "</i>",
        '<i rs="0" d="10" t="2000000" id="1"/>',                         # empty frame, attributes in another order
        '<i id="2" t="2000000" d="20" rs="1">', "  text is ignored ", _r(1, (300, 100), (300, 60)), "<!-- comment --></i>",
        '<i id="0" t="2000000" d="30" rs="1">', _r(42, (30, 40), (31, 10)), "</i>",    # frame id 0 again: replaces frame 0 in place
        "</camera>\n"])
    p = _write(tmp_path, body)
    H, W = 375, 1242
    nat, ids = S.load_streaks_from_xml(p, 1, W, H, with_ids=True)
    py, ids_py = S.load_streaks_from_xml_py(p, 1, W, H, with_ids=True)
    assert ids == ids_py == [0, 1, 2]
    _same(nat, py)
    assert [int(x) for x in nat[0]["pid"]] == [42] and len(nat[1]) == 0 and [int(x) for x in nat[2]["pid"]] == [1]
    # the first version of frame 0 on its own: order 7, 9, 11 with the replaced 7
    p2 = _write(tmp_path, body.split('<i rs="0"')[0] + "</camera>", "b_camera0.xml")
    f0 = S.load_streaks_from_xml(p2, 1, W, H)[0]
    _same([f0], [S.load_streaks_from_xml_py(p2, 1, W, H)[0]])
    assert [int(x) for x in f0["pid"]] == [7, 9, 11]
    assert tuple(f0["ip1"][0]) == (200, H - 300) and f0["type"].tolist() == [S.MEDIUM, S.BIG, S.SMALL]
    assert f0["max_width"].tolist() == [2, 4, 1] and f0["length"][2] == 22
    assert np.array_equal(f0["wp1"][2], [1, 2, 3]) and np.array_equal(f0["wp2"][2], [1, 2, 3.5])     # z negated
    # half-even rounding of the image positions: 7.5 -> 8, 8.5 -> 8 (after the y flip: 375 - 8.5 = 366.5 -> 366)
    assert tuple(f0["ip1"][2]) == (8, 366)


@pytest.mark.parametrize("body", [
    HEAD + '<i id="0" t="1" d="0" rs="1"><r pid="1" wp1="[0;0;-1]"/></i></camera>',          # missing attributes
    HEAD + '<i id="0" t="1" d="0" rs="1">' + _r(1, (1, 2), (3, "x")) + "</i></camera>",        # not a number
    HEAD + '<i id="0" t="1" d="0">' + _r(1, (1, 2), (3, 4)) + "</i></camera>",                  # frame without rs
    HEAD + '<i id="0" t="1" d="0" rs="1">' + _r(1, (1, 2), (3, 4)),                              # truncated file
    HEAD + '<i id="0" t="1" d="0" rs="1">' + _r(1, (1, 2), (3, 4))[:-10],                        # cut inside a tag
    "",
])
def test_corrupted_files_are_rejected_like_the_reference(tmp_path, body):
    """bad_weather.py:183-187,243-247: a file that does not parse, or a streak that cannot be read, aborts the run
    with the advice to delete the simulation -- never a silently shorter frame."""
    p = _write(tmp_path, body)
    with pytest.raises(_lib.RainError) as e:
        S.load_streaks_from_xml(p, 1, 640, 480)
    assert "corrupted particles simulation" in str(e.value)
    with pytest.raises(Exception):
        S.load_streaks_from_xml_py(p, 1, 640, 480)


def test_missing_file(tmp_path):
    with pytest.raises(_lib.RainError):
        S.load_streaks_from_xml(str(tmp_path / "nope.xml"), 1, 640, 480)


def test_native_loader_is_much_faster_than_etree(tmp_path):
    import time
    parts = synth.make_particles(1242, 375, 20, 3000, 2.0, seed=1)
    xml = str(tmp_path / "big_camera0.xml")
    synth.write_particles_xml(parts, xml, 2.0)
    t0 = time.perf_counter(); a = S.load_streaks_from_xml(xml, 1, 1242, 375); t1 = time.perf_counter()
    b = S.load_streaks_from_xml_py(xml, 1, 1242, 375); t2 = time.perf_counter()
    _same(a, b)
    assert (t1 - t0) < (t2 - t1), (t1 - t0, t2 - t1)
    print("native %.3f s, etree+numpy %.3f s for %.1f MB" % (t1 - t0, t2 - t1, os.path.getsize(xml) / 1e6))


def test_binary_particles_cache_follows_the_reference_pickle_rules(tmp_path):
    """The reference caches the parsed simulation beside the XML and reuses it while version, md5 of the XML and image shape
    are unchanged (common/bad_weather.py:155-178); ours is a binary file of packed records with the same three conditions
    (plus the render scale, which the reference folds into the image shape)."""
    import os
    import time
    parts = synth.make_particles(640, 480, 3, 120, 2.0, seed=9)
    x = str(tmp_path / "sim_camera0.xml")
    synth.write_particles_xml(parts, x, 2.0)
    plain, ids = S.load_streaks_from_xml(x, 1, 640, 480, with_ids=True)
    assert not os.path.exists(x + S.CACHE_SUFFIX)
    first, ids1 = S.load_streaks_from_xml(x, 1, 640, 480, with_ids=True, use_cache=True)
    assert os.path.exists(x + S.CACHE_SUFFIX) and ids1 == ids
    t = os.path.getmtime(x + S.CACHE_SUFFIX)
    again, ids2 = S.load_streaks_from_xml(x, 1, 640, 480, with_ids=True, use_cache=True)
    assert os.path.getmtime(x + S.CACHE_SUFFIX) == t and ids2 == ids
    for a, b, c in zip(plain, first, again):
        assert a.tobytes() == b.tobytes() == c.tobytes()
    # another image shape or render scale: the cache is rebuilt, and holds the new records
    time.sleep(0.01)
    other = S.load_streaks_from_xml(x, 1, 512, 384, use_cache=True)
    assert other[0].tobytes() == S.load_streaks_from_xml(x, 1, 512, 384)[0].tobytes() and other[0].tobytes() != plain[0].tobytes()
    # a changed XML (md5) invalidates it
    parts2 = synth.make_particles(640, 480, 3, 80, 2.0, seed=10)
    synth.write_particles_xml(parts2, x, 2.0)
    new = S.load_streaks_from_xml(x, 1, 512, 384, use_cache=True)
    assert new[0].tobytes() == S.load_streaks_from_xml(x, 1, 512, 384)[0].tobytes() and len(new[0]) != len(other[0])
    # garbage in the cache file is not fatal
    open(x + S.CACHE_SUFFIX, "wb").write(b"not an npz")
    assert S.load_streaks_from_xml(x, 1, 512, 384, use_cache=True)[0].tobytes() == new[0].tobytes()
