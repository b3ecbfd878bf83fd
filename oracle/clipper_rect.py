"""Restatement of the one pyclipper call the hot path makes (TEST INFRASTRUCTURE).

Reference call site: common/bad_weather.py:363-373 --

    pc.AddPath(clip, PT_CLIP, True); pc.AddPath(subj, PT_SUBJECT, True)
    solution = pc.Execute(CT_INTERSECTION, PFT_NONZERO, PFT_NONZERO);  solution[0]

with ``subj`` the axis-aligned env-map rectangle (0,0)-(cols,rows) and ``clip`` the 20/24
vertex FOV polygon in float pixel coordinates.

pyclipper 1.0.6 (Clipper 6.4, Angus Johnson) is NOT installed in this image and is not
vendored by the reference, so this is *parity unpinned*: the behaviour below is restated
from Clipper's published semantics for this specific input class:

  * vertices are cast to 64-bit integers by C truncation (toward zero);
  * the intersection of a polygon with a rectangle that contains it is the polygon
    itself; otherwise it is clipped against the four half planes, intersection points
    rounded to the nearest integer (Clipper's ``Round``);
  * consecutive duplicate vertices and collinear vertices are removed
    (``FixupOutPolygon`` with PreserveCollinear = false);
  * outer contours are returned with positive Clipper ``Area`` (counter-clockwise in
    a y-up frame).

Self-intersecting inputs (a view cone straddling the +-pi azimuth seam) are passed
through without the decomposition Vatti clipping would perform.  The CUDA path
restates exactly this function (csrc/streak_geom.cuh: clip_fov_polygon).
"""
from __future__ import annotations


def _trunc(v: float) -> int:
    return int(v)  # Python int() truncates toward zero, like the C cast in pyclipper


def _round_half_away(v: float) -> int:
    # Clipper::Round: (val < 0) ? (cInt)(val - 0.5) : (cInt)(val + 0.5)
    return int(v - 0.5) if v < 0 else int(v + 0.5)


def _clip_halfplane(pts, axis, bound, keep_less):
    out = []
    n = len(pts)
    for i in range(n):
        a = pts[i]
        b = pts[(i + 1) % n]
        ina = (a[axis] <= bound) if keep_less else (a[axis] >= bound)
        inb = (b[axis] <= bound) if keep_less else (b[axis] >= bound)
        if ina:
            out.append(a)
        if ina != inb:
            t = (bound - a[axis]) / float(b[axis] - a[axis])
            o = 1 - axis
            p = [0, 0]
            p[axis] = bound
            p[o] = _round_half_away(a[o] + t * (b[o] - a[o]))
            out.append((p[0], p[1]))
    return out


def clean_polygon(pts):
    """Remove consecutive duplicates and collinear vertices (repeat until stable)."""
    pts = list(pts)
    changed = True
    while changed and len(pts) >= 3:
        changed = False
        n = len(pts)
        for i in range(n):
            p, c, nx = pts[i - 1], pts[i], pts[(i + 1) % n]
            if c == p or c == nx or (c[1] - p[1]) * (nx[0] - c[0]) == (c[0] - p[0]) * (nx[1] - c[1]):
                del pts[i]
                changed = True
                break
    return pts


def area2(pts):
    """Twice Clipper's Area() (positive = counter-clockwise in a y-up frame)."""
    a = 0
    n = len(pts)
    for i in range(n):
        j = i - 1
        a += (pts[j][0] + pts[i][0]) * (pts[j][1] - pts[i][1])
    return -a


def intersect_with_rect(clip_pts, cols: int, rows: int):
    """Returns a list of integer paths (lists of (x, y)); empty list if nothing remains."""
    pts = [(_trunc(p[0]), _trunc(p[1])) for p in clip_pts]
    if any(p[0] < 0 for p in pts):
        pts = _clip_halfplane(pts, 0, 0, False)
    if pts and any(p[0] > cols for p in pts):
        pts = _clip_halfplane(pts, 0, cols, True)
    if pts and any(p[1] < 0 for p in pts):
        pts = _clip_halfplane(pts, 1, 0, False)
    if pts and any(p[1] > rows for p in pts):
        pts = _clip_halfplane(pts, 1, rows, True)
    pts = clean_polygon(pts)
    if len(pts) < 3:
        return []
    if area2(pts) < 0:
        pts = pts[::-1]
    return [[list(p) for p in pts]]
