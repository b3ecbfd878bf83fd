"""Stand-in for pyclipper (absent from this image) used ONLY to run the reference here.
Delegates to oracle.clipper_rect (see that module for the parity statement)."""
from oracle.clipper_rect import intersect_with_rect

PT_CLIP, PT_SUBJECT = 1, 0
CT_INTERSECTION = 0
PFT_NONZERO = 1


class Pyclipper:
    def __init__(self):
        self._clip = None
        self._subj = None

    def AddPath(self, path, poly_type, closed=True):
        if poly_type == PT_CLIP:
            self._clip = path
        else:
            self._subj = path

    def Execute(self, clip_type, subj_fill=PFT_NONZERO, clip_fill=PFT_NONZERO):
        assert clip_type == CT_INTERSECTION
        xs = [p[0] for p in self._subj]
        ys = [p[1] for p in self._subj]
        assert min(xs) == 0 and min(ys) == 0, "stand-in only handles the (0,0)-(cols,rows) rectangle"
        return intersect_with_rect(self._clip, max(xs), max(ys))
