"""Stand-in for natsort.natsorted (absent from this image)."""
import re


def _key(s):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", str(s))]


def natsorted(seq):
    return sorted(seq, key=_key)
