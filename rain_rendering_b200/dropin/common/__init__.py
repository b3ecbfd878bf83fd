"""Drop-in ``common`` package: put ``rain_rendering_b200/dropin`` FIRST on ``PYTHONPATH`` and the
reference's ``main.py`` runs unchanged on the B200 path --

    PYTHONPATH=/path/to/repo/rain_rendering_b200/dropin:/path/to/repo python main.py --dataset kitti ...

``common.generator`` / ``common.bad_weather`` / ``common.add_attenuation`` / ``common.solid_angle``
are served from here (the hot path).  Everything else the reference's callers import from
``common`` (``common.db``, the dataset plumbing; ``common.drop_depth_map``) is out of scope and
resolves to the reference's own files: this package extends ``__path__`` with the reference's
``common`` directory (``RAIN_REFERENCE_ROOT`` or the first other ``common`` package on ``sys.path``).
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.dirname(os.path.dirname(os.path.dirname(_here)))
if _repo not in sys.path:
    sys.path.append(_repo)      # so that ``rain_rendering_b200`` itself is importable


def _reference_common():
    root = os.environ.get("RAIN_REFERENCE_ROOT")
    cands = [os.path.join(root, "common")] if root else []
    for p in sys.path:
        c = os.path.join(p or os.getcwd(), "common")
        if os.path.abspath(c) != _here:
            cands.append(c)
    for c in cands:
        if os.path.isfile(os.path.join(c, "db.py")) and os.path.isfile(os.path.join(c, "generator.py")):
            return os.path.abspath(c)
    return None


_ref = _reference_common()
if _ref and _ref not in __path__:
    __path__.append(_ref)
