"""Drop-in ``config`` package: only ``config.nuscenes`` is served from here (the reference's plug-in cannot be imported at
all -- see that module); ``config.kitti`` / ``config.cityscapes`` / ``config.customdb`` resolve to the reference's own files:
this package extends ``__path__`` with the reference's ``config`` directory (``RAIN_REFERENCE_ROOT`` or the first other
``config`` package on ``sys.path``), exactly like the drop-in ``common`` package does."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))


def _reference_config():
    root = os.environ.get("RAIN_REFERENCE_ROOT")
    cands = [os.path.join(root, "config")] if root else []
    for p in sys.path:
        c = os.path.join(p or os.getcwd(), "config")
        if os.path.abspath(c) != _here:
            cands.append(c)
    for c in cands:
        if os.path.isfile(os.path.join(c, "kitti.py")) and os.path.isfile(os.path.join(c, "customdb.py")):
            return os.path.abspath(c)
    return None


_ref = _reference_config()
if _ref and _ref not in __path__:
    __path__.append(_ref)
