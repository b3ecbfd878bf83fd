"""Run the UNTOUCHED reference (``/root/reference``) in this container (TEST INFRASTRUCTURE).

The reference cannot travel to the GPU box, so this module is only used here, by
``oracle/make_golden.py`` (which commits small golden fixtures under ``tests/golden``) and by
the CPU tests that pin the oracle restatement against the live reference when it is present.

Stand-ins: ``oracle/ref_shims`` supplies natsort / glob2 /
matplotlib.pyplot / pexpect (trivial), and pyclipper / imutils (restated -- parity unpinned for
those two).  numpy>=1.24 compatibility: ``np.int``/``np.float`` aliases and an object-array proxy
for the ragged texture list at common/bad_weather.py:146.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("RAIN_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "common", "generator.py"))


class _NpProxy(types.ModuleType):
    """numpy look-alike whose ``array`` falls back to dtype=object for ragged lists."""

    def __init__(self):
        super().__init__("numpy_proxy")

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def array(obj, *a, **k):
        try:
            return np.array(obj, *a, **k)
        except ValueError:
            out = np.empty(len(obj), dtype=object)
            for i, o in enumerate(obj):
                out[i] = o
            return out


_loaded = {}


def load_reference():
    """Import the reference's ``main`` and ``common.*`` modules with the stand-ins installed."""
    if _loaded:
        return _loaded
    assert reference_available(), "reference tree not found at %s" % REFERENCE_ROOT
    if not hasattr(np, "int"):
        np.int = int          # used at common/bad_weather.py:834,848
    if not hasattr(np, "float"):
        np.float = float      # used at common/generator.py:384
    if not hasattr(np, "bool"):
        np.bool = bool
    for p in (_REPO, REFERENCE_ROOT, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    # make sure a previously imported drop-in ``common`` package does not shadow the reference
    for k in [k for k in sys.modules if k == "common" or k.startswith("common.")]:
        del sys.modules[k]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import common.bad_weather as bw          # noqa
        import common.generator as gen           # noqa
        import common.add_attenuation as att     # noqa
        import common.solid_angle as sa          # noqa
        import common.my_utils as mu             # noqa
        import main as ref_main                  # noqa
    assert os.path.abspath(bw.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), bw.__file__
    bw.np = _NpProxy()
    import matplotlib.pyplot as plt
    _loaded.update(dict(bw=bw, gen=gen, att=att, sa=sa, mu=mu, main=ref_main, plt=plt))
    return _loaded


def run_reference(paths: dict, dataset: str, fallrate: int, frames=None, capture_stages=True,
                  noise_scale=0.0, noise_std=0.0, opacity_attenuation=1.0, sequences="seq1", settings_override=None):
    """Drive ``main.check_arg`` + ``Generator.run()`` of the reference on a tree laid out by
    ``rain_rendering_b200.synth.write_dataset``.  Returns {frame_name: {...arrays...}}."""
    m = load_reference()
    plt, bw, gen, att = m["plt"], m["bw"], m["gen"], m["att"]
    plt.captured.clear()
    argv = ["--dataset", dataset, "-k", paths["dataset_root"], "-r", paths["particles"],
            "-sd", paths["streaks_db"], "-d", paths["dataset_root"], "--output", paths["output"],
            "-i", str(fallrate), "-s", sequences, "--noverbose", "--conflict_strategy", "overwrite",
            "-ns", repr(noise_scale), "-nv", repr(noise_std), "-oa", repr(opacity_attenuation)]
    if frames:
        argv += ["-ff", ",".join(str(f) for f in frames)]
    stages = {"fog": [], "env": [], "streaks": []}

    orig_fog = att.FogRain.fog_rain_layer
    orig_env = bw.EnvironmentMapGenerator.generate_map
    orig_add = bw.RainRenderer.add_drop_to_image

    def fog_wrap(self, image, depth):
        out = orig_fog(self, image, depth)
        stages["fog"].append(out.copy())
        stages["streaks"].append([])
        return out

    def env_wrap(self, background):
        out = orig_env(self, background)
        stages["env"].append(out.copy())
        return out

    def add_wrap(self, dataset_, env_map_xyY, solid_angle_map, drop_fov_pts, drop_minC, bg, rainy_bg,
                 rainy_mask, rainy_saturation_mask, drop, drop_dict, *a, **k):
        rec = dict(pid=drop_dict.pid, fov_pts=np.array(drop_fov_pts, copy=True),
                   minC_in=np.array(drop_minC, copy=True), patch_in=drop[..., 3].copy())
        try:
            res = orig_add(self, dataset_, env_map_xyY, solid_angle_map, drop_fov_pts, drop_minC, bg, rainy_bg,
                           rainy_mask, rainy_saturation_mask, drop, drop_dict, *a, **k)
        except Exception:
            rec["error"] = True
            stages["streaks"][-1].append(rec)
            raise
        rec["minC_out"] = np.array(res[5], copy=True)
        rec["patch_out"] = np.array(res[3], copy=True)     # blurred, cropped BGR+alpha actually composited
        stages["streaks"][-1].append(rec)
        return res

    if capture_stages:
        att.FogRain.fog_rain_layer = fog_wrap
        bw.EnvironmentMapGenerator.generate_map = env_wrap
        bw.RainRenderer.add_drop_to_image = add_wrap
    import common.db as ref_db
    saved_defaults = dict(ref_db._settings_defaults)
    if settings_override:
        ref_db._settings_defaults.update(settings_override)     # customdb does not set these keys, so the defaults apply
    cwd = os.getcwd()
    try:
        os.chdir(REFERENCE_ROOT)  # config.<dataset> modules are imported relative to the reference root
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            args = m["main"].check_arg(argv)
            g = gen.Generator(args)
            g.run()
    finally:
        os.chdir(cwd)
        ref_db._settings_defaults.clear()
        ref_db._settings_defaults.update(saved_defaults)
        att.FogRain.fog_rain_layer = orig_fog
        bw.EnvironmentMapGenerator.generate_map = orig_env
        bw.RainRenderer.add_drop_to_image = orig_add
    out = {}
    rainy = {os.path.basename(p)[:-4]: a for p, a in plt.captured.items() if os.sep + "rainy_image" + os.sep in p}
    masks = {os.path.basename(p)[:-4]: a for p, a in plt.captured.items() if os.sep + "rain_mask" + os.sep in p}
    for i, name in enumerate(sorted(rainy)):
        rec = dict(rainy_rgb=rainy[name], rain_mask=masks[name])
        if capture_stages:
            rec["fog"] = stages["fog"][i]
            rec["env"] = stages["env"][i]
            rec["streaks"] = stages["streaks"][i]
        out[name] = rec
    return out
