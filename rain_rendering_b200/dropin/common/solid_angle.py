"""common.solid_angle.get_solid_angles (common/solid_angle.py:5-29) computed on the GPU
(csrc/rr_kernels.cu: k_omega) for the (H, W) lat-long grid of the given map."""
import numpy as np

_cache = {}


def get_solid_angles(img):
    from rain_rendering_b200 import _lib
    from rain_rendering_b200.api import RainContext
    H, W_env = int(img.shape[0]), int(img.shape[1])
    if (H, W_env) not in _cache:
        ctx = RainContext(0)
        out = np.empty((H, W_env), np.float64)
        _lib.check(ctx.lib.rr_solid_angles(ctx.h, H, W_env, _lib.ptr(out)), "rr_solid_angles")
        ctx.close()
        _cache[(H, W_env)] = out
    return _cache[(H, W_env)].copy()
