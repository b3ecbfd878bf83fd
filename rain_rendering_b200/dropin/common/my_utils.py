"""common.my_utils of the reference, hot-path part (common/my_utils.py:55-96,172) plus the few
path helpers its callers use.  The colour conversions are 3x3 host utilities kept in numpy for
API compatibility; the renderer performs them on the GPU (csrc/rr_kernels.cu: k_env_prefix, rr_tint)."""
import os
import re

import numpy as np

try:
    from natsort import natsorted
except Exception:                      # natsort is optional: natural sort restated
    def natsorted(seq):
        return sorted(seq, key=lambda s: [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", str(s))])


def path_os_s(path):
    return re.sub(r"[/|\\]+", re.escape(os.sep) if os.sep == "\\" else os.sep, path)


def os_listdir(path):
    return natsorted(os.listdir(path))


def print_error(msg):
    print('\n\x1b[2;30;41m[ERROR]\x1b[0m  %s' % msg)


def print_success(msg):
    print('\n\x1b[2;30;42m[SUCCESS]\x1b[0m  %s' % msg)


def print_warning(msg):
    print('\x1b[2;30;43m[WARNING]\x1b[0m  %s' % msg)


_M = np.array([[0.49000, 0.31000, 0.20000], [0.17697, 0.81240, 0.01063], [0.00000, 0.01000, 0.99000]])
_M2 = np.array([[0.41847, -0.15866, -0.082835], [-0.091169, 0.25243, 0.015708], [0.0009209, -0.0025498, 0.1786]])


def convert_rgb_to_xyY(array):
    XYZ = np.dot(array, _M) / 0.17697
    X, Y, Z = XYZ[..., 0], XYZ[..., 1], XYZ[..., 2]
    with np.errstate(divide='ignore', invalid='ignore'):
        x = X / (X + Y + Z)
        y = Y / (X + Y + Z)
    return np.concatenate([x[..., None], y[..., None], Y[..., None]], axis=-1)


def convert_xyY_to_rgb(xyY):
    x, y, Y = xyY[..., 0], xyY[..., 1], xyY[..., 2]
    with np.errstate(divide='ignore', invalid='ignore'):
        X = (Y * x) / y
        Z = (Y * (1 - x - y)) / y
    return np.dot(np.concatenate([X[..., None], Y[..., None], Z[..., None]], axis=-1), _M2)


def crop_center(image, height, width):
    x1 = int((image.shape[0] - height) / 2)
    y1 = int((image.shape[1] - width) / 2)
    return image[x1:x1 + height, y1:y1 + width]


def particles_path(path, weather):
    return os.path.join(path, weather["weather"], "{}mm".format(weather["fallrate"]), '*_camera0.xml')
