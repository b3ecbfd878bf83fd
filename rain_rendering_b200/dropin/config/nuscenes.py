"""config.nuscenes of the reference (config/nuscenes.py), repaired so that BASELINE config C4 (nuScenes CAM_FRONT, rate
sweep) can be driven through the reference's own ``main.py``.

What is broken upstream and what this module does instead:
  * ``config/nuscenes.py`` sits next to a DIRECTORY ``config/nuscenes/``; ``import config.nuscenes`` picks the module, and its
    first statement ``from config.nuscenes.nusc_dataset import ...`` (config/nuscenes.py:4) then fails -- a module has no
    submodules.  Here the dataset helper is loaded by file path, and only when the nuscenes-devkit is installed.
  * ``results.json_file`` (config/nuscenes.py:28) is read although ``main.py`` has that option commented out (main.py:30):
    ``getattr(results, "json_file", None)``.
  * ``results.particles`` is set to a dict built from four undefined names (``scene_token, cameras, motions, durations``,
    config/nuscenes.py:56) -- a NameError -- and would break ``main.py:176`` (``os.path.join(results.particles, ...)``) if it
    did not: left alone, the generic resolution of ``common/db.py:sim`` + ``main.py:176-209`` applies.
  * ``main.py:152-158`` tests ``os.path.exists(results.images[seq])`` on what for nuScenes is a LIST of files: the lists
    returned here answer ``os.fspath()`` with the directory of their first file.
Without the devkit the frames come from a plain index: ``<dataset_root>/rain_b200_index.json`` (or ``--json_file``) holding
``{"scenes": {"<scene>": ["samples/CAM_FRONT/<file>.jpg", ...]}}``; with neither, every ``samples/CAM_FRONT`` image in
natural order forms one sequence named ``CAM_FRONT``.  Depth maps are ``<depth_root>/<image stem>.npy`` (float32 metres), as
upstream (config/nuscenes.py:59).  ``settings()`` is the reference's (config/nuscenes.py:64-86).
"""
import importlib.util
import json
import os

from common import my_utils


class PathList(list):
    """A list of files that ``os.path.exists`` / ``os.fspath`` treat as the directory holding them (main.py:152-158)."""

    def __fspath__(self):
        return os.path.dirname(self[0]) if len(self) else os.devnull + ".missing"


def _devkit_dataset(results):
    """NuScenesDataset of the reference's helper, or None when it (or the nuscenes-devkit) cannot be had."""
    import config as _pkg
    for base in _pkg.__path__:
        path = os.path.join(base, "nuscenes", "nusc_dataset.py")
        if os.path.isfile(path):
            try:
                spec = importlib.util.spec_from_file_location("rain_b200_nusc_dataset", path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
            except Exception:
                return None
            tokens = None
            jf = getattr(results, "json_file", None)
            if jf:
                with open(jf) as f:
                    tokens = json.load(f).get("sample_data_tokens")
            return mod.NuScenesDataset(version="v1.0-trainval", root=results.dataset_root, pretransform_data=False, preload_data=False,
                                       only_annotated=False, specific_tokens=tokens)
    return None


def _index(results):
    """{scene: [image paths relative to dataset_root]} without the devkit."""
    root = results.dataset_root
    for cand in (getattr(results, "json_file", None), os.path.join(root, "rain_b200_index.json")):
        if cand and os.path.isfile(cand):
            with open(cand) as f:
                data = json.load(f)
            if "scenes" in data:
                return {str(k): list(v) for k, v in data["scenes"].items()}
    cam = os.path.join(root, "samples", "CAM_FRONT")
    assert os.path.isdir(cam), ("nuScenes frames not found: no nuscenes-devkit, no index file and no " + cam)
    files = [f for f in my_utils.os_listdir(cam) if f.lower().endswith((".jpg", ".jpeg", ".png"))]
    return {"CAM_FRONT": [os.path.join("samples", "CAM_FRONT", f) for f in files]}


def resolve_paths(results):
    if "gan" in results.dataset:
        raise NotImplementedError("nuscenes_gan (bad_weather.py:213-219 rescaling) is out of scope of the B200 path")
    root = results.dataset_root
    ds = _devkit_dataset(results)
    if ds is not None:
        scenes = {s: list(ds.get_filepaths(s, "CAM_FRONT")) for s in sorted(set(ds.scene_tokens))}
    else:
        scenes = _index(results)
    unique = sorted(scenes)
    if getattr(results, "sequences", None):
        want = [s for s in str(results.sequences).split(",") if s != ""]
        if want and want[0].isnumeric():                       # config/nuscenes.py:11-17: indices into the sorted scene list
            want = [unique[int(s)] for s in want if int(s) < len(unique)]
        picked = [s for s in unique if any(s[:len(w)] == w for w in want)] if want else unique
    else:
        picked = unique
    assert len(picked) > 0, "There are no valid sequences folder in the dataset root."
    results.sequences = picked
    results.images = {s: PathList(os.path.join(root, p) for p in scenes[s]) for s in picked}
    results.depth = {s: PathList(os.path.join(results.depth_root, os.path.splitext(os.path.basename(p))[0] + ".npy") for p in scenes[s]) for s in picked}
    results.calib = {s: None for s in picked}
    return results


def settings():
    settings = {}
    settings["cam_focal"] = 5.5          # Focal length (mm)
    settings["cam_gain"] = 1.0
    settings["cam_f_number"] = 1.8       # F-Number
    settings["cam_focus_plane"] = 6.0    # Focus plane (meter)
    settings["cam_exposure"] = 5.0       # Camera exposure (ms)
    settings["cam_pos"] = [1.5, 1.5, 0.3]
    settings["cam_lookat"] = [1.5, 1.5, -1.]
    settings["cam_up"] = [0., 1., 0.]
    settings["cam_WH"] = [1600, 900]     # CAM_FRONT frames (the reference leaves KITTI's default, common/db.py:13: its simulator would image the wrong sensor)
    settings["cam_CCD_WH"] = [1600, 900]
    settings["sequences"] = {}
    return settings
