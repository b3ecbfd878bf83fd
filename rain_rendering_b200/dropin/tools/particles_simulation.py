"""Drop-in ``tools.particles_simulation``: the module the reference's ``main.py`` imports when a particles XML is
missing (main.py:194-209).  The reference drives the closed ``AHLSimulation`` binary through pexpect menus
(tools/simulation.py:259-469) for minutes per weather; here ``process()`` runs the library's particle simulator
(``rr_simulate_particles``, csrc/rr_sim.cu) and writes the very file the rest of the reference
expects -- ``<particles>/<dataset>/<seq>/<weather>/<R>mm/*_camera0.xml`` in the simulator's XML schema -- so
``main.py`` continues unchanged (it globs the file, ``DBManager.load_streaks_from_xml`` parses it).

``tools`` is a namespace package in the reference (no ``__init__.py``): with ``rain_rendering_b200/dropin`` first on
``PYTHONPATH`` this file shadows ``tools/particles_simulation.py`` while ``tools.simulation`` still resolves to the
reference's own file.

Statistical stand-in only: the binary cannot run outside its 2017 library set and its RNG is not reproducible.
"""
import json
import os
import sys

import numpy as np

force_recompute = False
particles_root = os.path.join('data', 'particles')

_ctx = None


def _context():
    global _ctx
    if _ctx is None:
        from rain_rendering_b200 import api
        _ctx = api.RainContext(int(os.environ.get("LOCAL_RANK", os.environ.get("RAIN_B200_DEVICE", "0"))))
    return _ctx


def n_camera_frames(options):
    """normal mode: sim_duration seconds at cam_hz frames per second; steps mode: one camera frame per step
    (common/db.py:44-66, tools/simulation.py:226,365-389)."""
    steps = options.get("sim_steps") or {}
    if options.get("sim_mode", "normal") == "steps" and steps:
        return int(max(len(v) for v in steps.values()))
    return max(1, int(round(float(options["sim_duration"]) * float(options["cam_hz"]))))


def per_frame_values(options, key, default, n):
    """sim_steps[key][i] applies from step i on and stays applied (common/db.py:57-58)."""
    steps = (options.get("sim_steps") or {}).get(key)
    out = np.full(n, float(default))
    if steps is not None and len(steps):
        v = np.asarray(steps, dtype=np.float64)
        m = min(n, len(v))
        out[:m] = v[:m]
        out[m:] = v[m - 1]
    return out


def _vec(a):
    return "[" + ";".join(repr(float(t)) for t in a) + "]"


def write_sim_xml(frames, path, exposure_ms, cam_hz):
    """frames: list of SIM_STREAK_DTYPE arrays (rr_sim_streak: the attributes of the simulator's <r> element).
    Floats are written with repr(), so the loader reads back the very doubles the device produced."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    tmp = path + ".part"
    with open(tmp, "w") as f:
        f.write('<?xml version="1.0"?>\n<camera statslevel="0">\n')
        for i, fr in enumerate(frames):
            f.write('  <i id="%d" t="%d" d="%d" rs="%d">\n' % (i, int(round(exposure_ms * 1e6)), int(round(i * 1e9 / cam_hz)), len(fr)))
            for r in fr:
                f.write('    <r pid="%d" wp1="%s" wd1="%r" wp2="%s" wd2="%r" ip1="%s" iw1="%r" ip2="%s" iw2="%r" />\n' % (
                    int(r["pid"]), _vec(r["wp1"]), float(r["wd1"]), _vec(r["wp2"]), float(r["wd2"]),
                    _vec(r["ip1"]), float(r["iw1"]), _vec(r["ip2"]), float(r["iw2"])))
            f.write("  </i>\n")
        f.write("</camera>\n")
    os.replace(tmp, path)


def simulate_to_xml(out_root, options, weather, redo=False, ctx=None):
    """One (sequence, weather): -> path of the XML written, or None when a simulation file already exists and
    ``redo`` is false (tools/simulation.py:262-269)."""
    out_dir = os.path.join(out_root, weather["weather"], "{}mm".format(weather["fallrate"]))
    os.makedirs(out_dir, exist_ok=True)
    if not redo and any("camera0.xml" in f for f in os.listdir(out_dir)):
        print("Simulation file exits {}, next!".format(out_dir))
        return None
    try:        # tools/simulation.py:271-281
        with open(os.path.join(out_dir, "sim_options.json"), "w") as fp:
            json.dump({k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in options.items() if k != "sequences"}, fp, default=str)
    except Exception as e:
        print(e)
    ctx = ctx or _context()
    n = n_camera_frames(options)
    W, H = int(options["cam_WH"][0]), int(options["cam_WH"][1])
    speed = per_frame_values(options, "cam_motion", 0.0, n)
    exposure = per_frame_values(options, "cam_exposure", options["cam_exposure"], n)
    focal = per_frame_values(options, "cam_focal", options["cam_focal"], n)
    rate = per_frame_values(options, "rain_fallrate", weather["fallrate"], n)
    frames = []
    i = 0
    while i < n:                    # one device call per run of frames that share their step parameters
        j = i + 1
        while j < n and (speed[j], exposure[j], focal[j], rate[j]) == (speed[i], exposure[i], focal[i], rate[i]):
            j += 1
        got, _ = ctx.simulate_particles(i, j - i, W, H, rate[i], focal_mm=focal[i], pix_size_um=float(options["cam_CCD_pixsize"]),
                                        exposure_ms=exposure[i], sim_hz=float(options["sim_hz"]), cam_speed_kmh=speed[i], seed=0)
        frames += got
        i = j
    path = os.path.join(out_dir, "b200sim_%.1fms_%dfps_camera0.xml" % (float(options["cam_exposure"]), int(options["cam_hz"])))
    write_sim_xml(frames, path, float(options["cam_exposure"]), float(options["cam_hz"]))
    sys.stdout.write(" simulated %d camera frames, %.0f streaks per frame -> %s\n" % (n, np.mean([len(f) for f in frames]) if frames else 0, path))
    return path


def process(sim, force_recompute=False):
    """Same signature and loops as the reference (tools/particles_simulation.py:23-73): every weather x every path."""
    path, options, weathers = sim["path"], sim["options"], sim["weather"]
    written = []
    for weather in weathers:
        for i in range(len(path)):
            p = simulate_to_xml(path[i], options[i], weather, redo=force_recompute)
            if p:
                written.append(p)
    print("All threads completed")
    return written


def process_sequences(sequences, weathers, force_recompute=False):
    """(dataset, sequence) pairs -> their simulation folders and per-sequence options (``common.db.sim``), then
    ``process`` -- the entry point of the reference's stand-alone script (tools/particles_simulation.py:8-20)."""
    from common import my_utils, db
    sims = [db.sim(name, my_utils.path_os_s(seq), os.path.join(particles_root, name)) for name, seq in sequences]
    return process({"path": [s["path"] for s in sims], "options": [s["options"] for s in sims], "weather": weathers},
                   force_recompute=force_recompute)
