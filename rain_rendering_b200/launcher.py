"""``main_threaded.py`` of the reference, re-targeted at GPUs: the same command line, the same job
split, the same per-job log files -- but the children run the reference's ``main.py`` against the
B200 drop-in (``rain_rendering_b200/dropin`` first on ``PYTHONPATH``) and are spread over the GPUs of
the box instead of over ten CPU processes.

    python -m rain_rendering_b200.launcher --dataset kitti --intensity 1,5,25 \
        --scene_threaded --frame_start 0 --frame_end 7481 [... any main.py argument ...]

Reference behaviour kept (main_threaded.py, astra-vision/rain-rendering):
  * one job per intensity (:160-174); with ``--scene_threaded`` one job per 41-frame range and intensity
    (:109-139), ``--frame_end`` required (:110)
  * every job gets ``--conflict_strategy skip --noverbose`` (:121-129,164-167), ``-v`` and
    ``--scene_threaded`` / ``--scenes_per_thread`` are stripped (:123,127,152-154)
  * stdout / stderr of a job go to ``automate_log_<R>mm_<start>_to_<end>[_step_<s>].txt`` and
    ``automate_error_...txt`` in the working directory (:23-32)
Differences: the reference builds five identical jobs per (range, intensity) because its scene loop
(:118,147-150) never appends ``--sequences`` (``"sequences" in args`` is always false); identical
jobs are emitted once.  Its ``--frames`` + ``--scene_threaded`` branch reads a non-existent ``args.jump``
(:140) and raises; here ``--frames`` is passed through untouched.  Concurrency is ``gpus x
jobs_per_gpu`` (default 2 per GPU: one decodes while the other renders) instead of 10.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import time

FRAMES_PER_JOB = 41          # main_threaded.py:111
HERE = os.path.dirname(os.path.abspath(__file__))


def parse(argv):
    """The launcher's own view of the command line (main_threaded.py:52-95); everything else is main.py's."""
    p = argparse.ArgumentParser(description="Rain renderer, GPU job fan-out", add_help=False)
    p.add_argument("--intensity", type=str, required=True)
    p.add_argument("--scene_threaded", action="store_true")
    p.add_argument("--frame_start", type=int, default=0)
    p.add_argument("--frame_end", type=int, default=None)
    p.add_argument("--frame_step", type=int)
    p.add_argument("--frames", type=str)
    p.add_argument("--scenes_per_thread", type=int, default=25)
    # launcher-only options (stripped from the children's command line)
    p.add_argument("--gpus", type=int, default=None, help="GPUs to use (default: all visible)")
    p.add_argument("--jobs_per_gpu", type=int, default=2)
    p.add_argument("--main", type=str, default="main.py", help="path of the reference's main.py")
    res, _ = p.parse_known_args(argv)
    res.intensity = [int(i) for i in res.intensity.split(",")]
    return res


def _strip(args, flag, n_values):
    out, i = [], 0
    while i < len(args):
        if args[i] == flag:
            i += 1 + n_values
            continue
        out.append(args[i])
        i += 1
    return out


def _set(args, flag, value):
    """Replace the value of ``flag``; append the pair when the flag is absent (the reference's
    ``_new_args[_new_args.index(flag) + 1] = ...`` raises in that case, e.g. a defaulted --frame_start)."""
    args = list(args)
    if flag in args:
        args[args.index(flag) + 1] = str(value)
    else:
        args += [flag, str(value)]
    return args


def build_jobs(argv):
    """-> list of child argument lists (for ``python main.py <args>``), in the reference's order."""
    a = parse(argv)
    base = list(argv)
    for flag, n in (("--scenes_per_thread", 1), ("--gpus", 1), ("--jobs_per_gpu", 1), ("--main", 1)):
        base = _strip(base, flag, n)
    jobs = []
    if a.scene_threaded:
        assert a.frame_end or a.frames, "--scene_threaded needs --frame_end (or --frames)"      # main_threaded.py:110
        base = _strip(_strip(base, "--scene_threaded", 0), "-v", 0)
        if a.frames:
            ranges = [(None, None)]
        else:
            ranges = [(s, min(s + FRAMES_PER_JOB, a.frame_end)) for s in range(a.frame_start, a.frame_end, FRAMES_PER_JOB)]
        for f0, f1 in ranges:
            for intensity in a.intensity:
                args = base + ["--conflict_strategy", "skip", "--noverbose"]
                args = _set(args, "--intensity", intensity)
                if f0 is not None:
                    args = _set(_set(args, "--frame_start", f0), "--frame_end", f1)
                if args not in jobs:
                    jobs.append(args)
    else:
        for intensity in a.intensity:
            args = _set(base + ["--conflict_strategy", "skip", "--noverbose"], "--intensity", intensity)
            jobs.append(args)
    return jobs


def log_pattern(args):
    """main_threaded.py:23-27"""
    d = {args[i]: args[i + 1] for i in range(0, len(args) - 1)}
    pat = "{}mm_{}_to_{}".format(d.get("--intensity", "NA"), d.get("--frame_start", 0), d.get("--frame_end", "NA"))
    if d.get("--frame_step"):
        pat += "_step_{}".format(d.get("--frame_step"))
    return pat


def visible_gpus():
    env = os.environ.get("RAIN_B200_GPUS")
    if env:
        return int(env)
    try:
        import torch
        n = torch.cuda.device_count()
        if n:
            return n
    except Exception:
        pass
    return 1


def child_env(gpu):
    env = dict(os.environ)
    repo = os.path.dirname(HERE)
    extra = [os.path.join(HERE, "dropin"), repo]
    env["PYTHONPATH"] = os.pathsep.join(extra + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    env["RAIN_B200_DEVICE"] = str(gpu)
    env.pop("LOCAL_RANK", None)
    return env


def run_jobs(jobs, n_gpus, jobs_per_gpu=2, main_py="main.py", python=None, cwd=None, poll_s=0.2, log_dir=None):
    """Run every job as ``python main.py <args>``; at most ``jobs_per_gpu`` children per GPU, a job goes to the
    least loaded GPU.  Returns [(args, gpu, returncode)] in completion order."""
    python = python or sys.executable
    log_dir = log_dir or cwd or os.getcwd()
    pending = list(jobs)
    running = []          # (Popen, args, gpu, logfile, errfile)
    done = []
    load = [0] * n_gpus
    while pending or running:
        while pending and min(load) < jobs_per_gpu:
            gpu = load.index(min(load))
            args = pending.pop(0)
            pat = log_pattern(args)
            logf = open(os.path.join(log_dir, "automate_log_" + pat + ".txt"), "a+")
            errf = open(os.path.join(log_dir, "automate_error_" + pat + ".txt"), "a+")
            print(">>> START on GPU %d: %s" % (gpu, " ".join(args)))
            child = subprocess.Popen([python, main_py] + list(args), stdout=logf, stderr=errf, env=child_env(gpu), cwd=cwd)
            running.append((child, args, gpu, logf, errf))
            load[gpu] += 1
        time.sleep(poll_s)
        for item in list(running):
            child, args, gpu, logf, errf = item
            rc = child.poll()
            if rc is None:
                continue
            logf.close(); errf.close()
            running.remove(item)
            load[gpu] -= 1
            done.append((args, gpu, rc))
            print("Job ended (rc %d): %s" % (rc, " ".join(args)))
    return done


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    a = parse(argv)
    jobs = build_jobs(argv)
    n_gpus = a.gpus or visible_gpus()
    print("%d jobs over %d GPU(s), %d per GPU" % (len(jobs), n_gpus, a.jobs_per_gpu))
    print("Note this script does not show real-time output to avoid cumbersome console scrolling. Check ad-hoc logs.")
    done = run_jobs(jobs, n_gpus, a.jobs_per_gpu, a.main)
    failed = [d for d in done if d[2] != 0]
    print("All jobs completed" + (", %d FAILED (see automate_error_*.txt)" % len(failed) if failed else ""))
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
