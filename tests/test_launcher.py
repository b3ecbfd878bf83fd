"""rain_rendering_b200.launcher: the reference's main_threaded.py job split (main_threaded.py:98-174) and
log naming (:23-32), with the children spread over GPUs.  Host logic only."""
import os
import sys

from rain_rendering_b200 import launcher as L


def test_one_job_per_intensity():
    jobs = L.build_jobs(["--dataset", "kitti", "--intensity", "1,25,100", "-s", "data_object/training", "--gpus", "4"])
    assert len(jobs) == 3
    for j, r in zip(jobs, (1, 25, 100)):
        assert j[j.index("--intensity") + 1] == str(r)
        assert j[-3:] == ["--conflict_strategy", "skip", "--noverbose"]         # main_threaded.py:164-167
        assert "--gpus" not in j and j[:2] == ["--dataset", "kitti"]


def test_scene_threaded_41_frame_ranges():
    argv = ["--dataset", "kitti", "--intensity", "5,50", "--scene_threaded", "--frame_start", "10", "--frame_end", "100", "-v",
            "--scenes_per_thread", "25", "--jobs_per_gpu", "3"]
    jobs = L.build_jobs(argv)
    # ranges [10,51) [51,92) [92,100) x two intensities, range-major like the reference's loops (:113-114); the
    # reference's five identical copies per pair (scene loop without --sequences) collapse to one
    got = [(j[j.index("--frame_start") + 1], j[j.index("--frame_end") + 1], j[j.index("--intensity") + 1]) for j in jobs]
    assert got == [("10", "51", "5"), ("10", "51", "50"), ("51", "92", "5"), ("51", "92", "50"), ("92", "100", "5"), ("92", "100", "50")]
    for j in jobs:
        assert "--scene_threaded" not in j and "-v" not in j and "--scenes_per_thread" not in j and "--jobs_per_gpu" not in j
        assert j.count("--noverbose") == 1 and j[j.index("--conflict_strategy") + 1] == "skip"
    assert L.log_pattern(jobs[2]) == "5mm_51_to_92"
    assert L.log_pattern(L.build_jobs(["--intensity", "7"])[0]) == "7mm_0_to_NA"
    assert L.log_pattern(["--intensity", "7", "--frame_start", "2", "--frame_end", "9", "--frame_step", "3"]) == "7mm_2_to_9_step_3"
    # a defaulted --frame_start is appended instead of raising like the reference's list.index
    j = L.build_jobs(["--intensity", "1", "--scene_threaded", "--frame_end", "50"])
    assert [(x[x.index("--frame_start") + 1], x[x.index("--frame_end") + 1]) for x in j] == [("0", "41"), ("41", "50")]


def test_jobs_are_spread_over_gpus_with_the_dropin_on_the_path(tmp_path):
    fake = tmp_path / "main.py"
    fake.write_text(
        "import os, sys, time, json\n"
        "time.sleep(0.3)\n"
        "print(json.dumps({'gpu': os.environ['RAIN_B200_DEVICE'], 'pp': os.environ['PYTHONPATH'].split(os.pathsep)[:2], 'args': sys.argv[1:]}))\n"
        "sys.exit(3 if '13' in sys.argv else 0)\n")
    jobs = L.build_jobs(["--dataset", "customdb", "--intensity", "1,5,13,25,50,100"])
    done = L.run_jobs(jobs, n_gpus=2, jobs_per_gpu=1, main_py=str(fake), cwd=str(tmp_path), poll_s=0.05)
    assert len(done) == 6 and sorted(d[1] for d in done) == [0, 0, 0, 1, 1, 1]
    assert [d[2] for d in done if "13" in d[0]] == [3] and sum(d[2] != 0 for d in done) == 1
    import json
    for args, gpu, rc in done:
        log = tmp_path / ("automate_log_%s.txt" % L.log_pattern(args))
        rec = json.loads(log.read_text().strip().splitlines()[-1])
        assert rec["gpu"] == str(gpu) and rec["args"] == args
        assert rec["pp"][0].endswith(os.path.join("rain_rendering_b200", "dropin")) and os.path.isdir(os.path.join(rec["pp"][1], "oracle"))
        assert (tmp_path / ("automate_error_%s.txt" % L.log_pattern(args))).exists()


import pytest


@pytest.mark.gpu
def test_launcher_really_renders_two_concurrent_jobs_on_the_gpu(tmp_path):
    """Two jobs of the main_threaded.py split (two frame ranges of one intensity) as two concurrent child processes on GPU 0,
    each driving the drop-in Generator over its own range -- through the launcher's own process management, environment
    (drop-in first on PYTHONPATH, RAIN_B200_DEVICE) and log files.  The reference's main.py is not on the GPU box, so the
    children run a stand-in that does what main.py:226-231 does with an argument object built for a synthetic tree."""
    import cv2
    import numpy as np
    from rain_rendering_b200 import synth
    root = str(tmp_path)
    W, H, nf = 320, 192, 6
    paths = synth.write_dataset(root, "customdb", "seq1", W, H, nf, 25, 500, seed=8, n_sim_frames=2)
    tools = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")
    fake = tmp_path / "main.py"
    fake.write_text(
        "import sys, json, argparse\n"
        "sys.path.insert(0, %r)\n"
        "import dropin_e2e\n"
        "p = argparse.ArgumentParser()\n"
        "p.add_argument('--intensity', type=int); p.add_argument('--frame_start', type=int, default=0); p.add_argument('--frame_end', type=int)\n"
        "p.add_argument('--conflict_strategy'); p.add_argument('--noverbose', action='store_true'); p.add_argument('--dataset')\n"
        "r = p.parse_args()\n"
        "from common.generator import Generator          # resolves to the drop-in: the launcher put it first on PYTHONPATH\n"
        "a = dropin_e2e.make_args(json.load(open(%r)), 'customdb', r.intensity)\n"
        "a.conflict_strategy, a.frame_start, a.frame_end, a.verbose = r.conflict_strategy, r.frame_start, r.frame_end, False\n"
        "g = Generator(a); g.run(); print('RENDERED', g.last_stats['frames'])\n" % (tools, str(tmp_path / "paths.json")))
    import json
    json.dump(paths, open(str(tmp_path / "paths.json"), "w"))
    os.environ["RAIN_B200_BATCH"] = "2"
    try:
        jobs = L.build_jobs(["--dataset", "customdb", "--intensity", "25", "--scene_threaded", "--frame_start", "0", "--frame_end", str(nf)])
        assert len(jobs) == 1                                    # 6 frames < 41: one range; split it by hand into two ranges
        jobs = [L._set(L._set(jobs[0], "--frame_start", 0), "--frame_end", 3), L._set(L._set(jobs[0], "--frame_start", 3), "--frame_end", nf)]
        done = L.run_jobs(jobs, n_gpus=1, jobs_per_gpu=2, main_py=str(fake), cwd=str(tmp_path), poll_s=0.1)
    finally:
        del os.environ["RAIN_B200_BATCH"]
    assert [d[2] for d in done] == [0, 0], [(tmp_path / ("automate_error_%s.txt" % L.log_pattern(j))).read_text()[-1500:] for j in jobs]
    for j in jobs:
        assert "RENDERED 3" in (tmp_path / ("automate_log_%s.txt" % L.log_pattern(j))).read_text()
    out_dir = os.path.join(paths["output"], "customdb", "seq1", "rain", "25mm")
    imgs = [cv2.imread(os.path.join(out_dir, "rainy_image", "%06d.png" % i)) for i in range(nf)]
    assert all(im is not None and im.shape == (H, W, 3) for im in imgs)
    assert all(os.path.exists(os.path.join(out_dir, "rain_mask", "%06d.png" % i)) for i in range(nf))
    assert all(np.abs(im.astype(int) - cv2.imread(os.path.join(paths["dataset_root"], "customdb", "seq1", "rgb", "%06d.png" % i)).astype(int)).mean() > 0.1
               for i, im in enumerate(imgs))                     # really rendered: rain and fog changed the frames
