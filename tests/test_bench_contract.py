"""bench.py's contract that can be checked without a GPU: the reference arm (the CPU algorithm on the host cores)
prints one JSON line with the agreed keys, and the roofline numerator is SURVEY 8(d)'s per-frame byte count."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_per_frame_is_the_survey_formula():
    sys.path.insert(0, ROOT)
    import bench
    W, H, N = 1242, 375, 403
    # SURVEY 8(d) with the round-2 boundary formats: the depth enters as the PNG's uint16 samples (2 bytes, was float32)
    # and the uint8 colormap index of the saved mask is one more output
    assert bench.algorithmic_bytes_per_frame(W, H, N) == 3 * W * H + 2 * W * H + (12 + 4 + 3 + 1) * W * H + 128 * N
    assert bench.REC_BYTES == 128 and bench.BATCH == 64 and bench.WORKLOAD == "C2"


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rainy frames/sec at 1242x375, 25mm/hr" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["steps"] == 1
    assert d["config"]["workload"].startswith("C2") and 300 < d["config"]["streaks_per_frame"] < 500
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "frame" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 0.01 < d["value"] < 100
    # a non-zero rank of a torchrun launch exits 0 without work and without output
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"), cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
