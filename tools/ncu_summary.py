"""Condense an `ncu --set full` report into the per-kernel summary CSV kept under profiles/ and the per-stage
DRAM traffic JSON bench.py reads (profiles/roofline_traffic.json).

  python tools/ncu_summary.py REPORT.ncu-rep OUT_SUMMARY.csv [OUT_TRAFFIC.json]
"""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]
STAGE = {"k_stats": "fog", "k_stats_final": "fog", "k_downscale2": "fog", "k_downscale2_final": "fog", "k_fext": "fog", "k_fog": "fog",
         "k_fext_pad": "fog", "k_fog_acs": "fog", "k_fog_roll": "fog",
         "k_env_map": "env", "k_env_prefix": "env", "k_ambient": "env", "k_plan": "setup", "k_setup": "setup", "k_scan": "setup",
         "k_raster": "raster", "k_blur": "blur", "k_composite": "composite", "k_frame_mean": "composite", "k_epilogue": "epilogue"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out_csv = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    traffic = {}
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + METRICS)
        for r in data:
            m_ = re.search(r"\b(k_\w+)", r[idx["Kernel Name"]])          # templates arrive as "void k_fog<(bool)1>(...)"
            name = m_.group(1) if m_ else r[idx["Kernel Name"]].split("(")[0]
            w.writerow([name] + ["%s %s" % (r[idx[m]], units[idx[m]]) if m in idx else "" for m in METRICS])
            b = sum(float(r[idx[m]]) * UNIT.get(units[idx[m]], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            st = STAGE.get(name)
            if st:
                traffic[st] = traffic.get(st, 0.0) + b
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as f:
            json.dump(traffic, f, indent=1)
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
