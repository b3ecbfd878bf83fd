"""Per-kernel hot source lines of an `ncu --set full --import-source on` report (needs -lineinfo builds):

  python tools/ncu_hot_lines.py REPORT.ncu-rep OUT.md [kernel ...]

For each kernel: the source lines (file:line, text) that collect the most warp-stall samples, with their share of the
kernel's samples and of its executed warp instructions, and the average number of active threads."""
import csv
import io
import subprocess
import sys


def lines_of(rep, kernel):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:^" + kernel + "$"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    out, cur_file, hdr, seen_launch = [], None, None, 0
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1]
        elif len(r) >= 2 and r[0] == "Function Name":
            pass
        elif r and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            # the column "Source" appears twice (CUDA text, SASS text): take positions
        elif hdr and len(r) > 10 and r[0] not in ("", "Line No"):
            try:
                out.append((cur_file, int(r[0]), r[1].strip(), int(r[hdr["# Samples"]]), int(r[hdr["Instructions Executed"]]),
                            int(r[hdr["Thread Instructions Executed"]])))
            except (ValueError, KeyError):
                pass
    return out


def main():
    rep, out_md = sys.argv[1], sys.argv[2]
    kernels = sys.argv[3:] or ["k_raster", "k_fog", "k_composite", "k_env_prefix", "k_setup", "k_blur", "k_env_map", "k_plan"]
    with open(out_md, "w") as f:
        f.write("# Hot source lines per kernel (%s)\n\nshare of the kernel's warp-stall samples / of its executed warp instructions / average active threads\n" % rep.split("/")[-1])
        for k in kernels:
            L = lines_of(rep, k)
            if not L:
                continue
            # several launches of the same kernel are concatenated: aggregate by (file, line)
            agg = {}
            for fl, ln, txt, smp, wi, ti in L:
                a = agg.setdefault((fl, ln), [txt, 0, 0, 0])
                a[1] += smp; a[2] += wi; a[3] += ti
            ts = sum(a[1] for a in agg.values()) or 1
            tw = sum(a[2] for a in agg.values()) or 1
            f.write("\n## %s\n\n| samples | instr | threads | where | source |\n|---|---|---|---|---|\n" % k)
            for (fl, ln), a in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
                f.write("| %.1f %% | %.1f %% | %.1f | %s:%d | `%s` |\n" % (100.0 * a[1] / ts, 100.0 * a[2] / tw, a[3] / max(a[2], 1),
                                                                   fl.split("/")[-1], ln, a[0][:110].replace("|", "\\|")))
    print("wrote", out_md)


if __name__ == "__main__":
    main()
