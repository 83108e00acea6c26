"""Per-source-line instruction counts of an ncu report captured with --import-source on
(`ncu -i rep --page source --print-source cuda,sass --csv`): the lines that executed the most
warp instructions, with their share of the kernel.

    python tools/ncu_lines.py <file.ncu-rep> [top_n] [kernel-name regex]
"""
import csv
import subprocess
import sys


def main(path, top=30, kernel=None):
    cmd = ["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"]
    if kernel:
        cmd += ["--kernel-name", "regex:" + kernel]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr, agg = "", None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            i_inst = hdr.index("Instructions Executed")
            i_thr = hdr.index("Thread Instructions Executed")
            i_smp = hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= i_thr or not r[0].isdigit():
            continue
        try:
            inst, thr, smp = int(r[i_inst]), int(r[i_thr]), int(r[i_smp])
        except ValueError:
            continue
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [0, 0, 0, r[1].strip()])
        a[0] += inst
        a[1] += thr
        a[2] += smp
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    print(f"total warp instructions {tot}, stall samples {tots}")
    print(f"{'warp inst':>12s} {'share':>6s} {'lanes':>6s} {'samples':>8s}  line")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{a[0]:12d} {100 * a[0] / tot:5.1f}% {a[1] / max(1, a[0]):6.1f} {100 * a[2] / tots:7.1f}%  {f}:{ln}: {a[3][:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30, sys.argv[3] if len(sys.argv) > 3 else None)
