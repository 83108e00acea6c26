"""Turn ncu artefacts brought back in gpurun_out/ into the text summaries committed
under profiles/ (the .ncu-rep files themselves are scratch).

    python tools/ncu_summary.py rep <file.ncu-rep>      # key metrics of every launch in the report
    python tools/ncu_summary.py launches <launches.csv> # per-kernel totals and shares
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("kernel:", d.get("Kernel Name", ("", "?"))[1])
        for k in KEYS:
            if k in d:
                print(f"  {k:72s} {d[k][1]:>18s} {d[k][0]}")
        stalls = [(float(v[1] or 0), h[len(STALL):-len("_per_issue_active.ratio")]) for h, v in d.items()
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")]
        print("  warp stall reasons (warps stalled per issue-active cycle), top 6:")
        for val, name in sorted(stalls, reverse=True)[:6]:
            print(f"    {name:28s} {val:.3f}")
        print()


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'total us':>12s} {'launches':>8s} {'avg us':>10s} {'share':>7s}  kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:12.1f} {c:8d} {t / c:10.2f} {100 * t / tot:6.2f}%  {k[:110]}")


if __name__ == "__main__":
    {"rep": rep, "launches": launches}[sys.argv[1]](sys.argv[2])
