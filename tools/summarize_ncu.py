"""Turns ncu outputs into the small text summaries kept under profiles/.
  python tools/summarize_ncu.py launches <launches.csv> <out.txt> [title]
  python tools/summarize_ncu.py full <report.ncu-rep> <out.txt> [title]"""
import collections
import csv
import subprocess
import sys

WANT = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3}.get(d["Metric Unit"], 1e-6)
        a = agg.setdefault(d["Kernel Name"][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(title + "\n(per-launch times under ncu are cold-cache and serialised: compare SHARES)\n\n")
        f.write("%-92s %6s %12s %7s\n" % ("kernel", "n", "total ms", "share"))
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:25]:
            f.write("%-92s %6d %12.3f %6.1f%%\n" % (k, a[0], a[1], 100 * a[1] / tot))
        f.write("\ntotal device time of all launches: %.3f ms over %d launches\n" % (tot, sum(a[0] for a in agg.values())))


def full(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(title + "\n(per-launch values; source report not tracked)\n\n")
        for r in rows[2:]:
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write("%-80s %s %s\n" % (w, r[i], units[i]))
            f.write("\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    {"launches": launches, "full": full}[mode](src, dst, title)
