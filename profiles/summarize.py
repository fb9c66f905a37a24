"""Summarise gpurun_out ncu artefacts into small tracked text files under profiles/.

    python profiles/summarize.py launches gpurun_out/r01_launches.csv profiles/r01_launches_summary.txt
    python profiles/summarize.py full     gpurun_out/r01_spmv.ncu-rep  profiles/r01_spmv_full.txt
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__shared_mem_per_block_dynamic",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__issue_active.avg.per_cycle_active",
]


def launches(src, dst):
    per = collections.OrderedDict()
    order = []
    with open(src) as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    for r in rows[1:]:
        name = r[ki].replace("<unnamed>::", "")
        ns = float(r[vi].replace(",", ""))
        per.setdefault(name, []).append(ns)
        order.append((name, ns, r[gi]))
    total = sum(sum(v) for v in per.values())
    with open(dst, "w") as out:
        out.write(f"# {src}: {len(order)} launches, {total / 1e6:.3f} ms total device time (ncu, cold-cache, serialised)\n")
        out.write(f"{'kernel':90s} {'n':>5s} {'avg_us':>10s} {'sum_ms':>10s} {'share':>7s}\n")
        for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            out.write(f"{name[:90]:90s} {len(v):5d} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / 1e6:10.3f} {100 * sum(v) / total:6.1f}%\n")
    print(open(dst).read())


def full(src, dst):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in txt.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as out:
        out.write(f"# {src} (ncu --set full --clock-control none), one block per captured launch\n")
        for r in rows[2:]:
            out.write(f"kernel: {r[hdr.index('Kernel Name')]}\n")
            for k in KEEP:
                if k in hdr:
                    i = hdr.index(k)
                    out.write(f"  {k:85s} {r[i]:>18s} {units[i]}\n")
            out.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
