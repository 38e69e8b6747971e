#!/usr/bin/env python
"""Turn ncu outputs (brought back in gpurun_out/) into the small text/JSON
summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r01.csv   > profiles/r01_launches.txt
  python profiles/summarize.py full     gpurun_out/prof_r01.ncu-rep   > profiles/r01_ncu_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.avg.per_second",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        ns = float(r[v].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[u], 1)
        a = agg.setdefault(r[k], [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("# %d launches, %.3f ms total device time" % (sum(a[0] for a in agg.values()), tot / 1e6))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%6.2f%%  n=%5d  mean=%10.1f us  total=%10.3f ms  %s" % (100 * a[1] / tot, a[0], a[1] / a[0] / 1e3, a[1] / 1e6, name))


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none: selected raw metrics per profiled launch")
    for r in rows[2:]:
        print("== %s  grid %s block %s" % (r[idx["Kernel Name"]], r[idx.get("Grid Size", 0)], r[idx.get("Block Size", 0)]))
        for m in KEEP:
            if m in idx:
                print("   %-85s %s %s" % (m, r[idx[m]], units[idx[m]]))


def traffic(path):
    """-> JSON fragment for profiles/traffic.json: DRAM bytes and warp instructions per profiled launch, keyed by kernel name"""
    import json
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = collections.OrderedDict()
    for r in rows[2:]:
        def val(m):
            return float(r[idx[m]].replace(",", "")) * scale.get(units[idx[m]], 1)
        out.setdefault(r[idx["Kernel Name"]], {"dram_bytes": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
                                              "warp_inst": int(val("smsp__inst_executed.sum")),
                                              "us": round(float(r[idx["gpu__time_duration.sum"]]) * {"us": 1, "ms": 1e3, "ns": 1e-3}[units[idx["gpu__time_duration.sum"]]], 1)})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
