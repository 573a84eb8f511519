#!/usr/bin/env python3
"""Summarise ncu outputs from gpurun_out/ into profiles/ (tracked).
  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.txt
  python tools/ncu_summary.py kernel   gpurun_out/prof_fast_r1.ncu-rep profiles/r1_fast_cells.txt
  python tools/ncu_summary.py traffic  gpurun_out/r2h_fe.ncu-rep gpurun_out/r2h_ba.ncu-rep profiles/r2h_traffic.json
"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    h = rows[0]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] == "ns" else v * 1e3 if r[iu] == "ms" else v
        k = r[ik].split("(")[0]
        agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (source {src}); cold-cache, serialised: compare SHARES\n")
        f.write(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'avg_us':>10s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:70]:70s} {v[0]:8d} {v[1]:12.1f} {v[1] / tot:7.3f} {v[1] / v[0]:10.1f}\n")
    print(open(dst).read())


def launches_batch(src, dst, n):
    """the launch list restricted to the launches over batches of n frames (a grid dimension == n): the device-resident leg of bench.py, whose per-kernel
    shares are what roofline.kernel_share_of_step reports (the whole-command list also holds the e2e leg's 32-frame batches and the single-frame latency calls)"""
    n = str(int(n))
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    h = rows[0]
    ik, ig, iv, iu = h.index("Kernel Name"), h.index("Grid Size"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        dims = [d.strip() for d in r[ig].strip("()").split(",")]
        if n not in dims or not r[ik].startswith(("orbs::", "void orbs::")):
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] == "ns" else v * 1e3 if r[iu] == "ms" else v
        k = r[ik].split("(")[0]
        agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (source {src}); only launches over batches of {n} frames (the device-resident leg); cold-cache, serialised: compare SHARES\n")
        f.write(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'avg_us':>10s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:70]:70s} {v[0]:8d} {v[1]:12.1f} {v[1] / tot:7.3f} {v[1] / v[0]:10.1f}\n")
    print(open(dst).read())


def kernel(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  (source {src})\n")
        for r in rows[2:]:
            f.write(f"## {r[h.index('Kernel Name')][:100]}  (launch id {r[0]})\n")
            for w in WANT:
                if w in h:
                    i = h.index(w)
                    f.write(f"{w:80s} {r[i]:>18s} {units[i]}\n")
    print(open(dst).read())


def traffic(fe_rep, ba_rep, dst, frames_per_launch=128):
    """dram__bytes_read + dram__bytes_write and duration per launch of every captured kernel -> the json bench.py reads for roofline.traffic"""
    import json
    out = {"source": f"ncu --set full --clock-control none ({fe_rep}, {ba_rep}); per launch"}
    for group, rep in (("frontend", fe_rep), ("ba", ba_rep)):
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        h, units = rows[0], rows[1]
        ik, ir, iw, it = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tm = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
        ks = collections.defaultdict(list)
        for r in rows[2:]:
            name = r[ik].split("(")[0].replace("void ", "").replace("orbs::", "")
            name = name.split("<")[0]
            if name.startswith("k_"): name = name[2:]
            f = lambda v: float(v.replace(",", ""))
            ks[name].append({"dram_bytes": int(f(r[ir]) * mult[units[ir]] + f(r[iw]) * mult[units[iw]]), "duration_us": round(f(r[it]) * tm[units[it]], 2)})
        out[group] = {"frames_per_launch": frames_per_launch, "kernels": ks} if group == "frontend" else {"kernels": ks}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({g: {k: v[0] for k, v in out[g]["kernels"].items()} for g in ("frontend", "ba")}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches_batch":
        launches_batch(sys.argv[2], sys.argv[3], sys.argv[4])
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
