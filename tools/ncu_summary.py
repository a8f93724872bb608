#!/usr/bin/env python
"""Summarise ncu captures from gpurun_out/ into profiles/ (tracked evidence).

    python tools/ncu_summary.py <tag> <prof.ncu-rep> [<launches.csv>] [--traffic <cfg>:<kind><<semantics>>]

Writes profiles/<tag>_key_metrics.csv (one `--set full` capture of the dominant kernel: time, DRAM
bytes, issue utilisation, occupancy, registers, shared memory), profiles/<tag>_details.txt (the
ncu details page) and, when a launch list is given, profiles/<tag>_launches.csv with the kernel
names shortened and each kernel's share of the listed GPU time.
"""
import csv
import io
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_active",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct",
    "launch__registers_per_thread",
    "launch__block_size",
    "launch__grid_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main() -> None:
    args = list(sys.argv[1:])
    traffic_key = None
    if "--traffic" in args:  # record dram__bytes_read + dram__bytes_write of this launch for bench.py's roofline.traffic
        k = args.index("--traffic")
        traffic_key = args[k + 1]
        del args[k:k + 2]
    tag, rep = args[0], args[1]
    launches = args[2] if len(args) > 2 else None
    os.makedirs("profiles", exist_ok=True)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(f"profiles/{tag}_key_metrics.csv", "w") as f:
        f.write(f"kernel,{vals[hdr.index('Kernel Name')]},\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k},{vals[i]},{units[i]}\n")
    if traffic_key:
        import json

        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            tot += float(vals[i]) * scale[units[i]]
        path = "profiles/traffic.json"
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur[traffic_key] = tot
        with open(path, "w") as f:
            json.dump(cur, f, indent=1, sort_keys=True)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    with open(f"profiles/{tag}_details.txt", "w") as f:
        f.write(det)
    if launches:
        lines = [l for l in open(launches) if l.startswith('"')]
        rd = list(csv.DictReader(io.StringIO("".join(lines))))
        tot = sum(float(r["Metric Value"]) for r in rd)
        with open(f"profiles/{tag}_launches.csv", "w") as f:
            f.write("id,kernel,block,grid,time_ns,share_of_listed_gpu_time\n")
            for r in rd:
                name = r["Kernel Name"]
                name = name.split("<")[0].split("(")[0].replace("void ", "")[:60]
                t = float(r["Metric Value"])
                f.write(f"{r['ID']},{name},\"{r['Block Size']}\",\"{r['Grid Size']}\",{t:.0f},{t / tot:.4f}\n")


if __name__ == "__main__":
    main()
