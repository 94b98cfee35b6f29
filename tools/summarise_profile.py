#!/usr/bin/env python
"""Summarise an ncu capture (gpurun_out/prof_<tag>.ncu-rep) + launch list into profiles/.

usage: python tools/summarise_profile.py <tag> [envs] [lanes]
Writes profiles/<tag>_k_run_frames.csv (selected raw metrics), profiles/<tag>_launches.csv (copy of the
launch list) and profiles/traffic.json (DRAM bytes per k_run_frames launch, read by bench.py).
"""
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]


def main():
    tag = sys.argv[1]
    envs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    lanes = sys.argv[3] if len(sys.argv) > 3 else "default"
    rep = ROOT / "gpurun_out" / f"prof_{tag}.ncu-rep"
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = ROOT / "profiles" / f"{tag}_k_run_frames.csv"
    out.parent.mkdir(exist_ok=True)
    sel = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value", f"# k_run_frames, {envs} envs, lanes={lanes}, ncu --set full --clock-control none"])
        for h, u, v in zip(hdr, units, vals):
            if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                try:
                    if "issue_stalled" in h and float(v) < 0.01:
                        continue
                except ValueError:
                    pass
                w.writerow([h, u, v])
                sel[h] = (u, v)

    def nbytes(name):
        u, v = sel[name]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

    traffic = {"dram_bytes_per_launch": nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"), "envs_per_launch": envs,
               "kernel": "k_run_frames", "source": f"profiles/{tag}_k_run_frames.csv", "note": "one ncu --set full capture (cold caches, serialised replays)"}
    json.dump(traffic, open(ROOT / "profiles" / "traffic.json", "w"), indent=1)
    ll = ROOT / "gpurun_out" / f"launches_{tag}.csv"
    if ll.exists():
        shutil.copy(ll, ROOT / "profiles" / f"{tag}_launches.csv")
    print(open(out).read())
    print(traffic)


if __name__ == "__main__":
    main()
