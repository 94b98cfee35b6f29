#!/usr/bin/env python
"""Summarise an ncu capture (gpurun_out/prof_<tag>.ncu-rep) into profiles/.

usage: python tools/summarise_profile.py <tag> <envs> <emulated instructions per env-step> [--headline]

Writes profiles/<tag>_k_run_frames.csv (selected raw metrics of the captured emulation kernel: k_run_frames_1 for batches of one
env per warp, k_run_frames otherwise) and, with --headline, profiles/kernel_counters.json: the per-launch counters bench.py
quotes in its `roofline.traffic` and `issue` records, stamped with the hash of the CUDA sources so that a stale capture is
never reported for a different kernel.
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]


def main():
    tag, envs, instr_per_env_step = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
    headline = "--headline" in sys.argv
    rep = ROOT / "gpurun_out" / f"prof_{tag}.ncu-rep"
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    kernel = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "k_run_frames"
    out = ROOT / "profiles" / f"{tag}_k_run_frames.csv"
    out.parent.mkdir(exist_ok=True)
    sel = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value", f"# {kernel}, {envs} envs, one launch, ncu --set full --clock-control none"])
        for h, u, v in zip(hdr, units, vals):
            if h in WANT or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                try:
                    if "issue_stalled" in h and float(v) < 0.01:
                        continue
                except ValueError:
                    pass
                w.writerow([h, u, v])
                sel[h] = (u, v)

    def num(name, scale=None):
        u, v = sel[name]
        return float(v) * ({"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u] if scale else 1)

    rec = {"kernel": kernel.split("(")[0], "envs_per_launch": envs, "dram_bytes_per_launch": num("dram__bytes_read.sum", 1) + num("dram__bytes_write.sum", 1),
           "warp_instructions_per_launch": num("smsp__inst_executed.sum"), "emulated_instructions_per_launch": envs * instr_per_env_step,
           "lanes_active": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
           "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
           "sm_active_share_of_elapsed": num("sm__cycles_active.avg") / num("sm__cycles_elapsed.avg"),
           "kernel_ms_under_ncu": num("gpu__time_duration.sum"), "source": f"profiles/{tag}_k_run_frames.csv",
           "note": "one ncu --set full capture (serialised replays, cold caches): counters, not timings, are quoted from it"}
    print(open(out).read())
    print(json.dumps(rec, indent=1))
    import bench

    rec["kernel_source_hash"] = bench.kernel_source_hash()
    json.dump(rec, open(ROOT / "profiles" / ("kernel_counters.json" if headline else f"{tag}_counters.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
