#!/usr/bin/env python
"""Headline benchmark: env-steps/s of the batched Game Boy environment (24 emulated frames + reward +
observation per env-step), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs-per-gpu E] [--impl reference]

* one process per GPU; for N > 1 the driver launches this file under torchrun (RANK / LOCAL_RANK /
  WORLD_SIZE / MASTER_* from the environment) and every rank owns E envs (weak scaling);
  the only collective is the NCCL all-reduce of the 72-double episode-info vector once per 32-step rollout.
* `value`   : whole-job env-steps/s with actions already resident in HBM and observations written into a
              device rollout tensor u8[32, E, 72*80*4]; CUDA-event timed, max over ranks.
* `e2e`     : the same metric through the host-buffer C-ABI call gbenv_step_host (pinned host memory;
              actions H2D and obs/reward/done D2H inside the timed region).
* `roofline`: algorithmic bytes (57,682 B per env-step) of the dominant kernel
              k_run_frames over its CUDA-event duration, against the measured HBM peak.
* `cpu_baseline` / `--impl reference`: the CPU oracle (a port of the reference's algorithm; PyBoy itself
              cannot be installed here) on the box's host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

ALGO_BYTES_PER_ENV_STEP = 57_682  # BASELINE.md section 4 / SURVEY.md 8d
ROLLOUT_T = 32
WORKLOAD = "pokelike synthetic ROM (MBC3, ~25% busy frames), random actions, full reward shaping, obs into device rollout u8[32,E,72,80,4]"


def workload_name(args) -> str:
    from pokegym_b200.tools import synth_rom

    if args.rom == "pokelike" and not args.state:
        return WORKLOAD
    what = f"synthetic ROM '{args.rom}'" if args.rom in synth_rom.rom_catalog() else f"user ROM {Path(args.rom).name}"
    start = f"reset from {Path(args.state).name}" if args.state else "booted for 60 frames"
    return f"{what}, {start}, random actions, full reward shaping, obs into device rollout u8[32,E,72,80,4]"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rom", default=os.environ.get("POKEGYM_ROM", "pokelike"), help="synthetic ROM name or path to a .gb file")
    ap.add_argument("--state", default=None, help="PyBoy .state file every env resets from (default: boot the ROM for 60 frames)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--also-envs", type=int, default=-1,
                    help="extra single-GPU leg at this env count, reported under 'large_batch' (0 = skip; -1 = one full wave: SMs x 20 warps x 32 envs)")
    ap.add_argument("--also-steps", type=int, default=12)
    return ap.parse_args()


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per k_run_frames launch from the committed ncu capture, scaled per env (or None)."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def build_rom(name: str) -> bytes:
    """`name` is a synthetic ROM of pokegym_b200.tools.synth_rom.rom_catalog() or a path to a user-supplied .gb file
    (BASELINE.json config 1: pokemon_red.gb; no ROM ships with the reference, so the default is synthetic)."""
    from pokegym_b200.tools import synth_rom

    if name in synth_rom.rom_catalog():
        fn, kw = synth_rom.rom_catalog()[name]
        return fn(**kw)
    return Path(name).read_bytes()


def start_envs(h, state_path):
    """Every env starts from the same point: the given PyBoy .state file, or the ROM booted for 60 frames."""
    if state_path:
        h.set_initial_template(h.add_state_template(Path(state_path).read_bytes()))
    else:
        h.tick(60, True)


def oracle_env_steps_per_s(rom: bytes, seconds: float, n_envs=None, state_path=None):
    """CPU oracle (port of the reference path) on all host cores; returns (value, cores, sample description)."""
    import numpy as np

    from pokegym_b200 import _capi
    import __graft_entry__ as g

    lib = _capi.GbEnvLib(g.build_oracle(), "oracle_")
    cores = os.cpu_count() or 1
    n = n_envs or max(cores * 4, 8)
    h = _capi.Handle(lib, n, rom)
    start_envs(h, state_path)
    obs = np.zeros((n, _capi.OBS_BYTES), dtype=np.uint8)
    rew = np.zeros(n)
    done = np.zeros(n, dtype=np.uint8)
    h.reset(obs)
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 8, (4096, n)).astype(np.uint8)
    for i in range(2):
        h.step(acts[i], obs, rew, done)
    steps, t0 = 0, time.perf_counter()
    while True:
        h.step(acts[(steps + 2) % 4096], obs, rew, done)
        steps += 1
        dt = time.perf_counter() - t0
        if dt >= seconds or steps >= 4000:
            break
    h.close()
    return n * steps / dt, cores, f"{n} envs x {steps} env-steps of the same ROM / reset state / action distribution in {dt:.1f} s on {cores} host threads"


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  PyBoy (the emulator the reference
    drives) is an un-vendored dependency that cannot be installed here, so this arm times the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rom = build_rom(args.rom)
    cores = os.cpu_count() or 1
    n = max(cores * 4, 8)
    import numpy as np

    from pokegym_b200 import _capi
    import __graft_entry__ as g

    lib = _capi.GbEnvLib(g.build_oracle(), "oracle_")
    h = _capi.Handle(lib, n, rom)
    start_envs(h, args.state)
    obs = np.zeros((n, _capi.OBS_BYTES), dtype=np.uint8)
    rew = np.zeros(n)
    done = np.zeros(n, dtype=np.uint8)
    h.reset(obs)
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 8, (args.warmup + args.steps, n)).astype(np.uint8)
    for i in range(args.warmup):
        h.step(acts[i], obs, rew, done)
    t0 = time.perf_counter()
    for i in range(args.steps):
        h.step(acts[args.warmup + i], obs, rew, done)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args), "envs_per_step": n, "act_freq": 24, "rom": args.rom,
                   "note": "bounded sample: each step advances `envs_per_step` envs on the host cores"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{n} envs x {args.steps} env-steps on {cores} host threads (oracle port; PyBoy + ROM unavailable)"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s": 24 * value,
    }
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g
    from pokegym_b200 import _capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.GbEnvLib(g.build_cuda() if rank == 0 or not _capi.DEFAULT_LIB.exists() else _capi.DEFAULT_LIB)
    E, K, W = args.envs_per_gpu, args.steps, max(args.warmup, 3)
    rom = build_rom(args.rom)
    h = _capi.Handle(lib, E, rom, device_id=local_rank)
    start_envs(h, args.state)  # every env starts from the same state and diverges through its actions
    rollout = torch.zeros((ROLLOUT_T, E, _capi.OBS_BYTES), dtype=torch.uint8, device=dev)
    reward = torch.zeros((ROLLOUT_T, E), dtype=torch.float64, device=dev)
    done = torch.zeros((ROLLOUT_T, E), dtype=torch.uint8, device=dev)
    info_sum = torch.zeros(_capi.INFO_SCALARS, dtype=torch.float64, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    actions = torch.randint(0, 8, (W + K, E), generator=gen, device=dev, dtype=torch.uint8)
    h.reset(rollout[0])

    def step(i):
        t = i % ROLLOUT_T
        h.step(actions[i], rollout[t], reward[t], done[t])
        if t == ROLLOUT_T - 1:  # once per rollout: reduce episode-info scalars across envs and ranks
            h.reduce_info(info_sum)
            if world > 1:
                dist.all_reduce(info_sum)

    for i in range(W):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    c0 = h.counters()
    k0 = h.kernel_time_total(0)
    k1 = h.kernel_time_total(1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for i in range(W, W + K):
        step(i)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    c1 = h.counters()
    k0b = h.kernel_time_total(0)
    k1b = h.kernel_time_total(1)
    if world > 1:
        dist.barrier()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * E * K / (ms / 1000.0)

    # ---- end to end through the host-buffer entry point (pinned memory, copies inside the timed region)
    n_e2e = max(3, min(args.e2e_steps, K))
    act_h = torch.randint(0, 8, (n_e2e + 3, E), dtype=torch.uint8).pin_memory()
    obs_h = torch.zeros((E, _capi.OBS_BYTES), dtype=torch.uint8).pin_memory()
    rew_h = torch.zeros(E, dtype=torch.float64).pin_memory()
    done_h = torch.zeros(E, dtype=torch.uint8).pin_memory()
    for i in range(3):
        h.step_host(act_h[i].numpy(), obs_h.numpy(), rew_h.numpy(), done_h.numpy())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(3, 3 + n_e2e):
        h.step_host(act_h[i].numpy(), obs_h.numpy(), rew_h.numpy(), done_h.numpy())
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * E * n_e2e / e2e_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (k_run_frames), from CUDA events on its launching stream
    run_ms = (k0b[0] - k0[0]) / max(1, k0b[1] - k0[1])
    wrap_ms = (k1b[0] - k1[0]) / max(1, k1b[1] - k1[1])
    peak, peak_src = measured_peak()
    achieved = ALGO_BYTES_PER_ENV_STEP * E / (run_ms / 1000.0) / 1e9
    traffic = recorded_traffic()
    instr = c1.instructions - c0.instructions
    line = {
        "metric": "env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args), "envs_per_gpu": E, "act_freq": 24, "rom": args.rom, "parallelism": f"env-sharded x{world}",
                   "l2_policy": f"working set {E * (16896 + 5760 + 23040 + 1152) / 1e6:.0f} MB of env state + obs per GPU exceeds the 126 MB L2; no explicit flush"},
        "frames_per_s": 24 * value,
        "emulated_instr_per_s": instr / (ms / 1000.0) * world,
        "roofline": {"bound": "hbm", "kernel": "k_run_frames", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (traffic or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                     "kernel_ms": run_ms, "kernel_share_of_step": run_ms / (ms / K), "wrap_kernels_ms": wrap_ms,
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * E},
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": E, "d2h_bytes_per_step": E * (_capi.OBS_BYTES + 8 + 1),
                "steps": n_e2e, "api": "gbenv_step_host (pinned host buffers)"},
        "gpu_launches": int(c1.kernel_launches - c0.kernel_launches),
        "clocks": clocks,
        "faults": int(c1.faults),
    }
    if args.also_envs < 0:
        args.also_envs = torch.cuda.get_device_properties(dev).multi_processor_count * 20 * 32  # 94,720 on a 148-SM B200
    if world == 1 and args.also_envs and args.also_envs != E:
        # the interpreter is issue/latency bound, so throughput keeps growing with the number of resident envs:
        # report the north_star's ">= 32k envs per B200" point next to the headline 4,096-env configuration
        try:
            h.close()
            del rollout
            torch.cuda.empty_cache()
            E2 = args.also_envs
            h2 = _capi.Handle(lib, E2, rom, device_id=local_rank)
            start_envs(h2, args.state)
            ro2 = torch.zeros((4, E2, _capi.OBS_BYTES), dtype=torch.uint8, device=dev)
            rw2 = torch.zeros(E2, dtype=torch.float64, device=dev)
            dn2 = torch.zeros(E2, dtype=torch.uint8, device=dev)
            act2 = torch.randint(0, 8, (3 + args.also_steps, E2), generator=gen, device=dev, dtype=torch.uint8)
            h2.reset(ro2[0])
            for i in range(3):
                h2.step(act2[i], ro2[i % 4], rw2, dn2)
            torch.cuda.synchronize()
            ev0.record()
            for i in range(3, 3 + args.also_steps):
                h2.step(act2[i], ro2[i % 4], rw2, dn2)
            ev1.record()
            torch.cuda.synchronize()
            ms2 = ev0.elapsed_time(ev1)
            line["large_batch"] = {"envs_per_gpu": E2, "value": E2 * args.also_steps / (ms2 / 1000.0), "unit": "env-steps/s", "steps": args.also_steps,
                                   "ms_per_step": ms2 / args.also_steps, "faults": int(h2.counters().faults)}
            h2.close()
        except Exception as e:
            line["large_batch"] = {"error": str(e)}
    try:
        v, cores, sample = oracle_env_steps_per_s(rom, args.cpu_baseline_seconds, state_path=args.state)
        line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample}
    except Exception as e:  # the baseline is reported, never required for the GPU number
        line["cpu_baseline"] = {"value": None, "unit": "env-steps/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
