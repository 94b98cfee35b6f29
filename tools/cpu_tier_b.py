#!/usr/bin/env python
"""BASELINE.md tier B: the reference's UNMODIFIED pokegym.Environment (loaded from /root/reference, third-party imports stubbed,
PyBoy replaced by a shim over the CPU oracle's emulator core -- tests/ref_shim.py), one process per host core, random actions.

Runs only where /root/reference is mounted (the build container; the GPU box has no reference tree), so the result is written
to profiles/<tag>_cpu_tier_b.json and bench.py quotes that file, saying where it was measured.  This is NOT a PyBoy number.

usage: python tools/cpu_tier_b.py [--seconds 8] [--procs N] [--tag r2]
"""
import argparse
import contextlib
import io
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def worker(rank, seconds, q):
    import numpy as np

    import __graft_entry__ as g
    import ref_shim
    from pokegym_b200 import _capi
    from pokegym_b200.tools import synth_rom

    lib = _capi.GbEnvLib(g.build_oracle(), "oracle_")
    rom = synth_rom.build_pokelike_rom()
    h = _capi.Handle(lib, 1, rom)
    h.tick(60, True)
    with tempfile.NamedTemporaryFile(suffix=".state", delete=False) as f:
        f.write(h.save_state(0))
        path = f.name
    env, _ = ref_shim.make_reference_env(rom, lib, path)
    rng = np.random.default_rng(rank)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        env.reset()
        for _ in range(20):
            env.step(int(rng.integers(0, 8)))
        n, crashes, t0 = 0, 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            try:
                env.step(int(rng.integers(0, 8)))
            except Exception:  # the reference raises on some game states of the synthetic ROM (e.g. IndexError in update_heat_map
                crashes += 1   # for coordinates outside its 444 x 436 map): a vectoriser would restart the env
                env.reset()
            n += 1
        dt = time.perf_counter() - t0
    os.unlink(path)
    q.put((n, dt, crashes))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=8.0)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--tag", default="r2")
    args = ap.parse_args()
    if not Path("/root/reference/pokegym/environment.py").exists():
        raise SystemExit("tier B needs the reference tree at /root/reference")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=worker, args=(r, args.seconds, q)) for r in range(args.procs)]
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    value = sum(n / dt for n, dt, _ in res)
    rec = {"tier": "B", "value": value, "unit": "env-steps/s", "cores": args.procs, "kind": "reference wrapper + oracle-CPU core (not PyBoy)",
           "sample": f"{args.procs} processes x {args.seconds:.0f} s, {sum(n for n, _, _ in res)} env-steps ({sum(c for _, _, c in res)} ended by an exception of the reference and restarted), pokelike synthetic ROM booted 60 frames, random actions",
           "measured_on": "build container host (no GPU box run: /root/reference is not mounted there)"}
    out = ROOT / "profiles" / f"{args.tag}_cpu_tier_b.json"
    out.write_text(json.dumps(rec, indent=1))
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
