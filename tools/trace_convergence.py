#!/usr/bin/env python
"""Convergence study for k_run_frames (CPU only; test infrastructure, not product code).

The device headers are compiled for the host with -DGB_SLOT_TRACE (tests/hostsim): every instruction an env executes is
recorded as (kind, physical ROM address, handler id, descriptor flags), with a marker wherever the lanes of a warp
re-converge (one iteration of the frame loop of run_frames_env).  From the traces of L envs -- what one warp with L lanes
would carry -- this script evaluates, interval by interval:

  * `all`   : the kernel's policy -- every lane executes its next instruction each slot.  Reported: slots per emulated
              instruction of one lane, distinct handler ids per slot, distinct instruction addresses per slot, share of
              slots with a memory access.
  * `minpc` : only the lanes at the lowest physical address execute (classic SIMT re-convergence heuristic).
  * `vote`  : the largest group of lanes at one address executes.

usage: python tools/trace_convergence.py [--rom pokelike] [--lanes 8] [--preroll 60] [--steps 2] [--mixed]
"""
import argparse
import ctypes as C
import subprocess
import sys
from collections import Counter
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "hostsim"))


def build():
    out = Path("/tmp/libhostsim_trace.so")
    src = ROOT / "tests" / "hostsim" / "hostsim.cpp"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DGB_SLOT_TRACE", "-Wno-unknown-pragmas", "-o", str(out), str(src)], check=True)
    return out


def split(trace):
    """one env's records -> list of intervals, each an array of records (markers removed)"""
    kind = trace >> 48
    cuts = np.nonzero(kind == 2)[0]
    out = []
    for a, b in zip(cuts, list(cuts[1:]) + [len(trace)]):
        out.append(trace[a + 1:b])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rom", default="pokelike")
    ap.add_argument("--lanes", type=int, default=8)
    ap.add_argument("--preroll", type=int, default=60)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--mixed", action="store_true")
    args = ap.parse_args()
    import driver
    from pokegym_b200.tools import synth_rom

    fn, kw = synth_rom.rom_catalog()[args.rom]
    rom = fn(**kw)
    driver.build = build  # the tracing build
    L = args.lanes
    hs = driver.HostSim(L, rom)
    hs.dll.hs_slots.restype = C.c_size_t
    hs.dll.hs_slots.argtypes = [C.c_void_p, C.c_size_t]
    if args.mixed:
        blobs = np.load(ROOT / "tests" / "golden" / "red_states_mixed.npz")["states"]
        for e in range(L):
            hs.load_blob(e, blobs[e % len(blobs)].tobytes())
    else:
        hs.tick(60, True)
    rng = np.random.default_rng(5)
    buf = np.empty(1, dtype=np.uint64)
    for _ in range(args.preroll):
        hs.run_action(rng.integers(0, 8, L).astype(np.uint8))
        hs.dll.hs_slots(buf.ctypes.data, 0)  # drop
    tot = Counter()
    hist_handlers, hist_addrs = Counter(), Counter()
    for _ in range(args.steps):
        hs.run_action(rng.integers(0, 8, L).astype(np.uint8))  # hs_run steps env 0 through its 24 frames, then env 1, ...
        n = hs.dll.hs_slots(None, 0)
        rec = np.empty(n, dtype=np.uint64)
        hs.dll.hs_slots(rec.ctypes.data, n)
        per_env = split_by_env(rec, L)
        nint = min(len(x) for x in per_env)
        for k in range(nint):
            seqs = [x[k] for x in per_env]
            lens = [len(s) for s in seqs]
            mx = max(lens)
            if mx == 0:
                continue
            tot["instr"] += sum(lens)
            tot["slots_all"] += mx
            for t in range(mx):
                live = [s[t] for s in seqs if len(s) > t]
                hs_ = {int(r >> np.uint64(24)) & 0xFF if int(r >> np.uint64(48)) == 0 else 99 for r in live}
                ad = {int(r) & 0xFFFFFF for r in live}
                hist_handlers[len(hs_)] += 1
                hist_addrs[len(ad)] += 1
                tot["lanes_live"] += len(live)
                if any((int(r >> np.uint64(32)) & 0x220) for r in live):
                    tot["slots_mem"] += 1
            # min-pc and vote policies
            for pol in ("minpc", "vote"):
                pos = [0] * L
                slots = 0
                grp = 0
                while True:
                    cur = [(int(seqs[i][pos[i]]) & 0xFFFFFF, i) for i in range(L) if pos[i] < lens[i]]
                    if not cur:
                        break
                    if pol == "minpc":
                        a = min(c[0] for c in cur)
                    else:
                        a = Counter(c[0] for c in cur).most_common(1)[0][0]
                    g = [i for (x, i) in cur if x == a]
                    for i in g:
                        pos[i] += 1
                    slots += 1
                    grp += len(g)
                tot["slots_" + pol] += slots
                tot["grp_" + pol] += grp
    per_lane = tot["instr"] / L
    print(f"rom {args.rom} lanes {L} mixed {args.mixed}: {tot['instr']} instructions, {per_lane:.0f} per lane")
    print(f"  all  : slots/lane-instr {tot['slots_all'] / per_lane:.3f}  lanes live/slot {tot['lanes_live'] / tot['slots_all']:.2f}  slots with memory access {tot['slots_mem'] / tot['slots_all']:.2f}")
    n = sum(hist_handlers.values())
    print("         distinct handlers per slot:", {k: round(v / n, 3) for k, v in sorted(hist_handlers.items())}, "mean", round(sum(k * v for k, v in hist_handlers.items()) / n, 2))
    print("         distinct addresses per slot:", {k: round(v / n, 3) for k, v in sorted(hist_addrs.items())}, "mean", round(sum(k * v for k, v in hist_addrs.items()) / n, 2))
    for pol in ("minpc", "vote"):
        print(f"  {pol:5s}: slots/lane-instr {tot['slots_' + pol] / per_lane:.3f}  lanes active/slot {tot['grp_' + pol] / tot['slots_' + pol]:.2f}")


def split_by_env(rec, L):
    """hs_run traces env 0's 24 frames, then env 1's, ...; a frame-loop marker (kind 2) opens every interval.  The traces
    are separated by running each env's instruction stream until its share of the frames is done: the harness marks
    the start of an env with kind 3."""
    kind = (rec >> np.uint64(48)).astype(np.int64)
    starts = list(np.nonzero(kind == 3)[0])
    assert len(starts) == L, (len(starts), L)
    out = []
    for a, b in zip(starts, starts[1:] + [len(rec)]):
        out.append(split(rec[a + 1:b]))
    return out


if __name__ == "__main__":
    main()
