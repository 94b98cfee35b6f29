#!/usr/bin/env python
"""Cross-check the CUDA emulator against real PyBoy.

The build image has neither PyBoy (`pyboy<2.0.0`, /root/reference/setup.py:12) nor a Game Boy ROM, so the dynamic
emulator core is "parity unpinned" against PyBoy itself.  A user who has both runs this:

    python tools/pyboy_crosscheck.py --rom pokemon_red.gb --state Bulbasaur.state --steps 200 [--seed 0] [--dump-dir out/]

It drives the same action sequence through
  (a) libgbenv.so (one env, `gbenv_run_action`: press, 24 ticks, release before tick 8, render the last tick), and
  (b) PyBoy, exactly as /root/reference/pokegym/pyboy_binding.py:71-91 does,
saves both emulators with `save_state` after every env-step and reports the first field of the v9 state file
(CPU registers, VRAM, OAM, LCD registers, scanline parameters, framebuffer, WRAM, HRAM, IO, timer, MBC, cart RAM,
joypad) that differs.  Exit status: 0 = identical for all steps, 1 = a difference was found, 3 = not run
(PyBoy, the ROM or a GPU is missing -- reported, never faked).
"""
from __future__ import annotations

import argparse
import io
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

# pyboy_binding.py:7-40 -- action index -> (press, release) WindowEvent names
ACTION_EVENTS = [("PRESS_ARROW_DOWN", "RELEASE_ARROW_DOWN"), ("PRESS_ARROW_LEFT", "RELEASE_ARROW_LEFT"), ("PRESS_ARROW_RIGHT", "RELEASE_ARROW_RIGHT"),
                 ("PRESS_ARROW_UP", "RELEASE_ARROW_UP"), ("PRESS_BUTTON_A", "RELEASE_BUTTON_A"), ("PRESS_BUTTON_B", "RELEASE_BUTTON_B"),
                 ("PRESS_BUTTON_START", "RELEASE_BUTTON_START"), ("PRESS_BUTTON_SELECT", "RELEASE_BUTTON_SELECT")]


def not_run(why: str) -> int:
    print(f"not run: {why}")
    return 3


def pyboy_run_action(pyboy, WindowEvent, action: int, frame_skip: int = 24):
    """pyboy_binding.run_action_on_emulator :71-91 with headless=True, fast_video=True."""
    press, release = (getattr(WindowEvent, n) for n in ACTION_EVENTS[action])
    pyboy.send_input(press)
    pyboy._rendering(False)
    for i in range(frame_skip):
        if i == 8:
            pyboy.send_input(release)
        if i == frame_skip - 1:
            pyboy._rendering(True)
        pyboy.tick()


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--rom", required=True)
    ap.add_argument("--state", required=True, help="PyBoy 1.6 save-state (v9, 142,610 bytes) to start both emulators from")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--frame-skip", type=int, default=24)
    ap.add_argument("--dump-dir", default=None, help="write <step>.gbenv.state / <step>.pyboy.state for the first differing step")
    args = ap.parse_args()

    if not Path(args.rom).exists():
        return not_run(f"ROM {args.rom} missing")
    if not Path(args.state).exists():
        return not_run(f"state {args.state} missing")
    try:
        from pyboy import PyBoy, WindowEvent
    except Exception as e:  # noqa: BLE001
        return not_run(f"PyBoy missing ({e}); install pyboy<2.0.0")
    try:
        import torch

        if not torch.cuda.is_available():
            return not_run("no CUDA device")
    except Exception as e:  # noqa: BLE001
        return not_run(f"torch missing ({e})")

    from pokegym_b200 import _capi
    from pokegym_b200.state_file import diff_states

    rom, blob = Path(args.rom).read_bytes(), Path(args.state).read_bytes()
    h = _capi.Handle(_capi.GbEnvLib(_capi.DEFAULT_LIB), 1, rom)
    h.load_template(h.add_state_template(blob))

    pyboy = PyBoy(args.rom, debugging=False, disable_input=False, window_type="headless", hide_window=True)
    pyboy.set_emulation_speed(0)
    pyboy.load_state(io.BytesIO(blob))

    def pyboy_state() -> bytes:
        f = io.BytesIO()
        pyboy.save_state(f)
        return f.getvalue()

    d0 = diff_states(pyboy_state(), h.save_state(0))
    if d0:
        print("after load_state:", d0)
    actions = torch.randint(0, 8, (args.steps,), generator=torch.Generator().manual_seed(args.seed)).to(torch.uint8)
    a_dev = torch.zeros(1, dtype=torch.uint8, device="cuda")
    for s in range(args.steps):
        a = int(actions[s])
        a_dev.fill_(a)
        h.run_action(a_dev, args.frame_skip)
        pyboy_run_action(pyboy, WindowEvent, a, args.frame_skip)
        mine, ref = h.save_state(0), pyboy_state()
        if mine != ref:
            print(f"step {s} (action {a}): states differ")
            for line in diff_states(ref, mine):
                print("  ", line)
            if args.dump_dir:
                out = Path(args.dump_dir)
                out.mkdir(parents=True, exist_ok=True)
                (out / f"{s}.gbenv.state").write_bytes(mine)
                (out / f"{s}.pyboy.state").write_bytes(ref)
            return 1
    print(f"identical: {args.steps} env-steps x {args.frame_skip} frames, all {len(blob):,} state bytes equal after every step")
    return 0


if __name__ == "__main__":
    sys.exit(main())
