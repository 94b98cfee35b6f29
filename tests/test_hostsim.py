"""The CUDA kernels' device code, compiled for the host (tests/hostsim, -DGB_HOSTSIM), against the CPU oracle.

This is how the interpreter / LCD / bus logic of pokegym_b200/csrc/*.cuh is checked in the `not gpu` tier: the same
headers nvcc compiles into libgbenv.so are compiled by g++ with the CUDA keywords stubbed out (gb_hd.h) and one env at
a time is stepped through run_frames_env -- the body of k_run_frames -- and compared, save-state for save-state, with
the oracle.  The harness is test infrastructure: nothing under pokegym_b200/ can load it.  The GPU tier repeats the same
comparisons on the real kernels (tests/test_gpu_parity.py)."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent / "hostsim"))

from pokegym_b200 import _capi  # noqa: E402
from pokegym_b200.state_file import diff_states  # noqa: E402
from pokegym_b200.tools import synth_rom  # noqa: E402


@pytest.fixture(scope="module")
def hostsim():
    import driver

    driver.build()
    return driver


def _compare(hs, cpu, n, where):
    for e in range(n):
        a, b = hs.save_state(e), cpu.save_state(e)
        assert a == b, f"{where}: env {e}: {diff_states(b, a)[:8]}"
        x = cpu.core_extra(e)
        assert hs.core_extra(e) == (x.stat_mode, x.ly_window, x.fault), where


@pytest.mark.parametrize("simt", [True, False], ids=["simt", "single"])
@pytest.mark.parametrize("rom_name,steps", [("pokelike", 10), ("conformance", 8), ("conformance_b", 6), ("pokelike_timer", 8), ("busy", 4),
                                            ("halt_edge", 30), ("lcd_probe", 10), ("lcd_probe_b", 10), ("divergent", 8)])
def test_device_code_on_host_matches_oracle(hostsim, oracle_lib, roms, rom_name, steps, simt):
    if rom_name not in synth_rom.rom_catalog():
        pytest.skip(f"{rom_name} ROM not in the catalog")
    rom, n = roms(rom_name), 5
    hs, cpu = hostsim.HostSim(n, rom, simt=simt), _capi.Handle(oracle_lib, n, rom)
    _compare(hs, cpu, n, "power-on")
    hs.tick(3, True)
    cpu.tick(3, True)
    _compare(hs, cpu, n, "after 3 rendered frames")
    rng = np.random.default_rng(123)
    for s in range(steps):
        act = rng.integers(0, 8, n).astype(np.uint8)
        hs.run_action(act)
        cpu.run_action(act)
        _compare(hs, cpu, n, f"{rom_name} step {s}")
    c = cpu.counters()
    assert hs.counters()[:2] == [c.instructions, c.cycles]
    hs.close()


@pytest.mark.parametrize("seed", range(40, 46))
def test_lazy_lcd_probe_seeds(hostsim, oracle_lib, seed):
    """More seeds of the LCD-observation program (reads / polls of LY and STAT, STAT / LYC / scroll / LCDC / LY writes,
    HALTs with and without pending interrupts) for the lazy LCD of gb_device.cuh."""
    rom, n = synth_rom.build_lcd_probe_rom(seed=seed, n_blocks=300 + 50 * (seed % 5)), 3
    hs, cpu = hostsim.HostSim(n, rom), _capi.Handle(oracle_lib, n, rom)
    rng = np.random.default_rng(seed)
    for s in range(6):
        act = rng.integers(0, 8, n).astype(np.uint8)
        hs.run_action(act)
        cpu.run_action(act)
        _compare(hs, cpu, n, f"seed {seed} step {s}")
    hs.close()


@pytest.mark.parametrize("defer", [True, False], ids=["deferred", "inline"])
@pytest.mark.parametrize("seed", [7, 8, 9])
def test_deferred_ppu_flushes_before_vram_writes(hostsim, oracle_lib, seed, defer):
    """Deferred PPU (gb_device.cuh render_flush / lcd_record_line, gb_kernels.cuh render_pending_lines): the lines of the frame
    that is rendered are only recorded at their HBlank and drawn after the frame loop; a write to VRAM / OAM (or an OAM DMA)
    while lines are pending must draw them first.  The LCD-observation program writes VRAM, OAM, scroll registers, LCDC and LY in
    the middle of frames, so the framebuffer (part of every save-state compared here) only matches the oracle -- which draws
    each line at its HBlank as PyBoy does -- if the flush rule holds.  Both forms of the renderer must agree with it."""
    import ctypes as C

    rom, n = synth_rom.build_lcd_probe_rom(seed=seed, n_blocks=400), 3
    hs, cpu = hostsim.HostSim(n, rom), _capi.Handle(oracle_lib, n, rom)
    hs.set_defer(defer)
    st = (C.c_ulonglong * 2)()
    hs.dll.hs_flush_stats(st)
    before = st[0]
    hs.tick(4, True)  # every frame rendered: the lines of frame k are flushed at the start of frame k + 1
    cpu.tick(4, True)
    _compare(hs, cpu, n, "after 4 rendered frames")
    rng = np.random.default_rng(seed)
    for s in range(12):
        act = rng.integers(0, 8, n).astype(np.uint8)
        hs.run_action(act)
        cpu.run_action(act)
        _compare(hs, cpu, n, f"seed {seed} step {s}")
    hs.dll.hs_flush_stats(st)
    assert (st[0] > before) == defer  # the flush path ran (deferred) / was never needed (inline)
    hs.close()


def test_every_fast_opcode_has_a_class(hostsim):
    """gb_classes.inc (tools/gen_classes.sh) lists every (handler, operand flags) pair of the per-opcode base table: the
    single-lane build dispatches on the class id, and an opcode without one would always take the slow tick."""
    import ctypes as C

    dll = C.CDLL(str(hostsim.build()))
    n = C.c_int(0)
    assert dll.hs_class_coverage(C.byref(n)) == 0
    assert 30 <= n.value < 255


def test_real_states_resume_identically(hostsim, oracle_lib, roms):
    """Reference save-states (committed under tests/golden) loaded into both and stepped."""
    from helpers import GOLDEN

    rom = roms("pokelike")
    blobs = [np.load(p)["start_state"].tobytes() for p in sorted(GOLDEN.glob("ref_wrapper_red_*.npz"))]
    assert blobs
    n = len(blobs)
    hs, cpu = hostsim.HostSim(n, rom), _capi.Handle(oracle_lib, n, rom)
    for e, b in enumerate(blobs):
        hs.load_blob(e, b)
        cpu.load_template(cpu.add_state_template(b), [e])
    _compare(hs, cpu, n, "after load")
    rng = np.random.default_rng(1)
    for s in range(12):
        act = rng.integers(0, 8, n).astype(np.uint8)
        hs.run_action(act)
        cpu.run_action(act)
        _compare(hs, cpu, n, f"real states step {s}")
    hs.close()
