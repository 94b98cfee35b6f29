"""Host-side tooling: SM83 assembler, synthetic ROM generator, state-file codec."""
import hashlib

import numpy as np
import pytest

from pokegym_b200.state_file import V9_LEN, diff_states, field_offsets, parse_state, serialize_state
from pokegym_b200.tools import sm83asm, synth_rom


def test_opcode_table_is_complete_and_unique():
    base = {v[0][0] for v in sm83asm.OPTABLE.values() if len(v[0]) == 1}
    cb = {v[0][1] for v in sm83asm.OPTABLE.values() if len(v[0]) == 2}
    assert len(cb) == 256
    assert base == set(range(256)) - set(sm83asm.ILLEGAL_OPCODES) - {0xCB}
    assert len(sm83asm.OPTABLE) == 245 - 1 + 256  # 244 base mnemonics (CB prefix excluded) + the CB page


def test_assembler_labels_and_relative_jumps():
    rom = bytearray(0x8000)
    a = sm83asm.Asm(rom)
    a.org(0x150)
    a.label("top")
    a.i("LD A,n", 0x12)
    a.i("JR NZ,e", "top")
    a.i("JP nn", "top")
    a.link()
    assert bytes(rom[0x150:0x157]) == bytes([0x3E, 0x12, 0x20, 0xFC, 0xC3, 0x50, 0x01])


def test_roms_are_deterministic_and_cover_every_opcode():
    r1, r2 = synth_rom.build_pokelike_rom(), synth_rom.build_pokelike_rom()
    assert r1 == r2 and len(r1) == 1 << 20
    assert r1[0x147] == 0x13 and r1[0x20B3] == 0x76  # MBC3+RAM+BATTERY; HALT where Pokemon Red's DelayFrame halts
    c = synth_rom.build_conformance_rom()
    assert synth_rom.build_conformance_rom.last_missing == []
    assert hashlib.sha256(c).hexdigest() == hashlib.sha256(synth_rom.build_conformance_rom()).hexdigest()


def test_state_codec_round_trip_on_a_generated_state(oracle_lib, roms):
    from pokegym_b200 import _capi

    h = _capi.Handle(oracle_lib, 1, roms("pokelike"))
    h.tick(30, True)
    blob = h.save_state(0)
    assert len(blob) == V9_LEN
    st = parse_state(blob)
    assert serialize_state(st) == blob
    assert st.cpu["PC"] == 0x20B3 and st.cpu["halted"] == 1 and st.lcd["LY"] == 144
    st.raw["wram"][0x123] ^= 0xFF
    d = diff_states(blob, serialize_state(st))
    assert d and d[0].startswith("wram: 1 bytes differ")
    assert field_offsets(9)["joypad"] == (142_608, 2)


def test_unsupported_blobs_are_rejected(oracle_lib, roms):
    from pokegym_b200 import _capi

    h = _capi.Handle(oracle_lib, 1, roms("pokelike"))
    with pytest.raises(_capi.GbEnvError):
        h.add_state_template(b"\x09" + bytes(100))
    with pytest.raises(ValueError):
        parse_state(bytes(10))


def test_oracle_emulator_is_deterministic_and_actions_matter(oracle_lib, roms):
    from pokegym_b200 import _capi

    outs = []
    for seed in (0, 0, 1):
        h = _capi.Handle(oracle_lib, 2, roms("pokelike"))
        h.tick(20, True)
        rng = np.random.default_rng(seed)
        for _ in range(8):
            h.run_action(rng.integers(0, 8, 2).astype(np.uint8))
        outs.append(h.save_state(1))
        assert h.counters().faults == 0
    assert outs[0] == outs[1] and outs[0] != outs[2]


def test_halt_edge_rom_cycles_through_every_scenario(oracle_lib, roms):
    """The HALT/LCD edge-case ROM keeps running on the oracle: every interrupt handler fires, the joypad-only HALT wakes on a button edge, and envs with different actions end in different states."""
    from pokegym_b200 import _capi

    h = _capi.Handle(oracle_lib, 3, roms("halt_edge"))
    rng = np.random.default_rng(1)
    for _ in range(12):
        h.run_action(rng.integers(0, 8, 3).astype(np.uint8))
    # one pass over the ten scenarios per env-step: it then sleeps in scenario 5 until the next step's button press
    vblank, stat, timer, joy = (int(x) for x in h.read_mem(0, 0xFF90, 4))
    assert int(h.read_mem(0, 0xFFA0, 1)[0]) == 5 and joy >= 8 and vblank >= 100 and stat >= 30 and timer >= 12, (vblank, stat, timer, joy)
    assert h.counters().faults == 0
    assert h.save_state(0) != h.save_state(1)


def test_pyboy_crosscheck_driver_matches_run_action(oracle_lib, roms):
    """tools/pyboy_crosscheck.py drives PyBoy the way pyboy_binding.run_action_on_emulator does; on the PyBoy-shaped
    shim over the oracle core that sequence must leave the same state as one gbenv_run_action call."""
    import importlib.util
    from pathlib import Path

    import ref_shim
    from pokegym_b200 import _capi

    spec = importlib.util.spec_from_file_location("pyboy_crosscheck", Path(__file__).resolve().parents[1] / "tools" / "pyboy_crosscheck.py")
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    rom = roms("pokelike")
    ref_shim.ShimConfig.rom, ref_shim.ShimConfig.oracle_lib = rom, oracle_lib
    fake = ref_shim.FakePyBoy("unused.gb")
    h = _capi.Handle(oracle_lib, 1, rom)
    for a in (4, 0, 7, 2, 2, 5):
        tool.pyboy_run_action(fake, ref_shim.WindowEvent, a)
        h.run_action(np.array([a], np.uint8))
        assert fake.handle.save_state(0) == h.save_state(0), diff_states(fake.handle.save_state(0), h.save_state(0))
    assert tool.not_run("x") == 3
