"""PyBoy save-state (.state) byte layout, reader and writer.

The reference loads these files with ``pyboy.load_state(BytesIO)``
(/root/reference/pokegym/pyboy_binding.py:59-69, environment.py:1241-1242).  PyBoy is not vendored,
so the layout below was verified against all 264 v9 fixtures of the reference tree
(SURVEY.md section 8c): every field has a fixed offset and the parser consumes exactly
142,610 bytes (v9) / 142,586 bytes (v7).

The device-side reset path parses the blob in C++ (csrc/gb_image.h: blob_to_image); this module is the host-side
mirror used by tools and tests (building synthetic states, diffing two states field by field).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

V9_LEN = 142_610
V7_LEN = 142_586
ROWS, COLS = 144, 160

# (name, length) in file order for a v9 DMG state
V9_FIELDS: List[Tuple[str, int]] = [
    ("header", 5),  # version, bootrom_enabled, key1, double_speed, cgb
    ("cpu", 18),  # A F B C D E, HL SP PC (u16 LE), IME halted stopped IE interrupt_queued IF
    ("vram", 0x2000),
    ("oam", 0xA0),
    ("lcd_regs", 11),  # LCDC BGP OBP0 OBP1 STAT LY LYC SCY SCX WY WX
    ("lcd_clock", 19),  # cgb, double_speed, clock u64, clock_target u64, next_stat_mode
    ("scanline_params", 144 * 5),  # SCX SCY WX(raw) WY tile_data_select
    ("screen", ROWS * COLS * 4),  # u32 LE 0xRRGGBBff
    ("wram", 0x2000),
    ("nonio0", 96),  # FEA0-FEFF
    ("io", 76),  # FF00-FF4B
    ("hram", 127),  # FF80-FFFE
    ("nonio1", 52),  # FF4C-FF7F
    ("timer", 8),  # DIV TIMA DIV_counter(u16) TIMA_counter(u16) TMA TAC
    ("mbc", 4),  # rombank rambank ram_enabled memorymodel
    ("cart_ram", 0x8000),
    ("joypad", 2),  # directional, standard
]


def field_offsets(version: int = 9) -> Dict[str, Tuple[int, int]]:
    """name -> (offset, length) for the given state version."""
    out: Dict[str, Tuple[int, int]] = {}
    off = 0
    for name, ln in V9_FIELDS:
        if version == 7:
            if name == "header":
                ln = 2
            elif name == "cpu":
                ln = 16
            elif name == "lcd_clock":
                ln = 0
        out[name] = (off, ln)
        off += ln
    return out


assert sum(ln for _, ln in V9_FIELDS) == V9_LEN
assert field_offsets(7)["joypad"][0] + 2 == V7_LEN


@dataclass
class GBState:
    """All fields of a DMG save-state as numpy byte arrays / ints."""

    version: int = 9
    raw: Dict[str, np.ndarray] = field(default_factory=dict)

    # -- convenience views -------------------------------------------------
    @property
    def cpu(self) -> Dict[str, int]:
        c = self.raw["cpu"].tobytes()
        A, F, B, C, D, E = c[0:6]
        HL, SP, PC = struct.unpack("<HHH", c[6:12])
        d = dict(A=A, F=F, B=B, C=C, D=D, E=E, HL=HL, SP=SP, PC=PC, IME=c[12], halted=c[13], stopped=c[14], IE=c[15])
        if self.version >= 8:
            d.update(interrupt_queued=c[16], IF=c[17])
        return d

    @property
    def lcd(self) -> Dict[str, int]:
        names = "LCDC BGP OBP0 OBP1 STAT LY LYC SCY SCX WY WX".split()
        d = dict(zip(names, self.raw["lcd_regs"].tolist()))
        if self.version >= 8:
            cgb, ds, clock, target, nsm = struct.unpack("<BBQQB", self.raw["lcd_clock"].tobytes())
            d.update(clock=clock, clock_target=target, next_stat_mode=nsm)
        return d

    @property
    def screen_rgb(self) -> np.ndarray:
        """What ``screen_ndarray()`` returns: (144, 160, 3) uint8, flag byte dropped."""
        return self.raw["screen"].reshape(ROWS, COLS, 4)[:, :, 1:]

    def mem(self, addr: int) -> int:
        """Value ``get_memory_value(addr)`` would return for RAM-backed addresses."""
        if 0x8000 <= addr < 0xA000:
            return int(self.raw["vram"][addr - 0x8000])
        if 0xC000 <= addr < 0xE000:
            return int(self.raw["wram"][addr - 0xC000])
        if 0xE000 <= addr < 0xFE00:
            return int(self.raw["wram"][addr - 0xE000])
        if 0xFE00 <= addr < 0xFEA0:
            return int(self.raw["oam"][addr - 0xFE00])
        if 0xFF80 <= addr < 0xFFFF:
            return int(self.raw["hram"][addr - 0xFF80])
        raise ValueError(f"address {addr:#06x} is not plain RAM in a state file")


def parse_state(blob: bytes) -> GBState:
    if len(blob) == V9_LEN and blob[0] == 9:
        version = 9
    elif len(blob) == V7_LEN and blob[0] == 7:
        version = 7
    else:
        raise ValueError(f"unsupported PyBoy state: version byte {blob[0] if blob else None}, {len(blob)} bytes")
    st = GBState(version=version)
    arr = np.frombuffer(blob, dtype=np.uint8)
    for name, (off, ln) in field_offsets(version).items():
        st.raw[name] = arr[off : off + ln].copy()
    return st


def serialize_state(st: GBState) -> bytes:
    """Inverse of :func:`parse_state` (byte-exact round trip)."""
    parts = []
    for name, (_, ln) in field_offsets(st.version).items():
        a = np.asarray(st.raw[name], dtype=np.uint8).reshape(-1)
        if a.size != ln:
            raise ValueError(f"field {name}: expected {ln} bytes, got {a.size}")
        parts.append(a.tobytes())
    return b"".join(parts)


def diff_states(a: bytes, b: bytes, version: int = 9) -> List[str]:
    """Human-readable list of differing fields between two equally-versioned blobs."""
    out = []
    for name, (off, ln) in field_offsets(version).items():
        xa = np.frombuffer(a[off : off + ln], dtype=np.uint8)
        xb = np.frombuffer(b[off : off + ln], dtype=np.uint8)
        if not np.array_equal(xa, xb):
            idx = np.nonzero(xa != xb)[0]
            first = ", ".join(f"+{i}:{xa[i]:02x}!={xb[i]:02x}" for i in idx[:6])
            out.append(f"{name}: {idx.size} bytes differ ({first})")
    return out
