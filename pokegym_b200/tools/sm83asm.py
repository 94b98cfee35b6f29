"""A small SM83 (Game Boy CPU) assembler used to author the synthetic test / bench ROMs.

No ROM ships with the reference (``*.gb`` is git-ignored, /root/reference/.gitignore:1-2) and none
can be fetched, so every ROM this project runs is generated here.  The mnemonic table is built from
the regular x/y/z structure of the opcode map; operands are ``n`` (imm8), ``nn`` (imm16 or label),
``e`` (relative label or signed offset).

    a = Asm(org=0x150)
    a.label("loop")
    a.i("LD A,n", 0x12)
    a.i("JR NZ,e", "loop")
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

R8 = ["B", "C", "D", "E", "H", "L", "(HL)", "A"]
RP = ["BC", "DE", "HL", "SP"]
RP2 = ["BC", "DE", "HL", "AF"]
CC = ["NZ", "Z", "NC", "C"]
ALU = ["ADD A,", "ADC A,", "SUB ", "SBC A,", "AND ", "XOR ", "OR ", "CP "]
ROT = ["RLC", "RRC", "RL", "RR", "SLA", "SRA", "SWAP", "SRL"]


def _build_table() -> Dict[str, Tuple[bytes, str]]:
    """mnemonic -> (opcode bytes, operand kind in {'', 'n', 'nn', 'e'})"""
    t: Dict[str, Tuple[bytes, str]] = {}

    def put(m: str, op: int, kind: str = "", prefix: bool = False):
        t[m] = (bytes([0xCB, op]) if prefix else bytes([op]), kind)

    put("NOP", 0x00)
    put("LD (nn),SP", 0x08, "nn")
    put("STOP", 0x10, "n")
    put("JR e", 0x18, "e")
    for i, c in enumerate(CC):
        put(f"JR {c},e", 0x20 + 8 * i, "e")
    for p, rp in enumerate(RP):
        put(f"LD {rp},nn", 0x01 + 16 * p, "nn")
        put(f"ADD HL,{rp}", 0x09 + 16 * p)
        put(f"INC {rp}", 0x03 + 16 * p)
        put(f"DEC {rp}", 0x0B + 16 * p)
    put("LD (BC),A", 0x02)
    put("LD (DE),A", 0x12)
    put("LD (HL+),A", 0x22)
    put("LD (HL-),A", 0x32)
    put("LD A,(BC)", 0x0A)
    put("LD A,(DE)", 0x1A)
    put("LD A,(HL+)", 0x2A)
    put("LD A,(HL-)", 0x3A)
    for y, r in enumerate(R8):
        put(f"INC {r}", 0x04 + 8 * y)
        put(f"DEC {r}", 0x05 + 8 * y)
        put(f"LD {r},n", 0x06 + 8 * y, "n")
    for y, m in enumerate(["RLCA", "RRCA", "RLA", "RRA", "DAA", "CPL", "SCF", "CCF"]):
        put(m, 0x07 + 8 * y)
    for y, d in enumerate(R8):
        for z, s in enumerate(R8):
            if y == 6 and z == 6:
                continue
            put(f"LD {d},{s}", 0x40 + 8 * y + z)
    put("HALT", 0x76)
    for y, m in enumerate(ALU):
        for z, s in enumerate(R8):
            put(f"{m}{s}", 0x80 + 8 * y + z)
        put(f"{m}n", 0xC6 + 8 * y, "n")
    for i, c in enumerate(CC):
        put(f"RET {c}", 0xC0 + 8 * i)
        put(f"JP {c},nn", 0xC2 + 8 * i, "nn")
        put(f"CALL {c},nn", 0xC4 + 8 * i, "nn")
    put("LDH (n),A", 0xE0, "n")
    put("LDH A,(n)", 0xF0, "n")
    put("ADD SP,e", 0xE8, "n")
    put("LD HL,SP+e", 0xF8, "n")
    for p, rp in enumerate(RP2):
        put(f"POP {rp}", 0xC1 + 16 * p)
        put(f"PUSH {rp}", 0xC5 + 16 * p)
    put("RET", 0xC9)
    put("RETI", 0xD9)
    put("JP HL", 0xE9)
    put("LD SP,HL", 0xF9)
    put("LD (C),A", 0xE2)
    put("LD (nn),A", 0xEA, "nn")
    put("LD A,(C)", 0xF2)
    put("LD A,(nn)", 0xFA, "nn")
    put("JP nn", 0xC3, "nn")
    put("DI", 0xF3)
    put("EI", 0xFB)
    put("CALL nn", 0xCD, "nn")
    for y in range(8):
        put(f"RST {y * 8:02X}", 0xC7 + 8 * y)
    for y, m in enumerate(ROT):
        for z, r in enumerate(R8):
            put(f"{m} {r}", 8 * y + z, prefix=True)
    for x, m in ((1, "BIT"), (2, "RES"), (3, "SET")):
        for y in range(8):
            for z, r in enumerate(R8):
                put(f"{m} {y},{r}", 64 * x + 8 * y + z, prefix=True)
    return t


OPTABLE = _build_table()
ILLEGAL_OPCODES = (0xD3, 0xDB, 0xDD, 0xE3, 0xE4, 0xEB, 0xEC, 0xED, 0xF4, 0xFC, 0xFD)

Operand = Union[int, str, None]


class Asm:
    """Assembles into a flat ROM image; ``org`` addresses are CPU addresses within ``bank``."""

    def __init__(self, rom: bytearray):
        self.rom = rom
        self.labels: Dict[str, int] = {}
        self.fixups: List[Tuple[int, str, str, int]] = []  # (rom offset, label, kind, cpu addr of next instr)
        self.bank = 0
        self.pc = 0

    # -- positioning ------------------------------------------------------
    def org(self, addr: int, bank: Optional[int] = None):
        if bank is not None:
            self.bank = bank
        elif addr < 0x4000:
            self.bank = 0
        self.pc = addr

    def _off(self, addr: Optional[int] = None) -> int:
        a = self.pc if addr is None else addr
        if a < 0x4000:
            return a
        return self.bank * 0x4000 + (a - 0x4000)

    def label(self, name: str):
        if name in self.labels:
            raise ValueError(f"duplicate label {name}")
        self.labels[name] = self.pc

    def here(self) -> int:
        return self.pc

    # -- emission ---------------------------------------------------------
    def db(self, *vals: int):
        for v in vals:
            self.rom[self._off()] = v & 0xFF
            self.pc += 1

    def dw(self, v: Union[int, str]):
        if isinstance(v, str):
            self.fixups.append((self._off(), v, "nn", 0))
            self.db(0, 0)
        else:
            self.db(v & 0xFF, (v >> 8) & 0xFF)

    def i(self, mnemonic: str, operand: Operand = None):
        try:
            opc, kind = OPTABLE[mnemonic]
        except KeyError:
            raise ValueError(f"unknown mnemonic {mnemonic!r}") from None
        self.db(*opc)
        if kind == "":
            if operand is not None:
                raise ValueError(f"{mnemonic} takes no operand")
        elif kind == "n":
            if not isinstance(operand, int):
                raise ValueError(f"{mnemonic} needs an int operand")
            self.db(operand & 0xFF)
        elif kind == "nn":
            self.dw(operand)  # type: ignore[arg-type]
        elif kind == "e":
            if isinstance(operand, str):
                self.fixups.append((self._off(), operand, "e", self.pc + 1))
                self.db(0)
            else:
                self.db(int(operand) & 0xFF)  # type: ignore[arg-type]

    def link(self):
        for off, name, kind, nxt in self.fixups:
            if name not in self.labels:
                raise ValueError(f"undefined label {name}")
            target = self.labels[name]
            if kind == "nn":
                self.rom[off] = target & 0xFF
                self.rom[off + 1] = target >> 8
            else:
                d = target - nxt
                if not -128 <= d <= 127:
                    raise ValueError(f"JR to {name} out of range ({d})")
                self.rom[off] = d & 0xFF
        self.fixups.clear()


def finalize_header(rom: bytearray, title: str = "SYNTH", cart_type: int = 0x13, rom_size_code: int = 0x05, ram_size_code: int = 0x03):
    """Fill the cartridge header (title, MBC3+RAM+BATTERY, 1 MiB / 32 KiB) and its checksums."""
    t = title.encode("ascii")[:15]
    rom[0x134 : 0x134 + 16] = t + bytes(16 - len(t))
    rom[0x147] = cart_type
    rom[0x148] = rom_size_code
    rom[0x149] = ram_size_code
    rom[0x14A] = 0x01
    rom[0x14B] = 0x33
    chk = 0
    for a in range(0x134, 0x14D):
        chk = (chk - rom[a] - 1) & 0xFF
    rom[0x14D] = chk
    rom[0x14E] = rom[0x14F] = 0
    s = sum(rom) & 0xFFFF
    rom[0x14E] = s >> 8
    rom[0x14F] = s & 0xFF
