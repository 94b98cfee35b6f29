"""Deterministic synthetic Game Boy ROMs (1 MiB, MBC3+RAM+BATTERY, 64 banks).

Pokemon Red cannot ship (BASELINE.json config 1: "user-supplied pokemon_red.gb") and no ROM exists in
any environment we control, so parity tests, smoke() and bench.py run on ROMs generated here
(SURVEY.md section 8d "Synthetic ROM").  ``build_pokelike_rom`` imitates the structure of Pokemon Red's
overworld loop so that the workload has the same shape:

* ``DelayFrame`` sits at 0x20AF with the ``HALT`` byte at 0x20B3, the address where 202/264 reference
  save-states are parked, so real fixtures resume cleanly; every unused byte is 0xFF (``RST 38``)
  and RST 38 recovers into the main loop, so the 62 mid-instruction fixtures also land safely.
* the VBlank handler does what Pokemon's does: scroll registers from HRAM shadows, a 120-byte
  tile-map row transfer, OAM DMA through the HRAM stub, DIV-driven RNG, joypad read.
* per frame: an "audio engine" in bank 2, a sprite/OAM builder in bank 1, and game logic in bank 3
  that walks the player, opens text boxes / menus (window on/off), starts battles, changes maps with
  an LCD-off tileset reload, saves to cart RAM -- and mutates the *same WRAM addresses Pokemon Red
  uses* for position, party, HP, badges, events, bag, Pokedex, moves, so the
  reward path sees changing inputs.

``build_conformance_rom`` is a seeded random-program generator covering every legal opcode.
"""
from __future__ import annotations

import random
from typing import Dict, Tuple

from .sm83asm import Asm, finalize_header, OPTABLE, R8

ROM_SIZE = 1 << 20

# HRAM variables (offsets for LDH)
H_DMA = 0x80
H_SCX, H_SCY, H_WY = 0xAE, 0xAF, 0xB0
H_JOYLAST, H_JOYRELEASED, H_JOYPRESSED, H_JOYHELD = 0xB1, 0xB2, 0xB3, 0xB4
H_BANK = 0xB8
H_WALK, H_DIR, H_TEXT, H_BATTLE = 0xC0, 0xC1, 0xC2, 0xC3
H_T0, H_T1, H_T2, H_T3 = 0xC4, 0xC5, 0xC6, 0xC7
H_MENU, H_ROW, H_BUSY = 0xC8, 0xC9, 0xCA
H_RADD, H_RSUB, H_FRAME, H_VBL = 0xD3, 0xD4, 0xD5, 0xD6
H_JOYINPUT = 0xF8

W_OAM = 0xC300
W_TILEMAP = 0xC3A0

MAP_TABLE = [0, 1, 2, 12, 13, 33, 40, 41, 54, 0x6C, 88, 59, 60, 61, 3, 4, 5, 6, 14, 15, 20, 21, 36, 51, 52, 134, 0xC2, 92, 7, 8, 9, 10]
ITEM_TABLE = [0x04, 0x14, 0xC4, 0x48, 0x06, 0x3E, 0x4A, 0x33, 0x0B, 0x10, 0xC5, 0xC6, 0x13, 0x1D, 0x28, 0x49]


def _xorshift_bytes(seed: int, n: int) -> bytes:
    rng = random.Random(seed)
    return bytes(rng.getrandbits(8) for _ in range(n))


def build_pokelike_rom(busy_iters: int = 40, seed: int = 1234, timer: bool = False, always_busy: bool = False) -> bytes:
    """Pokemon-like workload ROM.  ``busy_iters`` tunes the per-frame filler loop; ``timer`` enables
    TAC=0x05 with a TIMA handler; ``always_busy`` replaces HALT in DelayFrame by a spin (100 % busy)."""
    rom = bytearray([0xFF]) * ROM_SIZE
    a = Asm(rom)

    def ldh_w(off: int):  # LDH (off),A
        a.i("LDH (n),A", off)

    def ldh_r(off: int):  # LDH A,(off)
        a.i("LDH A,(n)", off)

    def st(addr: int):  # LD (addr),A
        a.i("LD (nn),A", addr)

    def ld(addr: int):  # LD A,(addr)
        a.i("LD A,(nn)", addr)

    def set_a(v: int):
        a.i("LD A,n", v)

    # ---------------------------------------------------------------- vectors
    for v in range(0, 0x40, 8):
        a.org(v)
        a.i("JP nn", "Recover")
    a.org(0x40)
    a.i("JP nn", "VBlank")
    a.org(0x48)
    a.i("RETI")
    a.org(0x50)
    a.i("JP nn", "TimerInt")
    a.org(0x58)
    a.i("RETI")
    a.org(0x60)
    a.i("RETI")
    a.org(0x100)
    a.i("NOP")
    a.i("JP nn", "Start")

    # ------------------------------------------------------------- bank 0 code
    a.org(0x150)
    a.label("Start")
    a.i("DI")
    a.i("LD SP,nn", 0xDFFF)
    a.i("CALL nn", "InitHardware")
    a.i("CALL nn", "InitGameRAM")
    a.i("JP nn", "EnterMain")

    a.label("Recover")  # entered from RST sled / fixtures: keep game RAM, re-arm the machine
    a.i("DI")
    a.i("LD SP,nn", 0xDFFF)
    a.i("CALL nn", "InitHardwareSoft")
    a.label("EnterMain")
    a.i("XOR A")
    a.i("LDH (n),A", 0x0F)
    a.i("EI")
    a.label("MainLoop")
    a.i("CALL nn", "DelayFrame")
    a.i("CALL nn", "JoypadEdges")
    set_a(2)
    a.i("CALL nn", "BankSwitch")
    a.i("CALL nn", 0x4000)  # AudioEngine (bank 2)
    set_a(1)
    a.i("CALL nn", "BankSwitch")
    a.i("CALL nn", 0x4000)  # SpriteUpdate (bank 1)
    set_a(3)
    a.i("CALL nn", "BankSwitch")
    a.i("CALL nn", 0x4000)  # GameLogic (bank 3)
    a.i("CALL nn", "BusyWork")
    a.i("JR e", "MainLoop")

    a.label("BankSwitch")  # A = bank
    ldh_w(H_BANK)
    st(0x2000)
    a.i("RET")

    a.label("Random")  # Pokemon's Random_: mixes DIV into hRandomAdd/hRandomSub; returns A
    a.i("PUSH BC")
    a.i("LDH A,(n)", 0x04)
    a.i("LD B,A")
    ldh_r(H_RADD)
    a.i("ADC A,B")
    ldh_w(H_RADD)
    a.i("LDH A,(n)", 0x04)
    a.i("LD B,A")
    ldh_r(H_RSUB)
    a.i("SBC A,B")
    ldh_w(H_RSUB)
    a.i("LD B,A")
    ldh_r(H_RADD)
    a.i("XOR B")
    a.i("RLCA")
    a.i("POP BC")
    a.i("RET")

    a.label("JoypadEdges")  # hJoyHeld / hJoyPressed from hJoyInput, like Pokemon's Joypad
    ldh_r(H_JOYINPUT)
    a.i("LD B,A")
    ldh_r(H_JOYLAST)
    a.i("LD E,A")
    a.i("XOR B")
    a.i("LD D,A")
    a.i("AND E")
    ldh_w(H_JOYRELEASED)
    a.i("LD A,D")
    a.i("AND B")
    ldh_w(H_JOYPRESSED)
    a.i("LD A,B")
    ldh_w(H_JOYLAST)
    ldh_w(H_JOYHELD)
    a.i("RET")

    a.label("ReadJoypad")  # Pokemon's ReadJoypad (home/joypad.asm): 6 + 10 settle reads
    set_a(1 << 5)
    a.i("LD C,n", 0)
    a.i("LDH (n),A", 0x00)
    for _ in range(6):
        a.i("LDH A,(n)", 0x00)
    a.i("CPL")
    a.i("AND n", 0x0F)
    a.i("SWAP A")
    a.i("LD B,A")
    set_a(1 << 4)
    a.i("LDH (n),A", 0x00)
    for _ in range(10):
        a.i("LDH A,(n)", 0x00)
    a.i("CPL")
    a.i("AND n", 0x0F)
    a.i("OR B")
    ldh_w(H_JOYINPUT)
    set_a(0x30)
    a.i("LDH (n),A", 0x00)
    a.i("RET")

    a.label("BusyWork")  # filler: checksum over wTileMap, busy_iters x 8 bytes
    ldh_r(H_BUSY)
    a.i("OR A")
    a.i("RET Z")
    a.i("LD B,A")
    a.i("LD HL,nn", W_TILEMAP)
    a.i("XOR A")
    a.label("bw_loop")
    for _ in range(8):
        a.i("ADD A,(HL)")
        a.i("RLCA")
        a.i("INC L")
    a.i("DEC B")
    a.i("JR NZ,e", "bw_loop")
    st(0xC0F0)
    a.i("RET")

    a.label("FarCopyData")  # A = bank, HL src, DE dst, BC count (Pokemon's FarCopyData)
    ldh_w(H_T3)
    ldh_r(H_BANK)
    a.i("PUSH AF")
    ldh_r(H_T3)
    a.i("CALL nn", "BankSwitch")
    a.i("CALL nn", "CopyData")
    a.i("POP AF")
    a.i("JP nn", "BankSwitch")

    a.label("FarReadByte")  # A = bank, HL address -> A
    ldh_w(H_T3)
    ldh_r(H_BANK)
    a.i("PUSH AF")
    ldh_r(H_T3)
    a.i("CALL nn", "BankSwitch")
    a.i("LD A,(HL)")
    ldh_w(H_T3)
    a.i("POP AF")
    a.i("CALL nn", "BankSwitch")
    ldh_r(H_T3)
    a.i("RET")

    a.label("DisableLCD")  # Pokemon's DisableLCD: mask the VBlank IRQ, wait for LY==145, clear LCDC bit 7
    a.i("LDH A,(n)", 0x40)
    a.i("BIT 7,A")
    a.i("RET Z")
    a.i("XOR A")
    a.i("LDH (n),A", 0x0F)
    a.i("LDH A,(n)", 0xFF)
    a.i("LD B,A")
    a.i("RES 0,A")
    a.i("LDH (n),A", 0xFF)
    a.label("dl_wait")
    a.i("LDH A,(n)", 0x44)
    a.i("CP n", 145)
    a.i("JR NZ,e", "dl_wait")
    a.i("LDH A,(n)", 0x40)
    a.i("RES 7,A")
    a.i("LDH (n),A", 0x40)
    a.i("LD A,B")
    a.i("LDH (n),A", 0xFF)
    a.i("RET")

    a.label("EnableLCD")
    a.i("LDH A,(n)", 0x40)
    a.i("SET 7,A")
    a.i("LDH (n),A", 0x40)
    a.i("RET")

    a.label("CopyData")  # HL src, DE dst, BC count (Pokemon's CopyData)
    a.i("LD A,(HL+)")
    a.i("LD (DE),A")
    a.i("INC DE")
    a.i("DEC BC")
    a.i("LD A,C")
    a.i("OR B")
    a.i("JR NZ,e", "CopyData")
    a.i("RET")

    a.label("InitHardware")
    a.i("CALL nn", "DisableLCD")
    # tile data: 6 KiB from bank 4
    set_a(4)
    a.i("CALL nn", "BankSwitch")
    a.i("LD HL,nn", 0x4000)
    a.i("LD DE,nn", 0x8000)
    a.i("LD BC,nn", 0x1800)
    a.i("CALL nn", "CopyData")
    a.i("LD HL,nn", 0x5800)
    a.i("LD DE,nn", 0x9800)
    a.i("LD BC,nn", 0x0800)
    a.i("CALL nn", "CopyData")
    a.label("InitHardwareSoft")
    # DMA stub -> HRAM
    a.i("LD HL,nn", "DMAStub")
    a.i("LD DE,nn", 0xFF80)
    a.i("LD BC,nn", 10)
    a.i("CALL nn", "CopyData")
    a.i("XOR A")
    ldh_w(H_SCX)
    ldh_w(H_SCY)
    ldh_w(H_WALK)
    ldh_w(H_TEXT)
    ldh_w(H_BATTLE)
    ldh_w(H_MENU)
    ldh_w(H_ROW)
    ldh_w(H_JOYLAST)
    a.i("LDH (n),A", 0x42)
    a.i("LDH (n),A", 0x43)
    a.i("LDH (n),A", 0x41)
    set_a(144)
    ldh_w(H_WY)
    a.i("LDH (n),A", 0x4A)
    set_a(7)
    a.i("LDH (n),A", 0x4B)
    set_a(0xE4)
    a.i("LDH (n),A", 0x47)
    set_a(0xD0)
    a.i("LDH (n),A", 0x48)
    set_a(0xE0)
    a.i("LDH (n),A", 0x49)
    set_a(busy_iters & 0xFF)
    ldh_w(H_BUSY)
    if timer:
        set_a(0x40)
        a.i("LDH (n),A", 0x06)  # TMA
        set_a(0x05)
        a.i("LDH (n),A", 0x07)  # TAC: enabled, /16
    else:
        a.i("XOR A")
        a.i("LDH (n),A", 0x07)
    set_a(0x0D)
    a.i("LDH (n),A", 0xFF)  # IE = VBlank | Timer | Serial (Pokemon's value)
    set_a(0xE3)
    a.i("LDH (n),A", 0x40)
    a.i("RET")

    a.label("DMAStub")
    a.db(0x3E, 0xC3, 0xE0, 0x46, 0x3E, 0x28, 0x3D, 0x20, 0xFD, 0xC9)

    a.label("InitGameRAM")
    # clear wOAMBuffer + sprite data + audio structs
    a.i("LD HL,nn", 0xC000)
    a.i("LD BC,nn", 0x0400)
    a.label("ig_clr")
    a.i("XOR A")
    a.i("LD (HL+),A")
    a.i("DEC BC")
    a.i("LD A,C")
    a.i("OR B")
    a.i("JR NZ,e", "ig_clr")
    # tile map <- bank 4 map data
    a.i("LD HL,nn", 0x5800)
    a.i("LD DE,nn", W_TILEMAP)
    a.i("LD BC,nn", 360)
    a.i("CALL nn", "CopyData")
    # party: 1 mon, species 0x99, level 6, HP 23/23, moves 0x21 0x2D 0 0
    set_a(1)
    st(0xD163)
    set_a(0x99)
    st(0xD164)
    st(0xD16B)
    set_a(0xFF)
    st(0xD165)
    a.i("XOR A")
    st(0xD16C)
    st(0xD18D)
    set_a(23)
    st(0xD16D)
    st(0xD18E)
    set_a(6)
    st(0xD18C)
    set_a(0x21)
    st(0xD173)
    set_a(0x2D)
    st(0xD174)
    # position: Oak's lab (map 40), X=5, Y=3; money 3000 BCD
    set_a(40)
    st(0xD35E)
    set_a(3)
    st(0xD361)
    set_a(5)
    st(0xD362)
    set_a(0x30)
    st(0xD348)
    set_a(0xFF)
    st(0xD31E)  # empty bag terminator
    a.i("RET")

    a.label("TimerInt")
    a.i("PUSH AF")
    ld(0xC0F1)
    a.i("INC A")
    st(0xC0F1)
    a.i("POP AF")
    a.i("RETI")

    # --------------------------------------------------------- VBlank handler
    a.label("VBlank")
    a.i("PUSH AF")
    a.i("PUSH BC")
    a.i("PUSH DE")
    a.i("PUSH HL")
    ldh_r(H_SCX)
    a.i("LDH (n),A", 0x43)
    ldh_r(H_SCY)
    a.i("LDH (n),A", 0x42)
    ldh_r(H_WY)
    a.i("LDH (n),A", 0x4A)
    # AutoBgMapTransfer: 6 rows x 20 bytes of wTileMap -> 0x9800 map, rotating third
    ldh_r(H_ROW)
    a.i("INC A")
    a.i("CP n", 3)
    a.i("JR C,e", "vb_rowok")
    a.i("XOR A")
    a.label("vb_rowok")
    ldh_w(H_ROW)
    a.i("LD HL,nn", W_TILEMAP)
    a.i("LD DE,nn", 0x9800)
    a.i("OR A")
    a.i("JR Z,e", "vb_rows")
    a.i("LD HL,nn", W_TILEMAP + 120)
    a.i("LD DE,nn", 0x9800 + 6 * 32)
    a.i("DEC A")
    a.i("JR Z,e", "vb_rows")
    a.i("LD HL,nn", W_TILEMAP + 240)
    a.i("LD DE,nn", 0x9800 + 12 * 32)
    a.label("vb_rows")
    a.i("LD B,n", 6)
    a.label("vb_row")
    for _ in range(20):
        a.i("LD A,(HL+)")
        a.i("LD (DE),A")
        a.i("INC E")
    a.i("LD A,E")
    a.i("ADD A,n", 12)
    a.i("LD E,A")
    a.i("JR NC,e", "vb_noc")
    a.i("INC D")
    a.label("vb_noc")
    a.i("DEC B")
    a.i("JR NZ,e", "vb_row")
    a.i("CALL nn", 0xFF80)  # OAM DMA
    # Random (Pokemon's VBlank RNG step)
    a.i("LDH A,(n)", 0x04)
    a.i("LD B,A")
    ldh_r(H_RADD)
    a.i("ADC A,B")
    ldh_w(H_RADD)
    a.i("LDH A,(n)", 0x04)
    a.i("LD B,A")
    ldh_r(H_RSUB)
    a.i("SBC A,B")
    ldh_w(H_RSUB)
    a.i("XOR A")
    ldh_w(H_VBL)
    ldh_r(H_FRAME)
    a.i("INC A")
    ldh_w(H_FRAME)
    a.i("CALL nn", "ReadJoypad")
    a.i("POP HL")
    a.i("POP DE")
    a.i("POP BC")
    a.i("POP AF")
    a.i("RETI")

    # ------------------------------------------------ DelayFrame at 0x20AF
    assert a.here() < 0x20AF, hex(a.here())
    a.org(0x20AF)
    a.label("DelayFrame")
    set_a(1)  # 20AF
    ldh_w(H_VBL)  # 20B1
    a.label("df_halt")
    assert a.here() == 0x20B3
    if always_busy:
        a.i("NOP")
    else:
        a.i("HALT")  # 20B3
    ldh_r(H_VBL)  # 20B4
    a.i("AND A")
    a.i("JR NZ,e", "df_halt")
    a.i("RET")

    # --------------------------------------------------- bank 2: audio engine
    a.org(0x4000, bank=2)
    a.i("LD HL,nn", 0xC000)
    a.i("LD C,n", 8)
    a.label("au_ch")
    a.i("DEC (HL)")  # note delay counter
    a.i("JR NZ,e", "au_next")
    # reload: pointer (HL+1,HL+2) into bank-2 song data
    a.i("PUSH HL")
    a.i("INC L")
    a.i("LD A,(HL+)")
    a.i("LD E,A")
    a.i("LD A,(HL)")
    a.i("AND n", 0x0F)
    a.i("OR n", 0x60)
    a.i("LD D,A")
    a.i("LD A,(DE)")  # next note
    a.i("INC DE")
    a.i("LD B,A")
    a.i("LD A,D")
    a.i("LD (HL-),A")
    a.i("LD A,E")
    a.i("LD (HL-),A")
    a.i("LD A,B")
    a.i("AND n", 0x0F)
    a.i("INC A")
    a.i("LD (HL),A")  # new delay
    a.i("LD A,B")
    a.i("SWAP A")
    a.i("AND n", 0x0F)
    a.i("ADD A,A")
    a.i("LD E,A")
    a.i("LD D,n", 0x68)
    a.i("LD A,(DE)")  # frequency table (bank 2, 0x6800)
    a.i("LDH (n),A", 0x13)
    a.i("INC E")
    a.i("LD A,(DE)")
    a.i("LDH (n),A", 0x14)
    set_a(0x77)
    a.i("LDH (n),A", 0x24)
    a.i("POP HL")
    a.label("au_next")
    # vibrato / envelope bookkeeping bytes
    a.i("PUSH HL")
    a.i("LD A,L")
    a.i("ADD A,n", 4)
    a.i("LD L,A")
    a.i("LD A,(HL)")
    a.i("ADD A,n", 3)
    a.i("LD (HL+),A")
    a.i("RRA")
    a.i("XOR (HL)")
    a.i("LD (HL+),A")
    a.i("AND n", 0x3F)
    a.i("LD (HL),A")
    a.i("POP HL")
    a.i("LD A,L")
    a.i("ADD A,n", 16)
    a.i("LD L,A")
    a.i("DEC C")
    a.i("JR NZ,e", "au_ch")
    a.i("RET")
    # song data + frequency table
    data = _xorshift_bytes(seed + 2, 0x1000)
    rom[2 * 0x4000 + 0x2000 : 2 * 0x4000 + 0x3000] = data
    rom[2 * 0x4000 + 0x2800 : 2 * 0x4000 + 0x2820] = _xorshift_bytes(seed + 3, 0x20)

    # ------------------------------------------------- bank 1: sprite update
    # 16 sprite structs at C100 + 16*i: [0]=picture id [1]=movement [2]=image idx [4]=Y [6]=X [7]=anim [9]=facing
    a.org(0x4000, bank=1)
    a.i("LD HL,nn", 0xC100)
    a.i("LD DE,nn", W_OAM)
    a.i("LD B,n", 16)
    a.label("sp_loop")
    a.i("LD A,(HL)")
    a.i("OR A")
    a.i("JR Z,e", "sp_hide")
    a.i("PUSH HL")
    a.i("INC L")
    a.i("INC L")
    a.i("LD C,(HL)")  # image index
    a.i("INC L")
    a.i("INC L")
    a.i("LD A,(HL)")  # Y
    a.i("INC L")
    a.i("INC L")
    a.i("LD H,(HL)")  # X
    a.i("LD L,A")
    # four OAM entries (2x2 tiles)
    for dy, dx, dt in ((0, 0, 0), (0, 8, 1), (8, 0, 2), (8, 8, 3)):
        a.i("LD A,L")
        a.i("ADD A,n", 16 + dy)
        a.i("LD (DE),A")
        a.i("INC E")
        a.i("LD A,H")
        a.i("ADD A,n", 8 + dx)
        a.i("LD (DE),A")
        a.i("INC E")
        a.i("LD A,C")
        a.i("ADD A,n", dt)
        a.i("LD (DE),A")
        a.i("INC E")
        a.i("LD A,C")
        a.i("AND n", 0xB0)  # attributes derived from image idx: palette / flips / priority
        a.i("LD (DE),A")
        a.i("INC E")
    a.i("POP HL")
    a.i("JR e", "sp_next")
    a.label("sp_hide")
    a.i("LD A,E")
    a.i("CP n", 0xA0)
    a.i("JR NC,e", "sp_next")
    a.i("LD C,n", 4)
    a.label("sp_hide1")
    set_a(160)
    a.i("LD (DE),A")
    a.i("INC E")
    a.i("INC E")
    a.i("INC E")
    a.i("INC E")
    a.i("DEC C")
    a.i("JR NZ,e", "sp_hide1")
    a.label("sp_next")
    a.i("LD A,L")
    a.i("AND n", 0xF0)
    a.i("ADD A,n", 16)
    a.i("LD L,A")
    a.i("LD A,E")
    a.i("CP n", 0xA0)
    a.i("JR NC,e", "sp_done")
    a.i("DEC B")
    a.i("JP NZ,nn", "sp_loop")
    a.label("sp_done")
    # NPC wander: every 32 frames nudge one NPC
    ldh_r(H_FRAME)
    a.i("AND n", 0x1F)
    a.i("RET NZ")
    a.i("CALL nn", "Random")
    a.i("LD C,A")
    a.i("AND n", 0x70)
    a.i("ADD A,n", 0x10)
    a.i("LD L,A")
    a.i("LD H,n", 0xC1)
    a.i("LD A,(HL)")
    a.i("OR A")
    a.i("RET Z")
    a.i("LD A,L")
    a.i("ADD A,n", 4)
    a.i("BIT 0,C")
    a.i("JR Z,e", "sp_wy")
    a.i("ADD A,n", 2)
    a.label("sp_wy")
    a.i("LD L,A")
    a.i("LD A,C")
    a.i("AND n", 0x08)
    a.i("SUB n", 4)
    a.i("ADD A,(HL)")
    a.i("AND n", 0x7F)
    a.i("LD (HL),A")
    a.i("RET")

    # --------------------------------------------------- bank 3: game logic
    a.org(0x4000, bank=3)
    a.label("GameLogic")
    ldh_r(H_BATTLE)
    a.i("OR A")
    a.i("JP NZ,nn", "BattleFrame")
    ldh_r(H_TEXT)
    a.i("OR A")
    a.i("JP NZ,nn", "TextFrame")
    ldh_r(H_WALK)
    a.i("OR A")
    a.i("JP NZ,nn", "WalkFrame")
    ldh_r(H_MENU)
    a.i("OR A")
    a.i("JP NZ,nn", "MenuFrame")
    # idle: dispatch on joypad
    ldh_r(H_JOYHELD)
    a.i("LD B,A")
    a.i("BIT 7,B")
    a.i("LD C,n", 0x00)
    a.i("JR NZ,e", "gl_walk")
    a.i("BIT 6,B")
    a.i("LD C,n", 0x04)
    a.i("JR NZ,e", "gl_walk")
    a.i("BIT 5,B")
    a.i("LD C,n", 0x08)
    a.i("JR NZ,e", "gl_walk")
    a.i("BIT 4,B")
    a.i("LD C,n", 0x0C)
    a.i("JR NZ,e", "gl_walk")
    ldh_r(H_JOYPRESSED)
    a.i("LD B,A")
    a.i("BIT 0,B")
    a.i("JP NZ,nn", "Interact")
    a.i("BIT 3,B")
    a.i("JP NZ,nn", "MenuOpen")
    a.i("BIT 2,B")
    a.i("JP NZ,nn", "SaveSRAM")
    a.i("RET")

    a.label("gl_walk")  # C = facing
    a.i("LD A,C")
    st(0xC109)
    ldh_w(H_DIR)
    # collision: map block byte from bank (map & 31) + 8 at 0x4000 + ((Y << 8 | X) & 0x3FFF)
    ld(0xD361)
    a.i("AND n", 0x3F)
    a.i("OR n", 0x40)
    a.i("LD H,A")
    ld(0xD362)
    a.i("LD L,A")
    ld(0xD35E)
    a.i("AND n", 0x1F)
    a.i("ADD A,n", 8)
    a.i("CALL nn", "FarReadByte")
    st(0xCFC6)  # wTileInFrontOfPlayer
    a.i("ADD A,C")  # direction-dependent, so a tile never blocks all four ways
    a.i("AND n", 0x07)
    a.i("RET Z")  # blocked: 1 in 8 (tile, direction) pairs
    set_a(8)
    ldh_w(H_WALK)
    a.i("RET")

    a.label("WalkFrame")
    a.i("DEC A")
    ldh_w(H_WALK)
    a.i("PUSH AF")
    # smooth scroll 2 px / frame
    ldh_r(H_DIR)
    a.i("LD C,A")
    a.i("OR A")
    a.i("JR NZ,e", "wf_1")
    ldh_r(H_SCY)
    a.i("ADD A,n", 2)
    ldh_w(H_SCY)
    a.i("JR e", "wf_anim")
    a.label("wf_1")
    a.i("CP n", 4)
    a.i("JR NZ,e", "wf_2")
    ldh_r(H_SCY)
    a.i("SUB n", 2)
    ldh_w(H_SCY)
    a.i("JR e", "wf_anim")
    a.label("wf_2")
    a.i("CP n", 8)
    a.i("JR NZ,e", "wf_3")
    ldh_r(H_SCX)
    a.i("SUB n", 2)
    ldh_w(H_SCX)
    a.i("JR e", "wf_anim")
    a.label("wf_3")
    ldh_r(H_SCX)
    a.i("ADD A,n", 2)
    ldh_w(H_SCX)
    a.label("wf_anim")
    # player sprite struct 0: picture 1, image idx from walk counter, centre of screen
    set_a(1)
    st(0xC100)
    ldh_r(H_WALK)
    a.i("AND n", 0x04)
    a.i("LD B,A")
    a.i("LD A,C")
    a.i("ADD A,A")
    a.i("ADD A,B")
    st(0xC102)
    set_a(60)
    st(0xC104)
    set_a(64)
    st(0xC106)
    a.i("POP AF")
    a.i("RET NZ")
    # step finished: move in RAM
    a.i("LD A,C")
    a.i("OR A")
    a.i("JR NZ,e", "ws_1")
    ld(0xD361)
    a.i("INC A")
    st(0xD361)
    a.i("JR e", "ws_done")
    a.label("ws_1")
    a.i("CP n", 4)
    a.i("JR NZ,e", "ws_2")
    ld(0xD361)
    a.i("DEC A")
    st(0xD361)
    a.i("JR e", "ws_done")
    a.label("ws_2")
    a.i("CP n", 8)
    a.i("JR NZ,e", "ws_3")
    ld(0xD362)
    a.i("DEC A")
    st(0xD362)
    a.i("JR e", "ws_done")
    a.label("ws_3")
    ld(0xD362)
    a.i("INC A")
    st(0xD362)
    a.label("ws_done")
    # map edge (coordinate outside 0..23) -> new map
    ld(0xD361)
    a.i("CP n", 24)
    a.i("JP NC,nn", "MapChange")
    ld(0xD362)
    a.i("CP n", 24)
    a.i("JP NC,nn", "MapChange")
    # wild encounter: 1 in 16 steps on "grass" tiles (tile bit 3)
    ld(0xCFC6)
    a.i("BIT 3,A")
    a.i("RET Z")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x0F)
    a.i("RET NZ")
    a.i("JP nn", "BattleStart")

    a.label("MapChange")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x1F)
    a.i("LD E,A")
    a.i("LD D,n", 0)
    a.i("LD HL,nn", "MapTable")
    a.i("ADD HL,DE")
    a.i("LD A,(HL)")
    st(0xD35E)
    a.i("CALL nn", "Random")
    a.i("AND n", 0x0F)
    a.i("ADD A,n", 4)
    st(0xD361)
    a.i("CALL nn", "Random")
    a.i("AND n", 0x0F)
    a.i("ADD A,n", 4)
    st(0xD362)
    a.i("XOR A")
    st(0xCD4D)
    ldh_w(H_SCX)
    ldh_w(H_SCY)
    # NPC population: (map & 7) + 2 sprites, positions from RNG
    a.i("LD HL,nn", 0xC110)
    a.i("LD B,n", 15)
    a.label("mc_npc")
    a.i("PUSH HL")
    ld(0xD35E)
    a.i("AND n", 0x07)
    a.i("ADD A,n", 2)
    a.i("CP B")
    a.i("LD A,n", 0)
    a.i("JR C,e", "mc_np0")
    a.i("CALL nn", "Random")
    a.i("OR n", 1)
    a.label("mc_np0")
    a.i("LD (HL+),A")
    a.i("INC L")
    a.i("CALL nn", "Random")
    a.i("LD (HL+),A")  # image idx (also drives attributes)
    a.i("INC L")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x7F)
    a.i("LD (HL+),A")  # Y
    a.i("INC L")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x7F)
    a.i("LD (HL),A")  # X
    a.i("POP HL")
    a.i("LD A,L")
    a.i("ADD A,n", 16)
    a.i("LD L,A")
    a.i("DEC B")
    a.i("JR NZ,e", "mc_npc")
    # tileset reload with the LCD off (Pokemon: DisableLCD / LoadTilesetTilePatternData / EnableLCD)
    a.i("CALL nn", "DisableLCD")
    ld(0xD35E)
    a.i("AND n", 0x1C)
    a.i("ADD A,n", 0x40)
    a.i("LD H,A")
    a.i("LD L,n", 0)
    a.i("LD DE,nn", 0x9000)
    a.i("LD BC,nn", 0x0600)
    ld(0xD35E)
    a.i("AND n", 0x03)
    a.i("ADD A,n", 4)
    a.i("CALL nn", "FarCopyData")
    # new tile map into wTileMap
    a.i("LD HL,nn", 0x5800)
    ld(0xD35E)
    a.i("ADD A,L")
    a.i("LD L,A")
    a.i("LD DE,nn", W_TILEMAP)
    a.i("LD BC,nn", 360)
    ld(0xD35E)
    a.i("AND n", 0x03)
    a.i("ADD A,n", 4)
    a.i("CALL nn", "FarCopyData")
    a.i("CALL nn", "EnableLCD")
    a.i("RET")

    a.label("MapTable")
    a.db(*MAP_TABLE)
    a.label("ItemTable")
    a.db(*ITEM_TABLE)

    a.label("Interact")
    # text box: font loaded flag, window at the bottom third, timer
    set_a(1)
    st(0xCFC4)
    set_a(96)
    ldh_w(H_WY)
    a.i("CALL nn", "Random")
    a.i("AND n", 0x1F)
    a.i("ADD A,n", 10)
    ldh_w(H_TEXT)
    a.i("CALL nn", "Random")
    a.i("AND n", 0x0F)
    a.i("ADD A,A")
    a.i("LD E,A")
    a.i("LD D,n", 0)
    a.i("LD HL,nn", "EventTable")
    a.i("ADD HL,DE")
    a.i("LD A,(HL+)")
    a.i("LD H,(HL)")
    a.i("LD L,A")
    a.i("JP HL")

    a.label("EventTable")
    for name in ["EvFlag", "EvFlag", "EvItem", "EvLevel", "EvHeal", "EvBadge", "EvSeen", "EvOwn", "EvMove", "EvParty", "EvHurt", "EvCut", "EvMoney", "EvFlag", "EvNone", "EvNone"]:
        a.dw(name)

    a.label("EvNone")
    a.i("RET")

    a.label("EvFlag")  # set a random event flag bit in D747..D885
    a.i("CALL nn", "Random")
    a.i("LD E,A")
    a.i("LD D,n", 0)
    a.i("LD HL,nn", 0xD747)
    a.i("ADD HL,DE")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x3F)
    a.i("LD E,A")
    a.i("ADD HL,DE")
    a.i("CALL nn", "RandomBit")
    a.i("OR (HL)")
    a.i("LD (HL),A")
    a.i("RET")

    a.label("RandomBit")  # A = 1 << (Random & 7)
    a.i("CALL nn", "Random")
    a.i("AND n", 0x07)
    a.i("LD B,A")
    set_a(1)
    a.i("INC B")
    a.label("rb_l")
    a.i("DEC B")
    a.i("RET Z")
    a.i("ADD A,A")
    a.i("JR e", "rb_l")

    a.label("EvItem")  # append an item to the bag (max 10 slots scanned by the wrapper)
    ld(0xD31D)
    a.i("CP n", 12)
    a.i("RET NC")
    a.i("LD C,A")
    a.i("INC A")
    st(0xD31D)
    a.i("CALL nn", "Random")
    a.i("AND n", 0x0F)
    a.i("LD E,A")
    a.i("LD D,n", 0)
    a.i("LD HL,nn", "ItemTable")
    a.i("ADD HL,DE")
    a.i("LD B,(HL)")
    a.i("LD A,C")
    a.i("ADD A,A")
    a.i("LD E,A")
    a.i("LD HL,nn", 0xD31E)
    a.i("ADD HL,DE")
    a.i("LD (HL),B")
    a.i("INC HL")
    set_a(1)
    a.i("LD (HL+),A")
    set_a(0xFF)
    a.i("LD (HL),A")
    a.i("RET")

    a.label("EvLevel")
    ld(0xD18C)
    a.i("CP n", 100)
    a.i("RET NC")
    a.i("INC A")
    st(0xD18C)
    ld(0xD18E)
    a.i("ADD A,n", 3)
    st(0xD18E)
    a.i("RET NC")
    ld(0xD18D)
    a.i("INC A")
    st(0xD18D)
    a.i("RET")

    a.label("EvHeal")
    ld(0xD18D)
    st(0xD16C)
    ld(0xD18E)
    st(0xD16D)
    a.i("RET")

    a.label("EvBadge")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x03)
    a.i("RET NZ")
    ld(0xD356)
    a.i("SCF")
    a.i("RLA")
    st(0xD356)
    a.i("RET")

    a.label("EvSeen")
    a.i("LD HL,nn", 0xD30A)
    a.i("JR e", "ev_dex")
    a.label("EvOwn")
    a.i("LD HL,nn", 0xD2F7)
    a.label("ev_dex")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x0F)
    a.i("LD E,A")
    a.i("LD D,n", 0)
    a.i("ADD HL,DE")
    a.i("CALL nn", "RandomBit")
    a.i("OR (HL)")
    a.i("LD (HL),A")
    a.i("RET")

    a.label("EvMove")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x03)
    a.i("LD E,A")
    a.i("LD D,n", 0)
    a.i("LD HL,nn", 0xD173)
    a.i("ADD HL,DE")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x7F)
    a.i("CP n", 8)
    a.i("JR NC,e", "em_st")
    set_a(15)  # Cut
    a.label("em_st")
    a.i("LD (HL),A")
    a.i("RET")

    a.label("EvParty")
    ld(0xD163)
    a.i("CP n", 6)
    a.i("RET NC")
    a.i("LD C,A")
    a.i("INC A")
    st(0xD163)
    # struct base = D16B + 0x2C * C
    a.i("LD HL,nn", 0xD16B)
    a.i("LD DE,nn", 0x2C)
    a.i("INC C")
    a.label("ep_mul")
    a.i("DEC C")
    a.i("JR Z,e", "ep_go")
    a.i("ADD HL,DE")
    a.i("JR e", "ep_mul")
    a.label("ep_go")
    a.i("CALL nn", "Random")
    a.i("OR n", 1)
    a.i("LD (HL+),A")  # species
    a.i("XOR A")
    a.i("LD (HL+),A")  # HP hi
    set_a(15)
    a.i("LD (HL),A")  # HP lo
    a.i("LD DE,nn", 6)
    a.i("ADD HL,DE")
    set_a(0x21)
    a.i("LD (HL),A")  # move 1
    a.i("LD DE,nn", 0x19)
    a.i("ADD HL,DE")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x07)
    a.i("ADD A,n", 3)
    a.i("LD (HL+),A")  # level
    a.i("XOR A")
    a.i("LD (HL+),A")
    set_a(15)
    a.i("LD (HL),A")  # max HP
    a.i("RET")

    a.label("EvHurt")
    ld(0xD16C)
    a.i("OR A")
    a.i("JR NZ,e", "eh_big")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x0F)
    a.i("LD B,A")
    ld(0xD16D)
    a.i("SUB B")
    a.i("JR NC,e", "eh_ok")
    a.i("XOR A")
    a.label("eh_ok")
    st(0xD16D)
    a.i("RET")
    a.label("eh_big")
    a.i("DEC A")
    st(0xD16C)
    a.i("RET")

    a.label("EvCut")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x03)
    a.i("RET NZ")
    set_a(0x3D)
    st(0xCD4D)
    set_a(1)
    st(0xCD6A)
    st(0xCFCB)
    a.i("RET")

    a.label("EvMoney")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x77)
    st(0xD349)
    a.i("RET")

    a.label("TextFrame")
    a.i("DEC A")
    ldh_w(H_TEXT)
    a.i("JR Z,e", "tf_close")
    ldh_r(H_JOYPRESSED)
    a.i("BIT 1,A")
    a.i("JR NZ,e", "tf_close")
    # print one character every 4 frames into the bottom rows of wTileMap
    ldh_r(H_FRAME)
    a.i("AND n", 0x03)
    a.i("RET NZ")
    ldh_r(H_TEXT)
    a.i("AND n", 0x0F)
    a.i("LD E,A")
    a.i("LD D,n", 0)
    a.i("LD HL,nn", W_TILEMAP + 14 * 20 + 1)
    a.i("ADD HL,DE")
    a.i("CALL nn", "Random")
    a.i("OR n", 0x80)
    a.i("LD (HL),A")
    a.i("RET")
    a.label("tf_close")
    a.i("XOR A")
    ldh_w(H_TEXT)
    st(0xCFC4)
    set_a(144)
    ldh_w(H_WY)
    a.i("RET")

    a.label("MenuOpen")
    set_a(1)
    ldh_w(H_MENU)
    st(0xCFC4)
    set_a(0xD3)
    st(0xCC30)
    set_a(0xC3)
    st(0xCC31)
    a.i("XOR A")
    st(0xCC26)
    st(0xCF13)
    set_a(6)
    a.i("LDH (n),A", 0x8C)
    a.i("CALL nn", "Random")
    a.i("AND n", 0x03)
    st(0xCF94)
    # draw the menu box into wTileMap (right columns)
    a.i("LD HL,nn", W_TILEMAP + 10)
    a.i("LD B,n", 14)
    a.label("mo_row")
    a.i("LD C,n", 10)
    a.label("mo_col")
    set_a(0x7F)
    a.i("LD (HL+),A")
    a.i("DEC C")
    a.i("JR NZ,e", "mo_col")
    a.i("LD DE,nn", 10)
    a.i("ADD HL,DE")
    a.i("DEC B")
    a.i("JR NZ,e", "mo_row")
    a.i("RET")

    a.label("MenuFrame")
    ldh_r(H_JOYPRESSED)
    a.i("LD B,A")
    a.i("BIT 1,B")
    a.i("JR NZ,e", "mf_close")
    a.i("BIT 3,B")
    a.i("JR NZ,e", "mf_close")
    a.i("BIT 7,B")
    a.i("JR Z,e", "mf_up")
    ld(0xCC26)
    a.i("INC A")
    a.i("AND n", 0x07)
    st(0xCC26)
    ld(0xCC30)
    a.i("ADD A,n", 0x28)
    st(0xCC30)
    a.i("RET")
    a.label("mf_up")
    a.i("BIT 6,B")
    a.i("JR Z,e", "mf_a")
    ld(0xCC26)
    a.i("DEC A")
    a.i("AND n", 0x07)
    st(0xCC26)
    ld(0xCC30)
    a.i("SUB n", 0x28)
    st(0xCC30)
    a.i("RET")
    a.label("mf_a")
    a.i("BIT 0,B")
    a.i("RET Z")
    ld(0xCC26)
    st(0xCF94)
    a.i("RET")
    a.label("mf_close")
    a.i("XOR A")
    ldh_w(H_MENU)
    st(0xCFC4)
    a.i("RET")

    a.label("SaveSRAM")
    set_a(0x0A)
    st(0x0000)
    a.i("CALL nn", "Random")
    a.i("AND n", 0x03)
    st(0x4000)
    a.i("CALL nn", "Random")
    a.i("AND n", 0x1F)
    a.i("OR n", 0xA0)
    a.i("LD D,A")
    a.i("LD E,n", 0)
    a.i("LD HL,nn", 0xD163)
    a.i("LD BC,nn", 64)
    a.i("CALL nn", "CopyData")
    a.i("LD A,(nn)", 0xA000)
    st(0xC0F2)
    a.i("XOR A")
    st(0x0000)
    a.i("LD A,(nn)", 0xA000)  # reads 0xFF once disabled
    st(0xC0F3)
    a.i("RET")

    a.label("BattleStart")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x3F)
    a.i("ADD A,n", 60)
    ldh_w(H_BATTLE)
    a.i("CALL nn", "Random")
    a.i("AND n", 0x01)
    a.i("INC A")
    st(0xD057)
    a.i("CALL nn", "Random")
    a.i("OR n", 1)
    st(0xD059)
    a.i("CALL nn", "Random")
    a.i("AND n", 0x1F)
    a.i("ADD A,n", 2)
    st(0xD8C5)
    set_a(1)
    st(0xCFC4)
    a.i("XOR A")
    ldh_w(H_WY)  # full-screen window, like Pokemon's battle screen
    st(0xCCD5)
    set_a(0xC1)
    st(0xCC30)
    set_a(0xC4)
    st(0xCC31)
    a.i("RET")

    a.label("BattleFrame")
    a.i("DEC A")
    ldh_w(H_BATTLE)
    a.i("JR Z,e", "bf_end")
    a.i("AND n", 0x0F)
    a.i("RET NZ")
    # one turn every 16 frames
    ld(0xCCD5)
    a.i("INC A")
    st(0xCCD5)
    a.i("CALL nn", "EvHurt")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x03)
    a.i("ADD A,A")
    a.i("ADD A,A")
    a.i("ADD A,A")
    a.i("ADD A,n", 0xA9)
    st(0xCC30)
    # fainted?
    ld(0xD16C)
    a.i("LD B,A")
    ld(0xD16D)
    a.i("OR B")
    a.i("RET NZ")
    set_a(0xFF)
    st(0xD057)
    set_a(2)
    ldh_w(H_BATTLE)
    a.i("RET")
    a.label("bf_end")
    ld(0xD057)
    a.i("CP n", 0xFF)
    a.i("JR NZ,e", "bf_won")
    a.i("CALL nn", "EvHeal")  # blackout: heal and warp
    a.i("CALL nn", "bf_clear")
    a.i("JP nn", "MapChange")
    a.label("bf_won")
    a.i("CALL nn", "Random")
    a.i("AND n", 0x01)
    a.i("CALL Z,nn", "EvLevel")
    a.i("CALL nn", "EvSeen")
    a.label("bf_clear")
    a.i("XOR A")
    st(0xD057)
    st(0xD059)
    st(0xCFC4)
    st(0xCC30)
    st(0xCC31)
    set_a(144)
    ldh_w(H_WY)
    a.i("RET")
    assert a.here() < 0x8000, hex(a.here())

    # ------------------------------------------------------------- data banks
    for bank in range(4, 8):  # tile patterns + tile maps
        rom[bank * 0x4000 : (bank + 1) * 0x4000] = _xorshift_bytes(seed + 10 + bank, 0x4000)
    rng = random.Random(seed + 99)
    # tile maps with structure (runs of the same tile), stored at 0x5800 of banks 4..7
    for bank in range(4, 8):
        base = bank * 0x4000 + 0x1800
        t = 0
        for i in range(0x800):
            if rng.random() < 0.3:
                t = rng.getrandbits(8)
            rom[base + i] = t
    for bank in range(8, 40):  # per-map collision/grass blocks
        rom[bank * 0x4000 : (bank + 1) * 0x4000] = _xorshift_bytes(seed + 100 + bank, 0x4000)
    a.link()
    finalize_header(rom, title="SYNTHPOKE")
    return bytes(rom)


# ----------------------------------------------------------------------------- conformance ROM


def _dest_reg(m: str) -> str:
    """Register an operand-less / imm8 mnemonic writes ('' if only A/F/B/C/L/(HL) are touched)."""
    head, _, rest = m.partition(" ")
    if head == "LD":
        d = rest.split(",")[0]
    elif head in ("INC", "DEC", "RLC", "RRC", "RL", "RR", "SLA", "SRA", "SWAP", "SRL"):
        d = rest
    elif head in ("RES", "SET"):
        d = rest.split(",")[1]
    else:
        d = ""
    return d if d in ("D", "E", "H") else ""


def build_conformance_rom(seed: int = 7, n_blocks: int = 600) -> bytes:
    """Random-program ROM exercising every legal opcode (base + CB page).

    A chain of straight-line blocks in banks 1..N.  Each block mutates registers and a 256-byte WRAM
    scratch page through (HL), does balanced stack traffic, one "special" (16-bit/SP/IO/MBC/LCDC/DMA/
    HALT/EI-DI forms) and hops to the next block through a random control-flow instruction; every few
    blocks all registers are appended to a log ring at 0xD000-0xDEFF.  VBlank, STAT(LYC), timer and
    joypad interrupts fire underneath.  It rarely halts, so it is also the 100 %-busy stress case.
    Oracle and CUDA are compared through full save-state equality.
    """
    rng = random.Random(seed)
    rom = bytearray([0xFF]) * ROM_SIZE
    a = Asm(rom)
    small = {0x08: ["PUSH HL", ("LD HL,nn", 0xDF00), "XOR (HL)", "LD (HL),A", "POP HL", "RET"], 0x10: ["INC A", "RET"], 0x18: ["CPL", "RET"],
             0x20: ["SCF", "RET"], 0x28: ["CCF", "RET"], 0x30: ["DAA", "RET"]}
    for v in range(0, 0x40, 8):
        a.org(v)
        if v in small:
            for ins in small[v]:
                if isinstance(ins, tuple):
                    a.i(*ins)
                else:
                    a.i(ins)
            assert a.here() <= v + 8
        else:
            a.i("JP nn", "Start")
    for v, name in ((0x40, "VBlank"), (0x48, "StatInt"), (0x50, "TimerInt"), (0x60, "JoyInt")):
        a.org(v)
        a.i("JP nn", name)
    a.org(0x58)
    a.i("RETI")
    a.org(0x100)
    a.i("NOP")
    a.i("JP nn", "Start")

    a.org(0x150)
    a.label("Start")
    a.i("DI")
    a.i("LD SP,nn", 0xDFF0)
    a.i("LD A,n", 1)
    a.i("LD (nn),A", 0x2100)
    a.i("XOR A")
    a.i("LDH (n),A", 0x0F)
    a.i("LD A,n", 0x1F)
    a.i("LDH (n),A", 0xFF)
    a.i("LD A,n", 0x40)
    a.i("LDH (n),A", 0x41)  # STAT: LYC interrupt
    a.i("LD A,n", 77)
    a.i("LDH (n),A", 0x45)
    a.i("LD A,n", 0x13)
    a.i("LDH (n),A", 0x06)
    a.i("LD A,n", 0x06)
    a.i("LDH (n),A", 0x07)  # TAC enabled /64
    a.i("LD HL,nn", 0xC800)
    a.i("LD DE,nn", 0xD000)
    a.i("EI")
    a.i("JP nn", "blk0")

    def handler(name: str, hram: int):
        a.label(name)
        a.i("PUSH AF")
        a.i("LDH A,(n)", hram)
        a.i("INC A")
        a.i("LDH (n),A", hram)
        a.i("POP AF")
        a.i("RETI")

    handler("VBlank", 0x90)
    handler("StatInt", 0x91)
    handler("TimerInt", 0x92)
    handler("JoyInt", 0x93)

    a.label("LogRegs")  # 8-byte record H L B C F A SPlo D at (DE); ring over 0xD000-0xDEFF
    a.i("PUSH AF")
    for r in ("H", "L", "B", "C"):
        a.i(f"LD A,{r}")
        a.i("LD (DE),A")
        a.i("INC E")
    a.i("PUSH HL")
    a.i("LD HL,SP+e", 2)
    a.i("LD A,(HL+)")
    a.i("LD (DE),A")
    a.i("INC E")
    a.i("LD A,(HL)")
    a.i("LD (DE),A")
    a.i("INC E")
    a.i("LD A,L")
    a.i("LD (DE),A")
    a.i("INC E")
    a.i("LD A,D")
    a.i("LD (DE),A")
    a.i("INC E")
    a.i("LD A,E")
    a.i("AND n", 0xF8)
    a.i("LD E,A")
    a.i("JR NZ,e", "lr_done")
    a.i("INC D")
    a.i("LD A,D")
    a.i("CP n", 0xDF)
    a.i("JR C,e", "lr_done")
    a.i("LD D,n", 0xD0)
    a.label("lr_done")
    a.i("POP HL")
    a.i("POP AF")
    a.i("RET")

    a.label("tramp")  # A = bank, continue at 0x4000
    a.i("LD (nn),A", 0x2100)
    a.i("JP nn", 0x4000)

    heads = ("ADD", "ADC", "SUB", "SBC", "AND", "XOR", "OR", "CP", "INC", "DEC", "LD", "RLC", "RRC", "RL", "RR", "SLA", "SRA", "SWAP", "SRL", "BIT", "RES", "SET")
    pool = []
    for m, (opc, kind) in OPTABLE.items():
        head = m.split(" ")[0]
        if head not in heads or kind not in ("", "n"):
            continue
        operands = m.partition(" ")[2]
        if head == "LDH" or any(t in operands for t in ("SP", "BC", "DE", "HL,", "(C)", "(HL+)", "(HL-)")) or operands == "HL":
            continue
        pool.append(m)
    misc = ["RLCA", "RRCA", "RLA", "RRA", "DAA", "CPL", "SCF", "CCF", "NOP"]
    cover = set()

    def emit(m: str, operand=None):
        cover.add(m)
        a.i(m, operand)

    def body(n_ops: int):
        for _ in range(n_ops):
            k = rng.random()
            if k < 0.80:
                m = rng.choice(pool)
                d = _dest_reg(m)
                pair = {"D": "DE", "E": "DE", "H": "HL"}.get(d)
                if pair:
                    emit(f"PUSH {pair}")
                emit(m, rng.getrandbits(8) if OPTABLE[m][1] == "n" else None)
                if pair:
                    emit(f"POP {pair}")
            elif k < 0.86:
                emit(rng.choice(misc))
            elif k < 0.90:
                emit(f"PUSH {rng.choice(['BC', 'AF'])}")
                emit(rng.choice(misc))
                emit(f"POP {rng.choice(['BC', 'AF'])}")
            elif k < 0.93:
                emit(rng.choice(["LD A,(HL+)", "LD A,(HL-)", "LD (HL+),A", "LD (HL-),A"]))
                emit("LD H,n", 0xC8)
            elif k < 0.95:
                emit(f"RST {rng.choice([0x08, 0x10, 0x18, 0x20, 0x28, 0x30]):02X}")
            elif k < 0.96:
                emit("LDH (n),A", rng.choice([0xA0, 0xA1, 0xA2, 0xA3, 0x04, 0x05, 0x0F, 0x41, 0x45, 0x47, 0x01, 0x24]))
            elif k < 0.98:
                emit("LDH A,(n)", rng.choice([0xA0, 0xA1, 0x04, 0x05, 0x0F, 0x41, 0x44, 0x00, 0x40, 0x46, 0x26, 0x01, 0xFF]))
            else:
                emit(rng.choice(["INC BC", "DEC BC", "ADD HL,BC", "ADD HL,HL"]))
                emit("LD H,n", 0xC8)

    specials = ["LD (nn),SP", "LD SP,HL", "ADD SP,e", "LD HL,SP+e", "LD (C),A", "LD A,(C)", "LD (nn),A", "LD A,(nn)", "LD A,(BC)", "LD (BC),A",
                "LD A,(DE)", "LD (DE),A", "JP HL", "CALL cc", "RET cc", "RETI", "STOP", "ADD HL,DE", "ADD HL,SP", "INC/DEC rp", "EI/DI", "HALT",
                "POP/PUSH DE/HL", "DMA", "LCDC", "MBC", "LD rp,nn", "JOYP"]
    bank = 1
    a.org(0x4000, bank=bank)
    for blk in range(n_blocks):
        if a.here() > 0x7D00:
            bank += 1
            a.i("LD A,n", bank)
            a.i("JP nn", "tramp")
            a.org(0x4000, bank=bank)
        a.label(f"blk{blk}")
        body(rng.randint(8, 40))
        sp = specials[blk % len(specials)] if blk < 2 * len(specials) else rng.choice(specials)
        if sp == "LD (nn),SP":
            emit("LD (nn),SP", 0xC900 + rng.getrandbits(7) * 2)
        elif sp == "LD SP,HL":
            emit("PUSH HL")
            emit("LD HL,SP+e", 2)
            emit("LD SP,HL")
            emit("DEC SP")
            emit("DEC SP")
            emit("POP HL")
        elif sp == "ADD SP,e":
            v = rng.choice([1, 2, 3, 0x7F, 0x10, 8])
            emit("ADD SP,e", (-v) & 0xFF)
            emit("ADD SP,e", v)
        elif sp == "LD HL,SP+e":
            emit("LD HL,SP+e", rng.getrandbits(8))
            emit("LD A,L")
            emit("XOR H")
            emit("LD HL,nn", 0xC800 + rng.getrandbits(8))
        elif sp in ("LD (C),A", "LD A,(C)"):
            emit("LD C,n", rng.choice([0xA4, 0xA5, 0xA6, 0x42, 0x43, 0x4A, 0x4B, 0x06]))
            emit(sp)
        elif sp in ("LD (nn),A", "LD A,(nn)"):
            emit(sp, rng.choice([0xC900, 0xCA00, 0xE900, 0xFEA0, 0xFE10, 0x8000, 0x9C00, 0xA100]) + rng.getrandbits(6))
        elif sp in ("LD A,(BC)", "LD (BC),A"):
            emit("LD B,n", 0xCA)
            emit(sp)
        elif sp in ("LD A,(DE)", "LD (DE),A"):
            emit("PUSH DE")
            emit("LD D,n", 0xCB)
            emit(sp)
            emit("POP DE")
        elif sp == "JP HL":
            emit("PUSH HL")
            emit("LD HL,nn", f"jphl{blk}")
            emit("JP HL")
            a.db(0xD3)  # never executed
            a.label(f"jphl{blk}")
            emit("POP HL")
        elif sp == "CALL cc":
            emit(f"CALL {rng.choice(['NZ', 'Z', 'NC', 'C'])},nn", "LogRegs")
            emit("CALL nn", "LogRegs")
        elif sp == "RET cc":
            emit("CALL nn", f"sub{blk}")
            emit("JR e", f"after{blk}")
            a.label(f"sub{blk}")
            emit(f"RET {rng.choice(['NZ', 'Z', 'NC', 'C'])}")
            emit("INC B")
            emit("RET")
            a.label(f"after{blk}")
        elif sp == "RETI":
            emit("CALL nn", f"sub{blk}")
            emit("JR e", f"after{blk}")
            a.label(f"sub{blk}")
            emit("DI")
            emit("RETI")
            a.label(f"after{blk}")
        elif sp == "STOP":
            emit("STOP", 0x00)
        elif sp in ("ADD HL,DE", "ADD HL,SP"):
            emit(sp)
            emit("LD A,H")
            emit("XOR L")
            emit("LD HL,nn", 0xC800 + rng.getrandbits(8))
        elif sp == "INC/DEC rp":
            emit(rng.choice(["INC SP", "DEC SP"]))
            emit(rng.choice(["INC SP", "DEC SP"]))
            emit("LD SP,nn", 0xDFF0)
            emit("PUSH DE")
            emit(rng.choice(["INC DE", "DEC DE"]))
            emit("LD A,E")
            emit("XOR D")
            emit("POP DE")
            emit(rng.choice(["INC HL", "DEC HL"]))
            emit("LD H,n", 0xC8)
        elif sp == "EI/DI":
            emit("DI")
            body(4)
            emit("EI")
        elif sp == "HALT":
            if rng.random() < 0.5:
                emit("DI")  # HALT with IME=0: wakes without dispatch
                emit("HALT")
                emit("EI")
            else:
                emit("HALT")
        elif sp == "POP/PUSH DE/HL":
            emit("PUSH DE")
            emit("PUSH HL")
            emit("POP DE")
            emit("POP HL")
            emit("PUSH DE")
            emit("PUSH HL")
            emit("POP DE")
            emit("POP HL")
        elif sp == "DMA":
            emit("LD A,n", rng.choice([0xC8, 0xC9, 0xD0, 0x80, 0x40]))
            emit("LDH (n),A", 0x46)
        elif sp == "LCDC":
            emit("LD A,n", rng.choice([0x91, 0xE3, 0xF7, 0x11, 0x63, 0x83, 0xB9, 0xE3, 0x91]))
            emit("LDH (n),A", 0x40)
            emit("LD A,n", rng.getrandbits(8))
            emit("LDH (n),A", rng.choice([0x42, 0x43, 0x4A, 0x4B, 0x47, 0x48, 0x49, 0x45]))
        elif sp == "MBC":
            emit("LD A,n", rng.choice([0x0A, 0x00, 0x0B, 0x1A]))
            emit("LD (nn),A", 0x0000 + rng.getrandbits(12))
            emit("LD A,n", rng.getrandbits(2))
            emit("LD (nn),A", 0x4000 + rng.getrandbits(12))
            emit("LD A,n", rng.getrandbits(8))
            emit("LD (nn),A", 0xA000 + rng.getrandbits(12))
            emit("LD A,(nn)", 0xA000 + rng.getrandbits(12))
            emit("LD (nn),A", 0x6000 + rng.getrandbits(12))
        elif sp == "LD rp,nn":
            emit("LD BC,nn", rng.getrandbits(16))
            emit("PUSH DE")
            emit("LD DE,nn", rng.getrandbits(16))
            emit("LD A,D")
            emit("ADD A,E")
            emit("POP DE")
        elif sp == "JOYP":
            emit("LD A,n", rng.choice([0x10, 0x20, 0x30, 0x00]))
            emit("LDH (n),A", 0x00)
            emit("LDH A,(n)", 0x00)
        nxt = f"blk{blk + 1}" if blk + 1 < n_blocks else "wrap"
        k = rng.random()
        if k < 0.3:
            emit("JP nn", nxt)
        elif k < 0.5:
            emit(f"JP {rng.choice(['NZ', 'Z', 'NC', 'C'])},nn", nxt)
            emit("JP nn", nxt)
        elif k < 0.7:
            emit(f"JR {rng.choice(['NZ', 'Z', 'NC', 'C'])},e", f"t{blk}")
            emit("INC C")
            a.label(f"t{blk}")
        elif k < 0.8:
            emit("JR e", f"t{blk}")
            a.db(0xDD)
            a.label(f"t{blk}")
        if blk % 7 == 0:
            emit("CALL nn", "LogRegs")
    a.label("wrap")
    a.i("LDH A,(n)", 0x90)
    a.i("BIT 0,A")
    a.i("JR Z,e", "wrap2")
    emit("RST 00")
    a.label("wrap2")
    emit("RST 38")
    a.link()
    finalize_header(rom, title="SYNTHCONF")
    build_conformance_rom.last_missing = sorted(m for m in OPTABLE if m not in cover)  # type: ignore[attr-defined]
    return bytes(rom)


def build_halt_edge_rom() -> bytes:
    """HALT / LCD-event edge cases, one scenario per frame, cycling (HRAM 0xA0 = scenario counter):

    0 plain HALT until VBlank                         5 HALT with only the joypad interrupt enabled (sleeps across
    1 SCX/SCY written just before HALT (scanline          frames until a button edge arrives)
      parameters pending while halted)                6 LCD switched off, spin, switched on again, HALT
    2 STAT mode-0 source armed, not enabled in IE     7 HALT with IME=0 and a pending enabled interrupt (wakes at once)
    3 LYC=100 STAT interrupt wakes the HALT mid-frame 8 window + sprites enabled, WX/WY/LCDC.4 changed before HALT
    4 LY written mid-frame, then HALT; TIMA running   9 STAT mode-2 + LYC sources armed and enabled; HALT twice
    Every handler counts into HRAM so missed or extra wake-ups change the state."""
    rom = bytearray([0xFF]) * ROM_SIZE
    a = Asm(rom)
    for v in range(0, 0x40, 8):
        a.org(v)
        a.i("RET")
    for v, name in ((0x40, "VBlank"), (0x48, "StatInt"), (0x50, "TimerInt"), (0x58, "SerialInt"), (0x60, "JoyInt")):
        a.org(v)
        a.i("JP nn", name)
    a.org(0x100)
    a.i("NOP")
    a.i("JP nn", "Start")
    a.org(0x150)

    def handler(name: str, hram: int):
        a.label(name)
        a.i("PUSH AF")
        a.i("LDH A,(n)", hram)
        a.i("INC A")
        a.i("LDH (n),A", hram)
        a.i("LDH A,(n)", 0x44)  # fold LY at wake-up time into a checksum
        a.i("PUSH HL")
        a.i("LD HL,nn", 0xC100)
        a.i("ADD A,(HL)")
        a.i("RLCA")
        a.i("LD (HL),A")
        a.i("POP HL")
        a.i("POP AF")
        a.i("RETI")

    handler("VBlank", 0x90)
    handler("StatInt", 0x91)
    handler("TimerInt", 0x92)
    handler("SerialInt", 0x94)
    handler("JoyInt", 0x93)

    def out(reg: int, v: int):
        a.i("LD A,n", v)
        a.i("LDH (n),A", reg)

    a.label("WaitVBlankLine")  # spin (no HALT) until LY == 145
    a.i("LDH A,(n)", 0x44)
    a.i("CP n", 145)
    a.i("JR NZ,e", "WaitVBlankLine")
    a.i("RET")

    a.label("Start")
    a.i("DI")
    a.i("LD SP,nn", 0xDFF0)
    out(0x0F, 0)
    out(0xFF, 0x01)
    out(0x40, 0x91)
    out(0x47, 0xE4)
    out(0x48, 0xD0)
    out(0x49, 0xE0)
    a.i("XOR A")
    a.i("LDH (n),A", 0xA0)
    # a few tiles, a map and two sprites so rendered frames are not blank
    a.i("LD HL,nn", 0x8000)
    a.i("LD B,n", 0)
    a.label("fill_tiles")
    a.i("LD A,L")
    a.i("XOR H")
    a.i("RRCA")
    a.i("LD (HL+),A")
    a.i("DEC B")
    a.i("JR NZ,e", "fill_tiles")
    a.i("LD HL,nn", 0x9800)
    a.i("LD B,n", 0)
    a.label("fill_map")
    a.i("LD A,L")
    a.i("AND n", 0x0F)
    a.i("LD (HL+),A")
    a.i("DEC B")
    a.i("JR NZ,e", "fill_map")
    a.i("LD HL,nn", 0xFE00)
    for v in (40, 30, 3, 0x00, 90, 100, 5, 0x30):
        a.i("LD A,n", v)
        a.i("LD (HL+),A")
    a.i("EI")

    a.label("MainLoop")
    a.i("LDH A,(n)", 0xA0)
    a.i("INC A")
    a.i("CP n", 10)
    a.i("JR C,e", "store_scn")
    a.i("XOR A")
    a.label("store_scn")
    a.i("LDH (n),A", 0xA0)
    for k in range(10):
        a.i("CP n", k)
        a.i("JP Z,nn", f"scn{k}")
    a.i("JP nn", "scn0")

    def end():
        a.i("JP nn", "MainLoop")

    a.label("scn0")
    a.i("HALT")
    end()

    a.label("scn1")
    a.i("LDH A,(n)", 0x90)
    a.i("LDH (n),A", 0x43)
    a.i("CPL")
    a.i("LDH (n),A", 0x42)
    a.i("HALT")
    end()

    a.label("scn2")
    out(0x41, 0x08)
    a.i("HALT")
    out(0x41, 0x00)
    out(0x0F, 0)
    end()

    a.label("scn3")
    out(0x45, 100)
    out(0x41, 0x40)
    out(0xFF, 0x03)
    a.i("HALT")  # woken by LYC=100
    a.i("HALT")  # then by VBlank
    out(0x41, 0x00)
    out(0xFF, 0x01)
    end()

    a.label("scn4")
    out(0x06, 0xF0)
    out(0x07, 0x05)
    out(0xFF, 0x05)
    a.i("LDH A,(n)", 0x44)
    a.i("ADD A,n", 7)
    a.i("LDH (n),A", 0x44)  # LY is writable in this PyBoy line
    a.i("HALT")
    a.i("HALT")
    out(0x07, 0x00)
    out(0xFF, 0x01)
    end()

    a.label("scn5")
    out(0x00, 0x10)  # select buttons
    out(0xFF, 0x10)
    a.i("HALT")  # sleeps until a joypad edge (may span many frames)
    a.i("LDH A,(n)", 0x00)  # which button: shifts this env's timing relative to its neighbours
    a.i("AND n", 0x0F)
    a.i("LD B,A")
    a.i("INC B")
    a.i("LD HL,nn", 0xC101)
    a.label("joy_spin")
    a.i("INC (HL)")
    a.i("DEC B")
    a.i("JR NZ,e", "joy_spin")
    out(0x00, 0x20)  # select directions for the next time round
    a.i("LDH A,(n)", 0x00)
    a.i("LD (nn),A", 0xC102)
    out(0xFF, 0x01)
    out(0x0F, 0)
    end()

    a.label("scn6")
    a.i("CALL nn", "WaitVBlankLine")
    out(0x40, 0x11)  # LCD off
    a.i("LD B,n", 200)
    a.label("off_spin")
    a.i("DEC B")
    a.i("JR NZ,e", "off_spin")
    out(0x40, 0x91)
    a.i("HALT")
    end()

    a.label("scn7")
    a.i("DI")
    out(0x0F, 0x01)  # VBlank already pending, IME=0: HALT falls through without dispatch
    a.i("HALT")
    a.i("INC B")
    out(0x0F, 0x00)
    a.i("EI")
    a.i("HALT")
    end()

    a.label("scn8")
    out(0x4A, 60)
    out(0x4B, 47)
    out(0x40, 0xF3)  # window + sprites, unsigned tile data
    a.i("HALT")
    out(0x40, 0xE3)  # signed tile data (scanline parameter 5)
    a.i("HALT")
    out(0x40, 0x91)
    end()

    a.label("scn9")
    out(0x45, 20)
    out(0x41, 0x60)
    out(0xFF, 0x03)
    a.i("HALT")
    a.i("HALT")
    out(0x41, 0x00)
    out(0xFF, 0x01)
    out(0x0F, 0)
    a.i("HALT")
    end()

    a.link()
    finalize_header(rom, title="HALTEDGE")
    return bytes(rom)


def build_lcd_probe_rom(seed: int = 3, n_blocks: int = 400) -> bytes:
    """A seeded random program of LCD *observations*: reads of LY / STAT / DIV, polling loops on LY and on the STAT mode,
    writes to SCX/SCY/WX/WY/LYC/LY/STAT/LCDC/IE/TAC, LCD off-on periods, HALTs and OAM DMA, separated by delays of
    random length so that every access lands at a different point of the scanline.  Everything read is folded into a
    WRAM checksum and every handler folds LY and STAT at dispatch time, so an LCD mode change applied at the wrong
    instruction boundary changes the state.  Exercises the lazy LCD of the CUDA interpreter (gb_device.cuh lcd_catch_up)
    against the oracle's event-per-tick loop."""
    rng = random.Random(seed)
    rom = bytearray([0xFF]) * ROM_SIZE
    a = Asm(rom)
    for v in range(0, 0x40, 8):
        a.org(v)
        a.i("RET")
    for v, name in ((0x40, "VBlank"), (0x48, "StatInt"), (0x50, "TimerInt"), (0x58, "SerialInt"), (0x60, "JoyInt")):
        a.org(v)
        a.i("JP nn", name)
    a.org(0x100)
    a.i("NOP")
    a.i("JP nn", "Start")
    a.org(0x150)

    def fold():  # checksum = rlca(checksum + A), kept at 0xC100
        a.i("PUSH HL")
        a.i("LD HL,nn", 0xC100)
        a.i("ADD A,(HL)")
        a.i("RLCA")
        a.i("LD (HL),A")
        a.i("POP HL")

    def handler(name: str, hram: int):
        a.label(name)
        a.i("PUSH AF")
        a.i("LDH A,(n)", hram)
        a.i("INC A")
        a.i("LDH (n),A", hram)
        a.i("LDH A,(n)", 0x44)
        fold()
        a.i("LDH A,(n)", 0x41)
        fold()
        a.i("POP AF")
        a.i("RETI")

    handler("VBlank", 0x90)
    handler("StatInt", 0x91)
    handler("TimerInt", 0x92)
    handler("SerialInt", 0x94)
    handler("JoyInt", 0x93)

    def out(reg: int, v: int):
        a.i("LD A,n", v)
        a.i("LDH (n),A", reg)

    a.label("Start")
    a.i("DI")
    a.i("LD SP,nn", 0xDFF0)
    out(0x0F, 0)
    out(0xFF, 0x01)
    out(0x40, 0x91)
    out(0x47, 0xE4)
    out(0x48, 0xD0)
    out(0x49, 0xE0)
    out(0x00, 0x20)
    a.i("LD HL,nn", 0x8000)
    a.i("LD B,n", 0)
    a.label("fill_tiles")
    a.i("LD A,L")
    a.i("XOR H")
    a.i("RRCA")
    a.i("LD (HL+),A")
    a.i("DEC B")
    a.i("JR NZ,e", "fill_tiles")
    a.i("LD HL,nn", 0x9800)
    a.i("LD BC,nn", 0x0800)
    a.label("fill_map")
    a.i("LD A,L")
    a.i("AND n", 0x0F)
    a.i("LD (HL+),A")
    a.i("DEC BC")
    a.i("LD A,B")
    a.i("OR C")
    a.i("JR NZ,e", "fill_map")
    a.i("LD HL,nn", 0xFE00)
    for v in (40, 30, 3, 0x00, 90, 100, 5, 0x30, 91, 20, 7, 0x60):
        a.i("LD A,n", v)
        a.i("LD (HL+),A")
    a.i("EI")
    a.label("MainLoop")
    uid = [0]

    def lab(prefix: str) -> str:
        uid[0] += 1
        return f"{prefix}{uid[0]}"

    def delay():
        n = rng.choice([1, 2, 3, 5, 9, 17, 33, 70, 150, 255])
        l = lab("dly")
        a.i("LD B,n", n)
        a.label(l)
        a.i("DEC B")
        a.i("JR NZ,e", l)
        if rng.random() < 0.3:
            for _ in range(rng.randrange(1, 4)):
                a.i("NOP")

    lcd_on = True
    for _ in range(n_blocks):
        delay()
        k = rng.randrange(20)
        if k < 3:
            a.i("LDH A,(n)", 0x44)
            fold()
        elif k < 6:
            a.i("LDH A,(n)", 0x41)
            fold()
        elif k == 6:
            a.i("LDH A,(n)", 0x04)
            a.i("AND n", 0xF0)  # DIV's low bits depend on nothing else we check; keep the slow-moving part
            fold()
        elif k == 7:  # poll LY (bounded)
            if lcd_on:
                l, e = lab("ply"), lab("plx")
                a.i("LD C,n", 200)
                a.label(l)
                a.i("LDH A,(n)", 0x44)
                a.i("CP n", rng.randrange(0, 154))
                a.i("JR Z,e", e)
                a.i("DEC C")
                a.i("JR NZ,e", l)
                a.label(e)
                a.i("LD A,C")
                fold()
        elif k == 8:  # poll the STAT mode (bounded)
            l, e = lab("pst"), lab("psx")
            a.i("LD C,n", 120)
            a.label(l)
            a.i("LDH A,(n)", 0x41)
            a.i("AND n", 3)
            a.i("CP n", rng.randrange(0, 4))
            a.i("JR Z,e", e)
            a.i("DEC C")
            a.i("JR NZ,e", l)
            a.label(e)
            a.i("LD A,C")
            fold()
        elif k == 9:
            a.i("LD A,(nn)", 0xC100)
            a.i("LDH (n),A", rng.choice([0x42, 0x43, 0x4A, 0x4B]))
        elif k == 10:
            out(0x45, rng.choice([0, 1, 17, 80, 143, 144, 150, 153, 200]))
        elif k == 11:
            out(0x41, rng.choice([0x00, 0x00, 0x00, 0x08, 0x10, 0x20, 0x40, 0x48, 0x78]))
        elif k == 12:
            out(0xFF, rng.choice([0x01, 0x01, 0x03, 0x05, 0x11, 0x13, 0x07]))
        elif k == 13:
            if lcd_on and rng.random() < 0.4:
                out(0x40, 0x11)
                lcd_on = False
            else:
                out(0x40, rng.choice([0x91, 0xB1, 0xF3, 0xE3, 0x93, 0x97, 0x81]))
                lcd_on = True
        elif k == 14:
            if not lcd_on:
                out(0x40, 0x91)
                lcd_on = True
            a.i("HALT")
        elif k == 15:
            a.i("LDH A,(n)", 0x44)
            a.i("ADD A,n", rng.choice([1, 3, 250]))
            a.i("LDH (n),A", 0x44)
        elif k == 16:
            out(0x06, rng.choice([0x00, 0xF0, 0x80]))
            out(0x07, rng.choice([0x00, 0x04, 0x05, 0x06, 0x07]))
        elif k == 17:
            out(0x46, 0xC1)
        elif k == 18:
            a.i(rng.choice(["DI", "EI"]))
        else:
            out(0x0F, rng.choice([0x00, 0x02, 0x04]))
    if not lcd_on:
        out(0x40, 0x91)
    out(0x07, 0x00)
    out(0xFF, 0x01)
    out(0x41, 0x00)
    a.i("EI")
    a.i("JP nn", "MainLoop")
    a.link()
    finalize_header(rom, title="LCDPROBE")
    return bytes(rom)


def rom_catalog() -> Dict[str, Tuple]:
    return {
        "pokelike": (build_pokelike_rom, {}),
        "pokelike_timer": (build_pokelike_rom, {"timer": True}),
        "busy": (build_pokelike_rom, {"always_busy": True}),
        "conformance": (build_conformance_rom, {}),
        "conformance_b": (build_conformance_rom, {"seed": 99, "n_blocks": 900}),  # second instruction stream / interrupt phase
        "halt_edge": (build_halt_edge_rom, {}),
        "lcd_probe": (build_lcd_probe_rom, {}),
        "lcd_probe_b": (build_lcd_probe_rom, {"seed": 11, "n_blocks": 700}),
    }
