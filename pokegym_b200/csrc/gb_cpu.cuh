// gb_cpu.cuh -- SM83 interpreter of k_run_frames (CPU.tick of PyBoy 1.6: interrupt check, HALT, one instruction,
// then the peripherals advance by its T-cycles; interrupt dispatch costs 0 cycles).
//
// Design (B200-first; nothing here resembles a CPU interpreter's fetch/decode/execute switch):
//   * The env's whole `Machine` lives in SHARED memory; only what every instruction touches is cached in registers:
//     the two packed register words, SP, PC, the ROM bank offset, a cycle countdown and a 2-bit mode word.  Everything
//     that is not plain execution of a ROM instruction (interrupt dispatch, HALT, RAM-resident code, IO / MBC / cart
//     RAM accesses, a running TIMA) is an out-of-line function that works on the shared-memory machine directly --
//     no by-value argument lists, no scratch copies, no stack.
//   * One 128-bit read-only load fetches the pre-decoded control word of the instruction (gb_predecode.h).  Operand
//     fetch (register byte by PRMT, immediate, or ONE shared copy of the bus read), register write-back (two PRMTs
//     with ready-made selectors) and the deferred bus write are the same straight-line code for all instructions, so
//     lanes of a warp that execute different instructions stay converged everywhere except inside the short handler
//     bodies, which the switch groups by handler id (lanes with the same id run together, whatever their PC).
//   * Time is ONE register: `rem` = cycles until this env's next LCD mode change.  lcd.clock and the DIV counter are
//     brought up to date (time_sync) only when something observes them.
#pragma once
#include "gb_device.cuh"
#include "gb_predecode.h"

__constant__ uint4 c_base_desc[512];  // per-opcode base descriptors (pd_build_base), uploaded once per process

#define MODE_ATTN 0x8000u   // pending interrupt / HALT / PyBoy's interrupt_queued latch: the tick starts in cpu_attention
#define MODE_POST 0x10000u  // HALT or a running TIMA: the tick ends in cpu_post_slow

struct RunCtx {  // uniform per launch
    const uint4 *rom_dec;
    uint32_t bank_mask;  // rom_banks - 1 when that is a power of two, else 0
};

__device__ __forceinline__ uint32_t hot_mode(const Machine &m) {
    const uint32_t attn = m.halted | m.iq | (m.iflag & m.ie & 0x1F), post = m.halted | (m.tmr & 0x04000000u);
    return (attn ? MODE_ATTN : 0u) | (post ? MODE_POST : 0u);
}
// cycles until LCD.tick has something to do: clock >= clock_target (LCD on) or >= one frame (LCD off)
__device__ __forceinline__ int hot_rem(const Machine &m) { return (int)(((m.lcdc & 0x80) ? m.target : FRAME_CYCLES) - m.clock); }
// bring lcd.clock and the DIV counter up to the interpreter's countdown
__device__ __forceinline__ void time_sync(Machine &m, int rem) {
    const uint32_t e = (uint32_t)(m.t_sync - rem);
    m.clock += e;
    m.divc += e;
    m.t_sync = rem;
}

// one thread per ROM offset
__global__ void k_predecode_rom(const uint8_t *rom, uint32_t rom_len, uint4 *out) {
    uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= rom_len) return;
    uint32_t in_bank = o & 0x3FFF;
    if (in_bank >= 0x3FFD) {  // operands would come from another bank: decode at run time
        out[o] = make_uint4(H_SLOW, 0, 0x32103210u, 0);
        return;
    }
    uint32_t ins = rom[o] | (rom[o + 1] << 8) | (rom[o + 2] << 16);
    out[o] = pd_finish(c_base_desc[(ins & 0xFF) == 0xCB ? (256u | ((ins >> 8) & 0xFF)) : (ins & 0xFF)], ins, o < 0x4000 ? o : 0x4000 + in_bank);
}

// byte-wise decode through the bus: RAM-resident code and the last three bytes of a ROM bank
__device__ GB_NOINLINE void cpu_decode_slow(Machine &m, uint32_t pc, uint32_t *out) {
    uint32_t ins = 0;
    for (uint32_t i = 0; i < 3; i++) ins |= bus_read_full(m, (pc + i) & 0xFFFF) << (8 * i);
    const uint4 d = pd_finish(c_base_desc[(ins & 0xFF) == 0xCB ? (256u | ((ins >> 8) & 0xFF)) : (ins & 0xFF)], ins, pc);
    out[0] = d.x; out[1] = d.y; out[2] = d.z; out[3] = d.w;
}

// Start of a tick that needs attention (CPU.tick / CPU.check_interrupts / CPU.handle_interrupt of PyBoy 1.6).
// Works on the spilled machine.  Returns 1 when an instruction is to be executed, else (T-cycles << 8).
__device__ GB_NOINLINE uint32_t cpu_attention(Machine &m) {
    if (!m.iq) {
        const uint32_t pending = m.iflag & m.ie & 0x1F;
        if (pending) {  // highest-priority pending source
            const uint32_t bit = pending & (0u - pending);
            if (m.halted) m.pc = (m.pc + 1) & 0xFFFF;
            if (m.ime) {
                m.iflag ^= bit;
                bus_write_full(m, (m.sp - 1) & 0xFFFF, m.pc >> 8);
                bus_write_full(m, (m.sp - 2) & 0xFFFF, m.pc & 0xFF);
                m.sp = (m.sp - 2) & 0xFFFF;
                m.pc = 0x40 + 8 * (31 - __clz(bit));
                m.ime = 0;
            }
            m.iq = 1;
            m.halted = 0;
            return 0;  // dispatch costs no cycles and executes nothing
        }
    } else if (m.halted) {  // debugger-only path in PyBoy: halted with a queued interrupt
        m.halted = 0;
        m.pc = (m.pc + 1) & 0xFFFF;
    }
    if (m.halted) return 4u << 8;
    m.iq = 0;
    return 1;
}

// End of a tick while halted or with TIMA running: HALT fast-forward (Motherboard.tick) and Timer.tick's TIMA half.
__device__ GB_NOINLINE uint32_t cpu_post_slow(Machine &m, uint32_t cycles) {
    if (m.halted) {  // fast-forward to the next LCD mode change / timer overflow
        const int a = (int)(m.target - m.clock), b = timer_cycles_to_interrupt(m), c = a < b ? a : b;
        cycles = c < 0 ? 0u : (uint32_t)c;
    }
    timer_tick_tima(m, cycles);
    return cycles;
}

// Interprets until this env's LCD clock reaches its next mode change.  Registers in/out by reference (they stay in
// registers: this function is inlined into the frame loop); `m` is the env's machine in shared memory.
__device__ __forceinline__ void cpu_run_to_event(Machine &m, const RunCtx &cx, uint32_t &bcde, uint32_t &hlaf, uint32_t &sp, uint32_t &pc,
                                                 uint32_t &rom_off, uint32_t &n_instr, uint32_t *scratch) {
    uint8_t *const memb = m.memb;
    const uint8_t *const rom = m.rom;
    int rem = hot_rem(m);
    m.t_sync = rem;
    uint32_t mode = hot_mode(m);
    auto rd8 = [&](uint32_t a) -> uint32_t {  // Motherboard.getitem: WRAM (+ echo), HRAM and ROM inline, the rest out of line
        if (a - 0xC000u < 0x3E00u) { const uint32_t i = MEM_WRAM + (a & 0x1FFF); return memb[((i >> 2) << 7) | (i & 3)]; }
        if (a - 0xFF80u < 0x7Fu) { const uint32_t i = MEM_HI + (a - 0xFE00); return memb[((i >> 2) << 7) | (i & 3)]; }
        if (a < 0x8000) return __ldg(rom + (a + (a >> 14) * rom_off));
        time_sync(m, rem);  // DIV is read from the synced counter
        return bus_read_slow(m, a);
    };
    do {
        uint32_t cyc = 0, wn = 0, wa = 0, wv = 0;
        uint4 d;
        bool execute = true, decoded = false;
        if ((pc | mode) & 0x8000u) {  // attention, or code outside the ROM
            if (mode & MODE_ATTN) {
                m.pc = pc; m.sp = sp;
                time_sync(m, rem);
                const uint32_t r = cpu_attention(m);
                pc = m.pc; sp = m.sp; rom_off = m.rom_off;  // the dispatch pushes through the full bus
                rem = hot_rem(m);
                m.t_sync = rem;
                mode = hot_mode(m);
                execute = r & 1;
                cyc = r >> 8;
            }
            if (execute && (pc & 0x8000u)) {
                time_sync(m, rem);
                cpu_decode_slow(m, pc, scratch);
                d = make_uint4(scratch[0], scratch[1], scratch[2], scratch[3]);
                decoded = true;
            }
        }
        if (execute) {
            if (!decoded) d = __ldg(cx.rom_dec + (pc + (pc >> 14) * rom_off));
            for (;;) {  // runs once; H_SLOW re-enters with the descriptor decoded on the fly
                // ---- operand fetch (uniform)
                const uint32_t w = d.w, imm16 = d.y & 0xFFFFu, op = d.x >> 24, f = hlaf >> 24;
                uint32_t v = gb_prmt(bcde, hlaf, w);  // byte 0 = source register (upper bytes: don't care)
                if (w & PDF_IMM) v = imm16;
                const uint32_t ar = gb_prmt(bcde, hlaf, w >> 16) & 0xFFFFu;
                uint32_t addr = (w & PDF_AIMM) ? imm16 : ar;
                if (w & PDF_ASP) addr = sp;
                if (w & PDF_RD) {
                    v = rd8(addr);
                    if (w & PDF_RD16) v |= rd8((addr + 1) & 0xFFFF) << 8;
                }
                // ---- handler
                uint32_t rv = v, next_pc = d.y >> 16;
                wv = v; wa = addr; wn = (w >> 9) & 1;
                cyc = (d.x >> 8) & 0xFF;
                bool again = false;
                const uint32_t h = d.x & 0xFF;
                if (h != H_MOV) switch (h) {  // plain moves (a third of all instructions) are done: rv = v
                case H_HLI: rv = ((ar + (uint32_t)(int32_t)(int8_t)op) & 0xFFFFu) | (v << 16); break;
                case H_ARITH: {  // branch-free: a subtraction adds the complement and inverts the carries
                    const uint32_t a = (hlaf >> 16) & 0xFF, ex = (d.x >> 16) & 0xFF, x = (v & 0xFF) ^ ex;
                    const uint32_t sum = a + x + ((((f >> 4) & op) ^ ex) & 1);  // carry in for ADC / SBC only
                    const uint32_t res = sum & 0xFF;
                    const uint32_t nf = (((((a ^ x ^ sum) & 0x10) << 1) | ((sum >> 4) & 0x10)) ^ (ex & 0x30)) | (ex & FLAG_N) | (res == 0 ? FLAG_Z : 0);
                    rv = res | (nf << 8);
                    break;
                }
                case H_LOGIC: {  // AND: a & v; XOR: a ^ v; OR: (a & v) | (a ^ v)
                    const uint32_t a = (hlaf >> 16) & 0xFF, ex = (d.x >> 16) & 0xFF, b = v & 0xFF;
                    const uint32_t res = ((a & b) & ex) | ((a ^ b) & op);
                    rv = res | (((op ? 0 : FLAG_H) | (res == 0 ? FLAG_Z : 0)) << 8);
                    break;
                }
                case H_INCDEC: {  // op = +1 / -1 (mod 256), ex = 0 / N|H: DEC inverts the half carry like a subtraction
                    const uint32_t b = v & 0xFF, ex = (d.x >> 16) & 0xFF, sum = b + op, res = sum & 0xFF;
                    const uint32_t nf = (f & FLAG_C) | ((((b ^ op ^ sum) & 0x10) << 1) ^ ex) | (res == 0 ? FLAG_Z : 0);
                    rv = res | (nf << 8);
                    wv = res;
                    break;
                }
                case H_ADD_HL: {
                    const uint32_t hl = hlaf & 0xFFFF, t = hl + ar;
                    const uint32_t nf = (f & FLAG_Z) | (((hl & 0xFFF) + (ar & 0xFFF)) > 0xFFF ? FLAG_H : 0) | (t > 0xFFFF ? FLAG_C : 0);
                    rv = (t & 0xFFFF) | (nf << 24);
                    break;
                }
                case H_JUMP: {
                    const uint32_t ex = (d.x >> 16) & 0xFF;
                    if (((f ^ ex) & op) == 0) { next_pc = imm16; cyc += ex & 0xF; }
                    break;
                }
                case H_CALL: {
                    const uint32_t ex = (d.x >> 16) & 0xFF;
                    if (((f ^ ex) & op) == 0) {  // high byte at SP-1 first, then low byte at SP-2
                        wa = (sp - 1) & 0xFFFF;
                        wv = __byte_perm(next_pc, 0, 0x4401);
                        wn = 2;
                        sp = (sp - 2) & 0xFFFF;
                        next_pc = imm16;
                        cyc += ex & 0xF;
                    }
                    break;
                }
                case H_RET: {
                    const uint32_t ex = (d.x >> 16) & 0xFF;
                    if (w & PDF_RETI) m.ime = 1;
                    if (((f ^ ex) & op) == 0) {
                        next_pc = v & 0xFFFF;
                        sp = (sp + 2) & 0xFFFF;
                        cyc += ex & 0xF;
                    }
                    break;
                }
                case H_PUSH:
                    wa = (sp - 1) & 0xFFFF;
                    wv = __byte_perm(ar, 0, 0x4401);
                    wn = 2;
                    sp = (sp - 2) & 0xFFFF;
                    break;
                case H_POP:
                    rv = v & (0xFF00u | op);
                    sp = (sp + 2) & 0xFFFF;
                    break;
                case H_ROT: {
                    const uint32_t b = v & 0xFF, c = (f >> 4) & 1;
                    uint32_t cout, res;
                    switch (op & 7) {
                    case 0: cout = b >> 7; res = (b << 1) | cout; break;        // RLC
                    case 1: cout = b & 1; res = (b >> 1) | (cout << 7); break;  // RRC
                    case 2: cout = b >> 7; res = (b << 1) | c; break;           // RL
                    case 3: cout = b & 1; res = (b >> 1) | (c << 7); break;     // RR
                    case 4: cout = b >> 7; res = b << 1; break;                 // SLA
                    case 5: cout = b & 1; res = (b >> 1) | (b & 0x80); break;   // SRA
                    case 6: cout = 0; res = (b >> 4) | (b << 4); break;         // SWAP
                    default: cout = b & 1; res = b >> 1; break;                 // SRL
                    }
                    res &= 0xFF;
                    const uint32_t nf = ((res == 0 && !(op & 8)) ? FLAG_Z : 0) | (cout ? FLAG_C : 0);  // RLCA..RRA clear Z
                    rv = res | (nf << 8);
                    wv = res;
                    break;
                }
                case H_BIT: rv = ((f & FLAG_C) | FLAG_H | ((v & imm16 & 0xFF) ? 0 : FLAG_Z)) << 8; break;
                case H_RESSET:
                    rv = (v & imm16 & 0xFF) | (imm16 >> 8);
                    wv = rv;
                    break;
                case H_SLOW:
                    time_sync(m, rem);
                    cpu_decode_slow(m, pc, scratch);
                    d = make_uint4(scratch[0], scratch[1], scratch[2], scratch[3]);
                    again = true;
                    break;
                default: {  // H_RARE: `op` is the opcode
                    const uint32_t a = (hlaf >> 16) & 0xFF, hl = hlaf & 0xFFFF;
                    switch (op) {
                    case 0x76: m.halted = 1; mode |= MODE_ATTN | MODE_POST; next_pc = pc; break;  // HALT: PC stays on the HALT byte
                    case 0x10: break;                                                              // STOP skips a byte
                    case 0xF3: m.ime = 0; break;
                    case 0xFB: m.ime = 1; break;  // PyBoy: EI takes effect immediately
                    case 0x27: {                  // DAA
                        uint32_t corr = ((f & FLAG_H) ? 0x06 : 0) | ((f & FLAG_C) ? 0x60 : 0), t = a;
                        if (f & FLAG_N) {
                            t -= corr;
                        } else {
                            if ((t & 0x0F) > 9) corr |= 0x06;
                            if (t > 0x99) corr |= 0x60;
                            t += corr;
                        }
                        t &= 0xFF;
                        hlaf = (hlaf & 0xFFFF) | (t << 16) | (((f & FLAG_N) | (t == 0 ? FLAG_Z : 0) | ((corr & 0x60) ? FLAG_C : 0)) << 24);
                        break;
                    }
                    case 0x2F: hlaf = (hlaf & 0xFFFF) | ((~a & 0xFF) << 16) | ((f | FLAG_N | FLAG_H) << 24); break;                // CPL
                    case 0x37: hlaf = (hlaf & 0x00FFFFFFu) | (((f & FLAG_Z) | FLAG_C) << 24); break;                               // SCF
                    case 0x3F: hlaf = (hlaf & 0x00FFFFFFu) | (((f & FLAG_Z) | ((f & FLAG_C) ^ FLAG_C)) << 24); break;              // CCF
                    case 0x08: wa = imm16; wv = sp; wn = 3; break;                                                                // LD (nn),SP
                    case 0xE8:
                    case 0xF8: {  // ADD SP,e / LD HL,SP+e
                        const uint32_t imm8 = imm16 & 0xFF;
                        const uint32_t nf = (((sp & 0xF) + (imm8 & 0xF)) > 0xF ? FLAG_H : 0) | (((sp & 0xFF) + imm8) > 0xFF ? FLAG_C : 0);
                        const uint32_t t = (sp + ((imm8 ^ 0x80) - 0x80)) & 0xFFFF;
                        hlaf = (hlaf & 0x00FFFFFFu) | (nf << 24);
                        if (op == 0xE8) sp = t;
                        else hlaf = (hlaf & 0xFFFF0000u) | t;
                        break;
                    }
                    case 0xE9: next_pc = hl; break;                   // JP HL
                    case 0xF9: sp = hl; break;                        // LD SP,HL
                    case 0x31: sp = imm16; break;                     // LD SP,nn
                    case 0x33: sp = (sp + 1) & 0xFFFF; break;         // INC SP
                    case 0x3B: sp = (sp + 0xFFFF) & 0xFFFF; break;    // DEC SP
                    case 0x39: {                                      // ADD HL,SP
                        const uint32_t t = hl + sp;
                        const uint32_t nf = (f & FLAG_Z) | (((hl & 0xFFF) + (sp & 0xFFF)) > 0xFFF ? FLAG_H : 0) | (t > 0xFFFF ? FLAG_C : 0);
                        hlaf = (hlaf & 0x00FF0000u) | (t & 0xFFFF) | (nf << 24);
                        break;
                    }
                    case 0xE2: wa = 0xFF00u | (bcde & 0xFF); wv = a; wn = 1; break;  // LD (FF00+C),A
                    case 0xF2:                                                       // LD A,(FF00+C)
                        time_sync(m, rem);
                        hlaf = (hlaf & 0xFF00FFFFu) | (bus_read_slow(m, 0xFF00u | (bcde & 0xFF)) << 16);
                        break;
                    default: m.fault = 1; break;  // illegal opcode: PyBoy raises; 1-byte 4-cycle NOP + sticky fault
                    }
                    break;
                }
                }
                if (again) continue;
                // ---- register write-back (uniform)
                bcde = gb_prmt(bcde, rv, d.z);
                hlaf = gb_prmt(hlaf, rv, d.z >> 16);
                pc = next_pc;
                n_instr++;
                break;
            }
        }
        // ---- bus writes, HALT, running TIMA: one divergent region for everything that is not register work
        if (wn | (mode & MODE_POST)) {
            if (wn) {
                // n = 1: byte wv at wa; 2: low byte of wv at wa, then high byte at wa - 1 (pushes); 3: low byte at wa, then
                // high byte at wa + 1 (LD (nn),SP)
                const uint32_t count = wn > 1 ? 2u : 1u, second = (wn == 2 ? wa - 1 : wa + 1) & 0xFFFF;
#pragma unroll 1
                for (uint32_t i = 0; i < count; i++) {
                    const uint32_t a = i ? second : wa, b = (i ? wv >> 8 : wv) & 0xFF;
                    if (a - 0xC000u < 0x3E00u) {  // WRAM and its echo
                        const uint32_t k = MEM_WRAM + (a & 0x1FFF);
                        memb[((k >> 2) << 7) | (k & 3)] = (uint8_t)b;
                    } else if (a - 0xFF80u < 0x7Fu || a - 0xFE00u < 0x100u) {  // HRAM, OAM
                        const uint32_t k = MEM_HI + (a - 0xFE00);
                        memb[((k >> 2) << 7) | (k & 3)] = (uint8_t)b;
                    } else if (a - 0x8000u < 0x2000u) {  // VRAM
                        const uint32_t k = MEM_VRAM + (a - 0x8000);
                        memb[((k >> 2) << 7) | (k & 3)] = (uint8_t)b;
                    } else if (a - 0x2000u < 0x2000u) {  // MBC3 ROM bank select (constant traffic in banked games)
                        uint32_t bank = b & 0x7F;
                        bank = bank ? bank : 1;
                        m.rombank = bank;
                        rom_off = (cx.bank_mask ? (bank & cx.bank_mask) : (bank % m.rom_banks)) * 0x4000u - 0x4000u;
                        m.rom_off = rom_off;
                    } else {  // IO registers, other MBC registers, cart RAM, OAM DMA
                        time_sync(m, rem);
                        bus_write_rare(&m, a, b);
                        rom_off = m.rom_off;
                        rem = hot_rem(m);
                        m.t_sync = rem;
                        mode = hot_mode(m);
                    }
                }
            }
            if (mode & MODE_POST) {
                time_sync(m, rem);
                cyc = cpu_post_slow(m, cyc);
                mode = hot_mode(m);
            }
        }
        rem -= (int)cyc;
    } while (rem > 0);
    time_sync(m, rem);
}
