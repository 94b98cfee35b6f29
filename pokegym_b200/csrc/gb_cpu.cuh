// gb_cpu.cuh -- SM83 interpreter over the pre-decoded ROM (CPU.tick of PyBoy: interrupt check, HALT, one
// instruction; returns T-cycles, interrupt dispatch costs 0).
//
// One 64-bit descriptor load (gb_predecode.h) replaces instruction fetch, opcode bit-field decoding, operand
// fetch and length / cycle computation; a dense switch on the handler id dispatches to short specialised
// bodies that work on the packed register words.  Data reads inline only the hot regions (work RAM, HRAM, ROM)
// and call an out-of-line function with by-value arguments otherwise; all stores are deferred to one write
// site at the end of the step, so `Machine` never has its address taken and stays in registers.
#pragma once
#include "gb_device.cuh"
#include "gb_predecode.h"

__constant__ uint32_t c_base_desc[512];  // per-opcode base descriptors (pd_build_base), uploaded once per process

__device__ __forceinline__ uint2 pd_decode_bytes(uint32_t ins, uint32_t pc) {  // ins = opcode | op1 << 8 | op2 << 16
    uint32_t op = ins & 0xFF, imm16 = (ins >> 8) & 0xFFFF, imm8 = imm16 & 0xFF;
    uint32_t b = c_base_desc[op == 0xCB ? (256u | imm8) : op];
    uint32_t lo = imm16;
    if (PD_H(b) == H_JR) lo = (pc + 2 + ((imm8 ^ 0x80) - 0x80)) & 0xFFFF;  // branch target
    if (op == 0xCB) lo = imm8 >> 6;                                        // rotate / BIT / RES / SET
    return make_uint2(PD_BASE_WORD0(b), lo | (((pc + PD_BASE_LEN(b)) & 0xFFFF) << 16));
}

// one thread per ROM offset
__global__ void k_predecode_rom(const uint8_t *rom, uint32_t rom_len, uint2 *out) {
    uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= rom_len) return;
    uint32_t in_bank = o & 0x3FFF;
    if (in_bank >= 0x3FFD) {  // operands would come from another bank: decode at run time
        out[o] = make_uint2(H_SLOW, 0);
        return;
    }
    uint32_t ins = rom[o] | (rom[o + 1] << 8) | (rom[o + 2] << 16);
    out[o] = pd_decode_bytes(ins, o < 0x4000 ? o : 0x4000 + in_bank);
}

// cold half of a data read: VRAM, cart RAM, OAM, IO array, IO registers -- inputs by value
__device__ __noinline__ uint32_t rd8_slow(uint32_t a, const uint8_t *memb, const uint8_t *cramb, uint32_t ram, uint32_t lcd0, uint32_t scroll,
                                          uint32_t pal_ie, uint32_t tim, uint32_t iflag) {
    if (a < 0xA000) {
        uint32_t i = MEM_VRAM + (a - 0x8000);
        return memb[((i >> 2) << 7) | (i & 3)];
    }
    if (a < 0xC000) {
        if (!(ram & 0xFF)) return 0xFF;
        uint32_t i = ((ram >> 8) & 3) * 0x2000u + (a - 0xA000);
        return cramb[((i >> 2) << 7) | (i & 3)];
    }
    if (a >= 0xFF00) {
        uint32_t r = io_reg_read(a, lcd0, scroll, pal_ie, tim, iflag);
        if (r != IO_NOT_A_REGISTER) return r;
    }
    uint32_t i = MEM_HI + (a - 0xFE00);
    return memb[((i >> 2) << 7) | (i & 3)];
}

__device__ __forceinline__ uint32_t rd8(Machine &m, uint32_t a) {  // Motherboard.getitem
    if (a - 0xC000u < 0x3E00u) return mem_rd(m, MEM_WRAM + (a & 0x1FFF));      // WRAM and its echo
    if (a >= 0xFF80 && a != 0xFFFF) return mem_rd(m, MEM_HI + (a - 0xFE00));    // HRAM
    if (a < 0x8000) return __ldg(m.rom + (a < 0x4000 ? a : a + m.rom_off));     // ROM data tables
    return rd8_slow(a, m.memb, m.cramb, m.ram_en | (m.rambank << 8), m.lcdc | (m.stat << 8) | (m.ly << 16) | (m.lyc << 24), m.scroll,
                    m.pal | (m.ie << 24), ((m.div + (m.divc >> 8)) & 0xFF) | m.tmr, m.iflag);
}

// Bus writes of one instruction, deferred so that a single write site exists.  n = 0 none; 1: one byte (v & 0xFF) at a;
// 2: low byte of v at a, then high byte at a - 1 (pushes); 3: low byte at a, then high byte at a + 1 (LD (nn),SP).
struct DeferredWrites {
    uint32_t n, a, v;
};

__device__ __forceinline__ void cpu_commit_writes(Machine &m, const DeferredWrites &w) {
    const uint32_t count = w.n > 1 ? 2u : w.n, second = (w.n == 2 ? w.a - 1 : w.a + 1) & 0xFFFF;
    for (uint32_t i = 0; i < count; i++) bus_write_full(m, i ? second : w.a, i ? (w.v >> 8) & 0xFF : w.v & 0xFF);
}

// Executes one CPU.tick; the instruction's bus writes are returned in `dw` and must be committed by the caller
// (cpu_commit_writes) before anything else observes the machine.
__device__ __forceinline__ uint32_t cpu_step(Machine &m, const uint2 *__restrict__ rom_dec, DeferredWrites &dw) {
    // only `wn` is initialised: address / value are read solely when wn != 0
    uint32_t wn = 0, wa, wv;
    uint32_t cycles = 0;
#define PUSH16(val)                                                          \
    do {  /* high byte at SP-1 first, then low byte at SP-2 */               \
        uint32_t _v = (val);                                                 \
        wa = (m.sp - 1) & 0xFFFF;                                            \
        wv = __byte_perm(_v, 0, 0x4401); /* (_v >> 8 & 0xFF) | (_v & 0xFF) << 8 */ \
        wn = 2;                                                              \
        m.sp = (m.sp - 2) & 0xFFFF;                                          \
    } while (0)
#define WRITE8(addr, val) do { wa = (addr) & 0xFFFF; wv = (val); wn = 1; } while (0)

    bool execute = true, decoded = false;
    uint32_t pc = m.pc;
    uint2 d = make_uint2(0, 0);
    // One divergent region guards everything that is not plain execution of a pre-decoded instruction: pending
    // interrupt, HALT, PyBoy's interrupt_queued latch, RAM-resident code and instructions whose operands straddle a
    // 16 KiB bank boundary (the last three bytes of a bank; k_predecode_rom marks them H_SLOW).  All inputs of the
    // predicate are in registers, so the hot path pays two logic ops, a compare and one never-taken branch.
    const uint32_t attention = m.halted | m.iq | (m.iflag & m.ie & 0x1F);
    if (attention | (uint32_t)(pc >= 0x8000) | (uint32_t)((pc & 0x3FFF) >= 0x3FFD)) {
        if (attention) {
            if (!m.iq) {
                uint32_t pending = m.iflag & m.ie & 0x1F;
                if (pending) {  // CPU.handle_interrupt for the highest-priority pending source
                    uint32_t bit = pending & (0u - pending);
                    if (m.halted) m.pc = (m.pc + 1) & 0xFFFF;
                    if (m.ime) {
                        m.iflag ^= bit;
                        PUSH16(m.pc);
                        m.pc = 0x40 + 8 * (31 - __clz(bit));
                        m.ime = 0;
                    }
                    m.iq = 1;
                    m.halted = 0;
                    execute = false;
                }
            } else if (m.halted) {  // debugger-only path in PyBoy: halted with a queued interrupt
                m.halted = 0;
                m.pc = (m.pc + 1) & 0xFFFF;
            }
            if (execute && m.halted) { dw.n = 0; return 4; }
            pc = m.pc;
        }
        if (execute && (pc >= 0x8000 || (pc & 0x3FFF) >= 0x3FFD)) {  // byte-wise decode through the bus
            uint32_t ins = 0;
            for (uint32_t i = 0; i < 3; i++) ins |= rd8(m, (pc + i) & 0xFFFF) << (8 * i);
            d = pd_decode_bytes(ins, pc);
            decoded = true;
        }
    }
    if (execute) {
        // ---- fetch: one descriptor load replaces opcode fetch, operand fetch and decode
        if (!decoded) d = __ldg(rom_dec + (pc < 0x4000 ? pc : pc + m.rom_off));
        const uint32_t dx = d.x, y = PD_Y(dx), z = PD_Z(dx), p = y >> 1;
        const uint32_t imm16 = d.y & 0xFFFF, imm8 = d.y & 0xFF;
        const uint32_t hl = reg_hl(m);
        uint32_t next_pc = d.y >> 16;
        cycles = PD_CYC(dx);
        switch (PD_H(dx)) {
        case H_NOP: break;
        case H_LD_R_R: set_reg8(m, y, reg8(m, z)); break;
        case H_LD_R_HL: set_reg8(m, y, rd8(m, hl)); break;
        case H_LD_HL_R: WRITE8(hl, reg8(m, z)); break;
        case H_LD_R_N: set_reg8(m, y, imm8); break;
        case H_LD_HL_N: WRITE8(hl, imm8); break;
        case H_LD_A_RP:
        case H_LD_RP_A: {
            uint32_t a = p == 0 ? (m.bcde & 0xFFFF) : p == 1 ? (m.bcde >> 16) : hl;
            if (PD_H(dx) == H_LD_A_RP) set_a(m, rd8(m, a));
            else WRITE8(a, reg_a(m));
            if (p >= 2) set_hl(m, hl + (p == 2 ? 1u : 0xFFFFu));
            break;
        }
        case H_LDH_N_A: WRITE8(0xFF00u | imm8, reg_a(m)); break;
        case H_LDH_A_N: set_a(m, rd8(m, 0xFF00u | imm8)); break;
        case H_LD_C_A: WRITE8(0xFF00u | (m.bcde & 0xFF), reg_a(m)); break;
        case H_LD_A_C: set_a(m, rd8(m, 0xFF00u | (m.bcde & 0xFF))); break;
        case H_LD_NN_A: WRITE8(imm16, reg_a(m)); break;
        case H_LD_A_NN: set_a(m, rd8(m, imm16)); break;
        case H_ALU_R: alu8(m, y, reg8(m, z)); break;
        case H_ALU_HL: alu8(m, y, rd8(m, hl)); break;
        case H_ALU_N: alu8(m, y, imm8); break;
        case H_INCDEC_R:
        case H_INCDEC_HL: {
            const bool mem = PD_H(dx) == H_INCDEC_HL;
            uint32_t v = mem ? rd8(m, hl) : reg8(m, y), res, nf = reg_f(m) & FLAG_C;
            if (!(z & 1)) {  // z == 4: INC, z == 5: DEC
                res = (v + 1) & 0xFF;
                nf |= ((v & 0xF) == 0xF ? FLAG_H : 0);
            } else {
                res = (v - 1) & 0xFF;
                nf |= FLAG_N | ((v & 0xF) == 0 ? FLAG_H : 0);
            }
            if (res == 0) nf |= FLAG_Z;
            set_f(m, nf);
            if (mem) WRITE8(hl, res);
            else set_reg8(m, y, res);
            break;
        }
        case H_LD_RP_NN: set_reg_pair(m, p, imm16); break;
        case H_INCDEC_RP: set_reg_pair(m, p, reg_pair(m, p) + ((y & 1) ? 0xFFFFu : 1u)); break;
        case H_ADD_HL: {
            uint32_t r = reg_pair(m, p), t = hl + r;
            set_f(m, (reg_f(m) & FLAG_Z) | (((hl & 0xFFF) + (r & 0xFFF)) > 0xFFF ? FLAG_H : 0) | (t > 0xFFFF ? FLAG_C : 0));
            set_hl(m, t);
            break;
        }
        case H_JR:
            if (y == 0 || condition(m, y & 3)) { next_pc = imm16; cycles += PD_TAKEN_EXTRA(dx); }
            break;
        case H_JP:
            if (y == 0 || condition(m, y & 3)) { next_pc = imm16; cycles += PD_TAKEN_EXTRA(dx); }
            break;
        case H_CALL:
            if (y == 0 || condition(m, y & 3)) { PUSH16(next_pc); next_pc = imm16; cycles += PD_TAKEN_EXTRA(dx); }
            break;
        case H_RETI: m.ime = 1;  // fall through
        case H_RET:
            if (y == 0 || condition(m, y & 3)) {
                next_pc = rd8(m, m.sp) | (rd8(m, (m.sp + 1) & 0xFFFF) << 8);
                m.sp = (m.sp + 2) & 0xFFFF;
                cycles += PD_TAKEN_EXTRA(dx);
            }
            break;
        case H_RST: PUSH16(next_pc); next_pc = y * 8; break;
        case H_PUSH: PUSH16((p == 3) ? ((reg_a(m) << 8) | reg_f(m)) : reg_pair(m, p)); break;
        case H_POP: {
            uint32_t v = rd8(m, m.sp) | (rd8(m, (m.sp + 1) & 0xFFFF) << 8);
            m.sp = (m.sp + 2) & 0xFFFF;
            if (p == 3) set_af(m, v >> 8, v & 0xF0);
            else set_reg_pair(m, p, v);
            break;
        }
        case H_CB_R:
        case H_CB_HL: {
            const bool mem = PD_H(dx) == H_CB_HL;
            const uint32_t x = imm16, f = reg_f(m);  // CB page: word 1 carries opcode bits 6-7
            uint32_t v = mem ? rd8(m, hl) : reg8(m, z), res;
            if (x == 1) {  // BIT: Z from the tested bit, H set, C kept
                set_f(m, (f & FLAG_C) | FLAG_H | (((v >> y) & 1) ? 0 : FLAG_Z));
            } else {
                if (x == 0) {
                    uint32_t c = (f >> 4) & 1, cout;
                    switch (y) {
                    case 0: cout = v >> 7; res = (v << 1) | cout; break;        // RLC
                    case 1: cout = v & 1; res = (v >> 1) | (cout << 7); break;  // RRC
                    case 2: cout = v >> 7; res = (v << 1) | c; break;           // RL
                    case 3: cout = v & 1; res = (v >> 1) | (c << 7); break;     // RR
                    case 4: cout = v >> 7; res = v << 1; break;                 // SLA
                    case 5: cout = v & 1; res = (v >> 1) | (v & 0x80); break;   // SRA
                    case 6: cout = 0; res = (v >> 4) | (v << 4); break;         // SWAP
                    default: cout = v & 1; res = v >> 1; break;                 // SRL
                    }
                    res &= 0xFF;
                    set_f(m, (res == 0 ? FLAG_Z : 0) | (cout ? FLAG_C : 0));
                } else {
                    res = (x == 2) ? (v & ~(1u << y)) : (v | (1u << y));  // RES / SET
                }
                if (mem) WRITE8(hl, res);
                else set_reg8(m, z, res);
            }
            break;
        }
        case H_ROT_A: {  // RLCA RRCA RLA RRA: Z N H cleared, C = bit shifted out
            const uint32_t a = reg_a(m), cin = (reg_f(m) >> 4) & 1, left = !(y & 1), thru = y >> 1;
            const uint32_t out = left ? (a >> 7) : (a & 1), in = thru ? cin : out;
            set_af(m, left ? ((a << 1) | in) : ((a >> 1) | (in << 7)), out ? FLAG_C : 0);
            break;
        }
        default: {  // H_RARE
            const uint32_t op = y, a = reg_a(m), f = reg_f(m);  // H_RARE: Y is the whole opcode
            switch (op) {
            case 0x76: m.halted = 1; next_pc = pc; break;              // HALT: PC stays on the HALT byte
            case 0x10: break;                                          // STOP skips a byte (length 2 in the descriptor)
            case 0xF3: m.ime = 0; break;
            case 0xFB: m.ime = 1; break;  // PyBoy: EI takes effect immediately
            case 0x27: {                                                              // DAA
                uint32_t corr = ((f & FLAG_H) ? 0x06 : 0) | ((f & FLAG_C) ? 0x60 : 0), t = a;
                if (f & FLAG_N) {
                    t -= corr;
                } else {
                    if ((t & 0x0F) > 9) corr |= 0x06;
                    if (t > 0x99) corr |= 0x60;
                    t += corr;
                }
                t &= 0xFF;
                set_af(m, t, (f & FLAG_N) | (t == 0 ? FLAG_Z : 0) | ((corr & 0x60) ? FLAG_C : 0));
                break;
            }
            case 0x2F: set_af(m, ~a, f | FLAG_N | FLAG_H); break;                     // CPL
            case 0x37: set_f(m, (f & FLAG_Z) | FLAG_C); break;                        // SCF
            case 0x3F: set_f(m, (f & FLAG_Z) | ((f & FLAG_C) ^ FLAG_C)); break;       // CCF
            case 0x08:  // LD (nn),SP: low byte first
                wa = imm16; wv = m.sp; wn = 3;
                break;
            case 0xE8:
            case 0xF8: {  // ADD SP,e / LD HL,SP+e
                uint32_t sp = m.sp;
                set_f(m, (((sp & 0xF) + (imm8 & 0xF)) > 0xF ? FLAG_H : 0) | (((sp & 0xFF) + imm8) > 0xFF ? FLAG_C : 0));
                uint32_t t = (sp + ((imm8 ^ 0x80) - 0x80)) & 0xFFFF;
                if (op == 0xE8) m.sp = t;
                else set_hl(m, t);
                break;
            }
            case 0xE9: next_pc = hl; break;           // JP HL
            case 0xF9: m.sp = hl; break;              // LD SP,HL
            default: m.fault = 1; break;  // illegal opcode: PyBoy raises; 1-byte 4-cycle NOP + sticky fault
            }
            break;
        }
        }
        m.pc = next_pc;
        m.n_instr++;
        m.iq = 0;
    }
    dw.n = wn; dw.a = wa; dw.v = wv;
#undef PUSH16
#undef WRITE8
    return cycles;
}
