// gb_predecode.h -- pre-decoded instruction descriptors (host + device).
//
// The cartridge ROM is shared by every env and never changes, so it is decoded ONCE per handle: for each ROM
// offset the library stores a 16-byte descriptor and the interpreter fetches one 128-bit word per instruction from
// this L2-resident table.  The descriptor is not a compressed opcode: it is the control word of a small uniform
// datapath (gb_cpu.cuh) --
//
//     operand fetch  ->  handler (one of 15 short bodies; plain moves need none)  ->  register write-back  ->  deferred bus write
//
// whose operand-fetch and write-back stages are the same straight-line code for every instruction.  The eight
// 8-bit SM83 registers live in two packed words (bcde = C|B<<8|E<<16|D<<24, hlaf = L|H<<8|A<<16|F<<24), i.e.
// bytes 0..7 of a PRMT source pair, so "which register" is a PRMT selector and the descriptor carries the selectors
// ready-made:
//
//   x: handler | op << 8 | ex << 16 | cycles << 24
//        cycles  T-cycles (condition false / unconditional)
//        ex      low nibble: extra T-cycles when a conditional branch is taken; bits 4 / 7: flag value the
//                condition expects (compared under the mask in `op`)
//        (H_ROT: ex = pd_rot_ex of the rotate kind)
//        op      handler operand: ALU op, INC/DEC select, signed +-1 for HL+/HL-/INC rr/DEC rr, condition mask
//                (0 = always, 0x80 = Z, 0x10 = C), rotate kind, POP low-byte mask, or the opcode (H_RARE)
//   y: imm16 | next_pc << 16      imm16: immediate, or a ready-made value -- the JR/JP/CALL/RST target, 0xFF00|n
//                                 for LDH, the BIT mask, the RES/SET and/or masks
//   z: sel_lo | sel_hi << 16      write-back: bcde = prmt(bcde, rv, sel_lo); hlaf = prmt(hlaf, rv, sel_hi), where rv is
//                                 the handler's 32-bit result (0x3210 = keep; nibble k = 4 + j puts result byte j
//                                 into register byte k).  8-bit results travel as res | new_F << 8, so "ADD writes
//                                 A and F", "CP writes F only" and "INC B writes B and F" differ only in selectors.
//   w: srcsel | flags << 4 | asel << 16 | class << 24
//        srcsel  PRMT nibble of the 8-bit source register (byte 0..7 of bcde:hlaf)
//        asel    PRMT selector of the 16-bit address / pair operand (BC 0x10, DE 0x32, HL 0x54, AF-for-PUSH 0x67); the class
//                bits above it select result bytes 2 and 3, which every user masks off
//        class   dense id of (handler, flags) (gb_classes.inc), 0xFF outside the fast set
//        flags   PDF_* below
//
// Code executed from RAM (the HRAM OAM-DMA stub) and instructions whose operand bytes straddle a 16 KiB bank
// boundary are decoded on the fly from the same per-opcode base table (pd_finish).
#pragma once
#include <stdint.h>

enum {
    H_MOV = 0,     // rv = v : every 8-bit load / store and LD rr,nn (the selectors / PDF_WR say where it goes); NOP
    H_HLI,       // rv = (pair + op) | v << 16 : LD A,(HL+-)  LD (HL+-),A  INC rr  DEC rr
    H_ARITH,     // ADD ADC SUB SBC CP on A with v   (op bit 0: carry in; ex = 0xFF for the subtractions)
    H_LOGIC,     // AND XOR OR on A with v           (ex = mask of a & v, op = mask of a ^ v)
    H_INCDEC,    // INC / DEC of v (register or (HL)): op = 1 / 0xFF, ex = 0 / N|H
    H_ADD_HL,    // ADD HL,rr (BC DE HL)
    H_JUMP,      // JR / JP, conditional or not (target ready in imm16)
    H_CALL,      // CALL cc / CALL / RST
    H_RET,       // RET cc / RET / RETI
    H_PUSH,
    H_POP,
    H_ROT,       // CB rotates / shifts / SWAP on v, and RLCA RRCA RLA RRA (op bit 3: Z forced clear)
    H_BIT,       // BIT b,v
    H_RESSET,    // RES / SET b,v
    H_CPL,       // CPL
    H_IME,       // DI / EI (op = the new IME; PyBoy: EI takes effect at once)
    H_JPHL,      // JP HL
    H_RARE,      // everything else, by opcode (never executed by the fast loop)
    H_SLOW,      // not pre-decodable here (instruction straddles a bank boundary): decoded on the fly by the slow tick
    H__COUNT
};

#define PDF_IMM 0x0010u   // v = imm16
#define PDF_RD 0x0020u    // v = bus[addr]
#define PDF_RD16 0x0040u  // v |= bus[addr + 1] << 8
#define PDF_AIMM 0x0080u  // addr = imm16 (else the pair selected by asel)
#define PDF_ASP 0x0100u   // addr = SP
#define PDF_WR 0x0200u    // one deferred byte store of the handler's `wv` at addr
#define PDF_RETI 0x0400u  // RETI: IME = 1
#define PDF_MOV 0x0800u   // handler is H_MOV (tested as a flag so that the dispatch of plain moves is one predicate, not a switch level)
#define PDF_JUMP 0x1000u  // handler is H_JUMP
#define PDF_INCDEC 0x2000u  // handler is H_INCDEC
#define PDF_ARITH 0x4000u   // handler is H_ARITH

#define PD_H(x) ((x) & 0xFFu)
#define PD_OP(x) (((x) >> 8) & 0xFFu)
#define PD_EX(x) (((x) >> 16) & 0xFFu)
#define PD_CYC(x) ((x) >> 24)

#if defined(__VECTOR_TYPES_H__)  // cuda_runtime.h (or the host-simulation shim) has defined uint4
typedef uint4 pd_desc_t;
#else
struct pd_desc_t { uint32_t x, y, z, w; };
#endif

// base-table entries keep instance-independent fields; y = constant imm | len << 16 | immediate kind << 24
enum { PDK_CONST = 0, PDK_IMM8, PDK_IMM16, PDK_JR, PDK_LDH };
#define PD_BASE_LEN(y) (((y) >> 16) & 3u)
#define PD_BASE_KIND(y) ((y) >> 24)

// register index of the opcode encoding (B C D E H L - A) -> byte number inside bcde:hlaf
static inline uint32_t pd_reg_byte(uint32_t idx) { return ((idx & 4) ? 4u : 0u) + ((idx ^ 1u) & 3u); }

struct pd_builder {
    uint32_t h, cyc, ex, op, imm, len, kind, sel_lo, sel_hi, srcsel, flags, asel;
};
static inline void pd_b_init(pd_builder *b, uint32_t h, uint32_t len, uint32_t cyc) {
    b->h = h; b->cyc = cyc; b->ex = 0; b->op = 0; b->imm = 0; b->len = len; b->kind = PDK_CONST;
    b->sel_lo = 0x3210; b->sel_hi = 0x3210; b->srcsel = 0; b->flags = 0; b->asel = 0x5454;
}
// result byte j of rv -> register byte k (0..7)
static inline void pd_b_write(pd_builder *b, uint32_t k, uint32_t j) {
    uint32_t *sel = k < 4 ? &b->sel_lo : &b->sel_hi;
    uint32_t sh = (k & 3) * 4;
    *sel = (*sel & ~(0xFu << sh)) | ((4u + j) << sh);
}
static inline void pd_b_src_reg(pd_builder *b, uint32_t idx) { b->srcsel = pd_reg_byte(idx); }
static inline void pd_b_pair(pd_builder *b, uint32_t p) {  // BC DE HL as the address / 16-bit operand
    static const uint32_t s[3] = {0x1010, 0x3232, 0x5454};
    b->asel = s[p];
}
static inline void pd_b_write_pair(pd_builder *b, uint32_t p) {  // rv bytes 0,1 -> pair p (BC DE HL)
    pd_b_write(b, p * 2, 0);
    pd_b_write(b, p * 2 + 1, 1);
}
// Dense id of the (handler, operand flags) pair (gb_classes.inc), 0xFF for whatever the fast set does not know.
static inline uint32_t pd_class_id(uint32_t h, uint32_t flags) {
#if !defined(GB_NO_CLASS_LOOKUP)
#define GB_CLS(id, H, F) \
    if (h == (H) && flags == (F)) return id;
#include "gb_classes.inc"
#undef GB_CLS
#endif
    (void)h; (void)flags;
    return 0xFFu;
}
#define PD_CLASS(w) ((w) >> 24)
#define PD_NO_CLASS_W 0xFF000000u  // w of a control word without flags and outside every class (H_SLOW)

static inline pd_desc_t pd_b_done(const pd_builder *b) {
    pd_desc_t d;
    d.x = b->h | (b->op << 8) | (b->ex << 16) | (b->cyc << 24);
    d.y = (b->imm & 0xFFFFu) | (b->len << 16) | (b->kind << 24);
    d.z = b->sel_lo | (b->sel_hi << 16);
    d.w = (b->srcsel & 0xFu) | (b->flags & 0xFFF0u) | (b->h == H_MOV ? PDF_MOV : 0u) | (b->h == H_JUMP ? PDF_JUMP : 0u) |
          (b->h == H_INCDEC ? PDF_INCDEC : 0u) | (b->h == H_ARITH ? PDF_ARITH : 0u) | ((b->asel & 0xFFu) << 16);
    d.w |= pd_class_id(b->h, d.w & 0xFFF0u) << 24;
    return d;
}

// condition cc (NZ Z NC C) -> mask in op, expected flag value in ex
static inline void pd_b_cond(pd_builder *b, uint32_t cc, uint32_t taken_extra) {
    uint32_t mask = (cc & 2) ? 0x10u : 0x80u;
    b->op = mask;
    b->ex = taken_extra | ((cc & 1) ? mask : 0u);
}

// rotate / shift kind y = RLC RRC RL RR SLA SRA SWAP SRL -> `ex` of H_ROT: bit 0 shifts right, bits 1 / 2 / 3 shift in the
// bit that falls out / the carry flag / bit 7 (SRA), bit 4 swaps the nibbles instead
static inline uint32_t pd_rot_ex(uint32_t y) {
    static const uint32_t t[8] = {0x02, 0x03, 0x04, 0x05, 0x00, 0x09, 0x10, 0x01};
    return t[y & 7];
}

// ALU group y = ADD ADC SUB SBC AND XOR OR CP
static inline void pd_b_alu(pd_builder *b, uint32_t y) {
    if (y >= 4 && y <= 6) {
        b->h = H_LOGIC;
        b->ex = y == 5 ? 0u : 0xFFu;
        b->op = y == 4 ? 0u : 0xFFu;
    } else {
        b->h = H_ARITH;
        b->ex = (y == 2 || y == 3 || y == 7) ? 0xFFu : 0u;
        b->op = (y == 1 || y == 3) ? 1u : 0u;
    }
}

// Per-opcode base descriptors (256 base + 256 CB page).  Cycle counts: the pastraiser table PyBoy 1.6 uses.
static inline void pd_build_base(pd_desc_t *t) {
    for (uint32_t opc = 0; opc < 256; opc++) {
        const uint32_t x = opc >> 6, y = (opc >> 3) & 7, z = opc & 7, p = y >> 1, q = y & 1;
        pd_builder b;
        pd_b_init(&b, H_RARE, 1, 4);
        b.op = opc;
        if (x == 1) {
            if (opc != 0x76) {  // LD r,r' / LD r,(HL) / LD (HL),r
                pd_b_init(&b, H_MOV, 1, (y == 6 || z == 6) ? 8 : 4);
                if (z == 6) b.flags |= PDF_RD;
                else pd_b_src_reg(&b, z);
                if (y == 6) b.flags |= PDF_WR;
                else pd_b_write(&b, pd_reg_byte(y), 0);
            }
        } else if (x == 2) {  // ALU A,r / ALU A,(HL)
            pd_b_init(&b, H_ARITH, 1, z == 6 ? 8 : 4);
            pd_b_alu(&b, y);
            if (z == 6) b.flags |= PDF_RD;
            else pd_b_src_reg(&b, z);
            if (y != 7) pd_b_write(&b, 6, 0);  // CP keeps A
            pd_b_write(&b, 7, 1);
        } else if (x == 0) {
            switch (z) {
            case 0:
                if (y == 0) {
                    pd_b_init(&b, H_MOV, 1, 4);  // NOP: a move that writes nothing
                } else if (y == 1) {
                    b.len = 3; b.cyc = 20; b.kind = PDK_IMM16;  // LD (nn),SP (rare)
                } else if (y == 2) {
                    b.len = 2;  // STOP skips a byte (rare)
                } else {
                    pd_b_init(&b, H_JUMP, 2, y == 3 ? 12 : 8);
                    b.kind = PDK_JR;
                    if (y >= 4) pd_b_cond(&b, y & 3, 4);
                }
                break;
            case 1:
                if (q == 0) {
                    if (p == 3) { b.len = 3; b.cyc = 12; b.kind = PDK_IMM16; }  // LD SP,nn (rare)
                    else {
                        pd_b_init(&b, H_MOV, 3, 12);
                        b.kind = PDK_IMM16; b.flags |= PDF_IMM;
                        pd_b_write_pair(&b, p);
                    }
                } else {
                    if (p == 3) { b.cyc = 8; }  // ADD HL,SP (rare)
                    else {
                        pd_b_init(&b, H_ADD_HL, 1, 8);
                        pd_b_pair(&b, p);
                        pd_b_write(&b, 4, 0); pd_b_write(&b, 5, 1); pd_b_write(&b, 7, 3);
                    }
                }
                break;
            case 2: {  // LD (BC/DE/HL+/HL-),A and LD A,(BC/DE/HL+/HL-)
                pd_b_init(&b, p < 2 ? H_MOV : H_HLI, 1, 8);
                pd_b_pair(&b, p < 2 ? p : 2);
                if (p >= 2) b.op = p == 2 ? 1u : 0xFFu;
                if (q == 0) {
                    b.srcsel = 6; b.flags |= PDF_WR;
                    if (p >= 2) { pd_b_write(&b, 4, 0); pd_b_write(&b, 5, 1); }
                } else {
                    b.flags |= PDF_RD;
                    if (p >= 2) { pd_b_write(&b, 4, 0); pd_b_write(&b, 5, 1); pd_b_write(&b, 6, 2); }
                    else pd_b_write(&b, 6, 0);
                }
                break;
            }
            case 3:
                if (p == 3) { b.cyc = 8; }  // INC SP / DEC SP (rare)
                else {
                    pd_b_init(&b, H_HLI, 1, 8);
                    pd_b_pair(&b, p);
                    b.op = q ? 0xFFu : 1u;
                    pd_b_write_pair(&b, p);
                }
                break;
            case 4:
            case 5:
                pd_b_init(&b, H_INCDEC, 1, y == 6 ? 12 : 4);
                b.op = (z & 1) ? 0xFFu : 1u;
                b.ex = (z & 1) ? 0x60u : 0u;
                if (y == 6) b.flags |= PDF_RD | PDF_WR;
                else { pd_b_src_reg(&b, y); pd_b_write(&b, pd_reg_byte(y), 0); }
                pd_b_write(&b, 7, 1);
                break;
            case 6:
                pd_b_init(&b, H_MOV, 2, y == 6 ? 12 : 8);
                b.kind = PDK_IMM8; b.flags |= PDF_IMM;
                if (y == 6) b.flags |= PDF_WR;
                else pd_b_write(&b, pd_reg_byte(y), 0);
                break;
            default:
                if (y < 4) {  // RLCA RRCA RLA RRA: the CB rotate of A with Z forced clear
                    pd_b_init(&b, H_ROT, 1, 4);
                    b.op = y | 8u; b.ex = pd_rot_ex(y); b.srcsel = 6;
                    pd_b_write(&b, 6, 0); pd_b_write(&b, 7, 1);
                }
                else if (y == 5) {  // CPL
                    pd_b_init(&b, H_CPL, 1, 4);
                    pd_b_write(&b, 6, 0); pd_b_write(&b, 7, 1);
                }
                break;  // DAA SCF CCF: rare
            }
        } else {
            switch (z) {
            case 0:
                if (y < 4) {
                    pd_b_init(&b, H_RET, 1, 8);
                    b.flags |= PDF_RD | PDF_RD16 | PDF_ASP;
                    pd_b_cond(&b, y, 12);
                } else if (y == 4 || y == 6) {  // LDH (n),A / LDH A,(n)
                    pd_b_init(&b, H_MOV, 2, 12);
                    b.kind = PDK_LDH; b.flags |= PDF_AIMM;
                    if (y == 4) { b.srcsel = 6; b.flags |= PDF_WR; }
                    else { b.flags |= PDF_RD; pd_b_write(&b, 6, 0); }
                } else {
                    b.len = 2; b.cyc = y == 5 ? 16 : 12; b.kind = PDK_IMM8;  // ADD SP,e / LD HL,SP+e (rare)
                }
                break;
            case 1:
                if (q == 0) {
                    pd_b_init(&b, H_POP, 1, 12);
                    b.flags |= PDF_RD | PDF_RD16 | PDF_ASP;
                    b.op = p == 3 ? 0xF0u : 0xFFu;
                    if (p == 3) { pd_b_write(&b, 6, 1); pd_b_write(&b, 7, 0); }
                    else pd_b_write_pair(&b, p);
                } else if (p < 2) {
                    pd_b_init(&b, H_RET, 1, 16);
                    b.flags |= PDF_RD | PDF_RD16 | PDF_ASP | (p == 1 ? PDF_RETI : 0);
                } else if (p == 3) {
                    b.cyc = 8;  // LD SP,HL (rare)
                } else {
                    pd_b_init(&b, H_JPHL, 1, 4);
                }
                break;
            case 2:
                if (y < 4) {
                    pd_b_init(&b, H_JUMP, 3, 12);
                    b.kind = PDK_IMM16;
                    pd_b_cond(&b, y, 4);
                } else if (y == 4 || y == 6) {
                    b.cyc = 8;  // LD (FF00+C),A / LD A,(FF00+C) (rare handler: the address needs C at run time)
                } else {  // LD (nn),A / LD A,(nn)
                    pd_b_init(&b, H_MOV, 3, 16);
                    b.kind = PDK_IMM16; b.flags |= PDF_AIMM;
                    if (y == 5) { b.srcsel = 6; b.flags |= PDF_WR; }
                    else { b.flags |= PDF_RD; pd_b_write(&b, 6, 0); }
                }
                break;
            case 3:
                if (y == 0) { pd_b_init(&b, H_JUMP, 3, 16); b.kind = PDK_IMM16; }
                else if (y == 6 || y == 7) { pd_b_init(&b, H_IME, 1, 4); b.op = y & 1; }  // DI / EI
                break;  // CB prefix (own page), illegal: rare
            case 4:
                if (y < 4) {
                    pd_b_init(&b, H_CALL, 3, 12);
                    b.kind = PDK_IMM16;
                    pd_b_cond(&b, y, 12);
                }
                break;
            case 5:
                if (q == 0) {
                    pd_b_init(&b, H_PUSH, 1, 16);
                    if (p == 3) b.asel = 0x6767;  // F low, A high
                    else pd_b_pair(&b, p);
                } else if (p == 0) {
                    pd_b_init(&b, H_CALL, 3, 24);
                    b.kind = PDK_IMM16;
                }
                break;
            case 6:
                pd_b_init(&b, H_ARITH, 2, 8);
                pd_b_alu(&b, y);
                b.kind = PDK_IMM8; b.flags |= PDF_IMM;
                if (y != 7) pd_b_write(&b, 6, 0);
                pd_b_write(&b, 7, 1);
                break;
            default:  // RST: a CALL to a constant
                pd_b_init(&b, H_CALL, 1, 16);
                b.imm = y * 8;
                break;
            }
        }
        t[opc] = pd_b_done(&b);
    }
    for (uint32_t opc = 0; opc < 256; opc++) {
        const uint32_t x = opc >> 6, y = (opc >> 3) & 7, z = opc & 7;
        pd_builder b;
        pd_b_init(&b, x == 0 ? H_ROT : x == 1 ? H_BIT : H_RESSET, 2, z == 6 ? 16 : 8);
        if (z == 6) b.flags |= PDF_RD | (x == 1 ? 0 : PDF_WR);
        else pd_b_src_reg(&b, z);
        if (x == 0) {
            b.op = y; b.ex = pd_rot_ex(y);
            pd_b_write(&b, 7, 1);
        } else if (x == 1) {
            b.imm = 1u << y;
            pd_b_write(&b, 7, 1);
        } else {
            b.imm = x == 2 ? (0xFFu & ~(1u << y)) : (0xFFu | ((1u << y) << 8));  // and mask | or mask << 8
        }
        if (x != 1 && z != 6) pd_b_write(&b, pd_reg_byte(z), 0);
        t[256 + opc] = pd_b_done(&b);
    }
}

// instance fields: ins = opcode | op1 << 8 | op2 << 16 at address pc, b = its base entry
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline pd_desc_t pd_finish(pd_desc_t b, uint32_t ins, uint32_t pc) {
    const uint32_t imm16 = (ins >> 8) & 0xFFFFu, imm8 = imm16 & 0xFFu, len = PD_BASE_LEN(b.y);
    uint32_t imm = b.y & 0xFFFFu;
    switch (PD_BASE_KIND(b.y)) {
    case PDK_IMM8: imm = imm8; break;
    case PDK_IMM16: imm = imm16; break;
    case PDK_JR: imm = (pc + 2 + ((imm8 ^ 0x80u) - 0x80u)) & 0xFFFFu; break;
    case PDK_LDH: imm = 0xFF00u | imm8; break;
    default: break;
    }
    b.y = imm | (((pc + len) & 0xFFFFu) << 16);
    return b;
}
