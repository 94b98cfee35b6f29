// gb_predecode.h -- pre-decoded instruction descriptors (host + device).
//
// The cartridge ROM is shared by every env and never changes, so it is decoded ONCE per handle: for each
// ROM offset the library stores an 8-byte descriptor {handler, register fields, length, cycle counts,
// immediate, precomputed relative-jump target}.  The interpreter then fetches one 64-bit word per
// instruction from this L2-resident table and jumps straight to a specialised handler -- no opcode
// bit-field decoding, operand fetching or length/cycle computation on the hot path.  Code executed from
// RAM (the HRAM OAM-DMA stub) and the last bytes of each 16 KiB bank are decoded on the fly from the same
// per-opcode base table.
#pragma once
#include <stdint.h>

// descriptor word 0:  handler | taken_extra << 6 | Y << 8 | Z << 16 | cycles << 24   (byte-aligned: one PRMT / shift each)
//   Y  opcode bits 3-5: destination register / ALU op / bit index / condition (H_RARE: the whole opcode byte)
//   Z  opcode bits 0-2: source register
//   cycles  T-cycles, condition false / unconditional;  taken_extra  (T-cycles when the condition holds - cycles) / 4
// descriptor word 1:  imm16 | next_pc << 16   (H_JR: the branch target instead of imm16; CB page: opcode bits 6-7)
#define PD_H(x) ((x) & 63u)
#define PD_TAKEN_EXTRA(x) ((((x) >> 6) & 3u) * 4u)
#if defined(__CUDA_ARCH__)
#define PD_Y(x) __byte_perm((x), 0, 0x4441)
#define PD_Z(x) __byte_perm((x), 0, 0x4442)
#else
#define PD_Y(x) (((x) >> 8) & 0xFFu)
#define PD_Z(x) (((x) >> 16) & 0xFFu)
#endif
#define PD_CYC(x) ((x) >> 24)

enum {
    H_SLOW = 0,  // not pre-decodable here (instruction straddles a bank boundary): decode on the fly
    H_NOP, H_LD_R_R, H_LD_R_HL, H_LD_HL_R, H_LD_R_N, H_LD_HL_N, H_LD_A_RP, H_LD_RP_A,
    H_LDH_N_A, H_LDH_A_N, H_LD_C_A, H_LD_A_C, H_LD_NN_A, H_LD_A_NN,
    H_ALU_R, H_ALU_HL, H_ALU_N, H_INCDEC_R, H_INCDEC_HL, H_LD_RP_NN, H_INCDEC_RP, H_ADD_HL,
    H_JR, H_JP, H_CALL, H_RET, H_RETI, H_RST, H_PUSH, H_POP, H_CB_R, H_CB_HL, H_ROT_A, H_RARE,
    H__COUNT
};

// `len` travels in bits 28-29 of the BASE entry only (pd_split_len strips it); cycles <= 24 needs bits 24-28
static inline uint32_t pd_make(uint32_t h, uint32_t y, uint32_t z, uint32_t len, uint32_t cyc, uint32_t cyc2, uint32_t op) {
    if (h == H_RARE) y = op;
    return h | (((cyc2 - cyc) / 4) << 6) | (y << 8) | (z << 16) | (cyc << 24) | (len << 29);
}
#define PD_BASE_LEN(x) ((x) >> 29)
#define PD_BASE_WORD0(x) ((x) & 0x1FFFFFFFu)

// Per-opcode base descriptors (256 base + 256 CB page).  Cycle counts: the pastraiser table PyBoy 1.6 uses.
// For conditional control flow the Y field holds the condition: 0 = always, 4..7 = NZ Z NC C.
static inline void pd_build_base(uint32_t *t) {
    for (uint32_t op = 0; op < 256; op++) {
        uint32_t x = op >> 6, y = (op >> 3) & 7, z = op & 7, p = y >> 1, q = y & 1;
        uint32_t d = pd_make(H_RARE, y, z, 1, 4, 4, op);
        if (x == 1) {
            if (op != 0x76) d = z == 6 ? pd_make(H_LD_R_HL, y, z, 1, 8, 8, op) : y == 6 ? pd_make(H_LD_HL_R, y, z, 1, 8, 8, op) : pd_make(H_LD_R_R, y, z, 1, 4, 4, op);
        } else if (x == 2) {
            d = z == 6 ? pd_make(H_ALU_HL, y, z, 1, 8, 8, op) : pd_make(H_ALU_R, y, z, 1, 4, 4, op);
        } else if (x == 0) {
            switch (z) {
            case 0:
                if (y == 0) d = pd_make(H_NOP, 0, 0, 1, 4, 4, op);
                else if (y == 3) d = pd_make(H_JR, 0, 0, 2, 12, 12, op);
                else if (y >= 4) d = pd_make(H_JR, y, 0, 2, 8, 12, op);
                else if (y == 1) d = pd_make(H_RARE, y, z, 3, 20, 20, op);  // LD (nn),SP
                else if (y == 2) d = pd_make(H_RARE, y, z, 2, 4, 4, op);    // STOP skips a byte
                break;
            case 1: d = q == 0 ? pd_make(H_LD_RP_NN, y, z, 3, 12, 12, op) : pd_make(H_ADD_HL, y, z, 1, 8, 8, op); break;
            case 2: d = q == 0 ? pd_make(H_LD_RP_A, y, z, 1, 8, 8, op) : pd_make(H_LD_A_RP, y, z, 1, 8, 8, op); break;
            case 3: d = pd_make(H_INCDEC_RP, y, z, 1, 8, 8, op); break;
            case 4:
            case 5: d = y == 6 ? pd_make(H_INCDEC_HL, y, z, 1, 12, 12, op) : pd_make(H_INCDEC_R, y, z, 1, 4, 4, op); break;
            case 6: d = y == 6 ? pd_make(H_LD_HL_N, y, z, 2, 12, 12, op) : pd_make(H_LD_R_N, y, z, 2, 8, 8, op); break;
            default:
                if (y < 4) d = pd_make(H_ROT_A, y, z, 1, 4, 4, op);  // RLCA RRCA RLA RRA
                break;  // DAA CPL SCF CCF: rare
            }
        } else {
            switch (z) {
            case 0:
                if (y < 4) d = pd_make(H_RET, 4 + y, 0, 1, 8, 20, op);
                else if (y == 4) d = pd_make(H_LDH_N_A, y, z, 2, 12, 12, op);
                else if (y == 6) d = pd_make(H_LDH_A_N, y, z, 2, 12, 12, op);
                else if (y == 5) d = pd_make(H_RARE, y, z, 2, 16, 16, op);  // ADD SP,e
                else d = pd_make(H_RARE, y, z, 2, 12, 12, op);              // LD HL,SP+e
                break;
            case 1:
                if (q == 0) d = pd_make(H_POP, y, z, 1, 12, 12, op);
                else if (p == 0) d = pd_make(H_RET, 0, 0, 1, 16, 16, op);
                else if (p == 1) d = pd_make(H_RETI, 0, 0, 1, 16, 16, op);
                else if (p == 3) d = pd_make(H_RARE, y, z, 1, 8, 8, op);  // LD SP,HL
                break;  // JP HL: rare, 4 cycles
            case 2:
                if (y < 4) d = pd_make(H_JP, 4 + y, 0, 3, 12, 16, op);
                else if (y == 4) d = pd_make(H_LD_C_A, y, z, 1, 8, 8, op);
                else if (y == 5) d = pd_make(H_LD_NN_A, y, z, 3, 16, 16, op);
                else if (y == 6) d = pd_make(H_LD_A_C, y, z, 1, 8, 8, op);
                else d = pd_make(H_LD_A_NN, y, z, 3, 16, 16, op);
                break;
            case 3:
                if (y == 0) d = pd_make(H_JP, 0, 0, 3, 16, 16, op);
                break;  // CB prefix (own page), DI, EI, illegal: rare
            case 4:
                if (y < 4) d = pd_make(H_CALL, 4 + y, 0, 3, 12, 24, op);
                break;
            case 5:
                if (q == 0) d = pd_make(H_PUSH, y, z, 1, 16, 16, op);
                else if (p == 0) d = pd_make(H_CALL, 0, 0, 3, 24, 24, op);
                break;
            case 6: d = pd_make(H_ALU_N, y, z, 2, 8, 8, op); break;
            default: d = pd_make(H_RST, y, z, 1, 16, 16, op); break;
            }
        }
        t[op] = d;
    }
    for (uint32_t op = 0; op < 256; op++) {
        uint32_t y = (op >> 3) & 7, z = op & 7;
        t[256 + op] = z == 6 ? pd_make(H_CB_HL, y, z, 2, 16, 16, op) : pd_make(H_CB_R, y, z, 2, 8, 8, op);
    }
}
