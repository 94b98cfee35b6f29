// gb_device.cuh -- device-side Game Boy (DMG + MBC3) machine: SM83 interpreter, bus, timer, LCD
// state machine and scanline renderer, one env per thread.
//
// Replaces PyBoy 1.6.x as driven by /root/reference/pokegym/pyboy_binding.py:71-91 (24 x PyBoy.tick
// per env step) and every get/set_memory_value call of the reference wrapper.  Semantics follow the
// instruction-granular PyBoy model of SURVEY.md Appendix A (peripherals advance after each whole
// instruction, interrupt dispatch costs 0 cycles, HALT fast-forwards to the next LCD event, one
// TIMA increment and one LCD mode change per tick at most ...).  Structure is B200-first and shares
// nothing with the CPU oracle: registers are packed words decoded with bit-field arithmetic, all
// env state lives in word-interleaved HBM arrays (gb_layout.cuh), and the main loop is organised
// around LCD events so the 32 envs of a warp re-converge at every scanline boundary.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gb_layout.cuh"

#define FLAG_Z 0x80u
#define FLAG_N 0x40u
#define FLAG_H 0x20u
#define FLAG_C 0x10u

#define IRQ_VBLANK 0x01u
#define IRQ_STAT 0x02u
#define IRQ_TIMER 0x04u
#define IRQ_SERIAL 0x08u
#define IRQ_JOYPAD 0x10u

#define FRAME_CYCLES 70224u

struct Machine {
    // SM83 registers: C|B<<8|E<<16|D<<24 and L|H<<8|A<<16|F<<24 (pairs are native little-endian halves)
    uint32_t bcde, hlaf, sp, pc;
    uint32_t ime, halted, stopped, iq, fault, ie, iflag;
    // LCD
    uint32_t lcdc, stat, ly, lyc, scy, scx, wy, wx, bgp, obp0, obp1;
    uint32_t stat_mode, next_mode, disable_renderer, frame_done;
    uint32_t clock, target;
    // timer
    uint32_t div, tima, tma, tac, divc, timac;
    // MBC3
    uint32_t rombank, rambank, ram_en, memorymodel, rom_off;
    // joypad + renderer bookkeeping
    uint32_t joy_dir, joy_std, lp_dirty, blank_shade;
    int ly_window;
    uint32_t hdr;
    // statistics
    uint32_t n_instr, n_cycles;
    // memory (pointers already offset to this env's lane inside its tile)
    uint8_t *memb;   // plain RAM, byte i at memb[((i >> 2) << 7) | (i & 3)]
    uint8_t *cramb;  // cart RAM, same addressing
    uint32_t *fb;    // word w at fb[w * 32]
    uint2 *lp;       // scanline y at lp[y * 32]
    const uint8_t *rom;
    uint32_t rom_banks;
};

// ------------------------------------------------------------------------------- state load/store

__device__ __forceinline__ void machine_bind(Machine &m, const DevArrays &d, int tile, int lane) {
    m.memb = (uint8_t *)(d.mem + il_index(tile, MEM_WORDS, 0, lane));
    m.cramb = (uint8_t *)(d.cram + il_index(tile, CRAM_WORDS, 0, lane));
    m.fb = d.fb + il_index(tile, FB_WORDS, 0, lane);
    m.lp = (uint2 *)(d.lp + (size_t)tile * LP_WORDS * GB_TILE) + lane;
    m.rom = d.rom;
    m.rom_banks = d.rom_banks;
}

__device__ __forceinline__ void machine_set_rombank(Machine &m, uint32_t bank) {
    m.rombank = bank;
    m.rom_off = (bank % m.rom_banks) * 0x4000u - 0x4000u;  // rom[addr + rom_off] for 0x4000 <= addr < 0x8000
}

__device__ inline void machine_load(Machine &m, const DevArrays &d, int tile, int lane) {
    machine_bind(m, d, tile, lane);
    const uint32_t *r = d.regs + il_index(tile, R_WORDS, 0, lane);
    uint32_t w;
    m.bcde = r[R_BCDE * 32];
    m.hlaf = r[R_HLAF * 32];
    w = r[R_SPPC * 32];
    m.sp = w & 0xFFFF;
    m.pc = w >> 16;
    w = r[R_INT * 32];
    m.ime = w & 1; m.halted = (w >> 1) & 1; m.stopped = (w >> 2) & 1; m.iq = (w >> 3) & 1; m.fault = (w >> 4) & 1;
    m.ie = (w >> 8) & 0xFF; m.iflag = (w >> 16) & 0xFF;
    w = r[R_LCD0 * 32];
    m.lcdc = w & 0xFF; m.stat = (w >> 8) & 0xFF; m.ly = (w >> 16) & 0xFF; m.lyc = w >> 24;
    w = r[R_LCD1 * 32];
    m.scy = w & 0xFF; m.scx = (w >> 8) & 0xFF; m.wy = (w >> 16) & 0xFF; m.wx = w >> 24;
    w = r[R_LCD2 * 32];
    m.bgp = w & 0xFF; m.obp0 = (w >> 8) & 0xFF; m.obp1 = (w >> 16) & 0xFF;
    m.stat_mode = (w >> 24) & 3; m.next_mode = (w >> 26) & 3; m.disable_renderer = (w >> 28) & 1; m.frame_done = (w >> 29) & 1;
    m.clock = r[R_CLOCK * 32];
    m.target = r[R_TARGET * 32];
    w = r[R_TIMER * 32];
    m.div = w & 0xFF; m.tima = (w >> 8) & 0xFF; m.tma = (w >> 16) & 0xFF; m.tac = w >> 24;
    m.divc = r[R_DIVC * 32];
    m.timac = r[R_TIMAC * 32];
    w = r[R_MBC * 32];
    m.rambank = (w >> 8) & 0xFF; m.ram_en = (w >> 16) & 0xFF; m.memorymodel = w >> 24;
    machine_set_rombank(m, w & 0xFF);
    w = r[R_JOY * 32];
    m.joy_dir = w & 0xFF; m.joy_std = (w >> 8) & 0xFF; m.ly_window = (int)(int8_t)((w >> 16) & 0xFF); m.lp_dirty = w >> 24;
    m.hdr = r[R_HDR * 32];
    m.blank_shade = r[R_MISC * 32] & 0xFF;
    m.n_instr = 0;
    m.n_cycles = 0;
}

__device__ inline void machine_store(const Machine &m, const DevArrays &d, int tile, int lane) {
    uint32_t *r = d.regs + il_index(tile, R_WORDS, 0, lane);
    r[R_BCDE * 32] = m.bcde;
    r[R_HLAF * 32] = m.hlaf;
    r[R_SPPC * 32] = (m.sp & 0xFFFF) | (m.pc << 16);
    r[R_INT * 32] = m.ime | (m.halted << 1) | (m.stopped << 2) | (m.iq << 3) | (m.fault << 4) | (m.ie << 8) | (m.iflag << 16);
    r[R_LCD0 * 32] = m.lcdc | (m.stat << 8) | (m.ly << 16) | (m.lyc << 24);
    r[R_LCD1 * 32] = m.scy | (m.scx << 8) | (m.wy << 16) | (m.wx << 24);
    r[R_LCD2 * 32] = m.bgp | (m.obp0 << 8) | (m.obp1 << 16) | ((m.stat_mode | (m.next_mode << 2) | (m.disable_renderer << 4) | (m.frame_done << 5)) << 24);
    r[R_CLOCK * 32] = m.clock;
    r[R_TARGET * 32] = m.target;
    r[R_TIMER * 32] = m.div | (m.tima << 8) | (m.tma << 16) | (m.tac << 24);
    r[R_DIVC * 32] = m.divc;
    r[R_TIMAC * 32] = m.timac;
    r[R_MBC * 32] = m.rombank | (m.rambank << 8) | (m.ram_en << 16) | (m.memorymodel << 24);
    r[R_JOY * 32] = m.joy_dir | (m.joy_std << 8) | (((uint32_t)m.ly_window & 0xFF) << 16) | (m.lp_dirty << 24);
    r[R_HDR * 32] = m.hdr;
    r[R_MISC * 32] = m.blank_shade;
}

// ----------------------------------------------------------------------------------- plain memory

__device__ __forceinline__ uint32_t mem_rd(const Machine &m, uint32_t i) { return m.memb[((i >> 2) << 7) | (i & 3)]; }
__device__ __forceinline__ void mem_wr(Machine &m, uint32_t i, uint32_t v) { m.memb[((i >> 2) << 7) | (i & 3)] = (uint8_t)v; }
__device__ __forceinline__ uint32_t mem_rd_word(const Machine &m, uint32_t w) { return ((const uint32_t *)m.memb)[w << 5]; }
__device__ __forceinline__ void mem_wr_word(Machine &m, uint32_t w, uint32_t v) { ((uint32_t *)m.memb)[w << 5] = v; }
__device__ __forceinline__ uint32_t vram_rd16(const Machine &m, uint32_t i) {  // i even: both bytes share a word
    return *(const uint16_t *)(m.memb + (((i >> 2) << 7) | (i & 2)));
}

// ------------------------------------------------------------------------------------------ joypad

__device__ __forceinline__ uint32_t joypad_pull(const Machine &m, uint32_t v) {  // Interaction.pull
    uint32_t p14 = (v >> 4) & 1, p15 = (v >> 5) & 1, b = (v | 0xCF) & 0xFF;
    if (p14 != p15) b &= p14 ? m.joy_std : m.joy_dir;  // only one group selected: byte &= 4-bit nibble
    return b;
}

__device__ __forceinline__ void joypad_event(Machine &m, int button, int pressed) {  // Interaction.key_event
    uint32_t bit = 1u << (button & 3);
    uint32_t &reg = (button < 4) ? m.joy_dir : m.joy_std;
    uint32_t before = reg;
    reg = pressed ? (reg & ~bit) : (reg | bit);
    if ((before ^ reg) & before) m.iflag |= IRQ_JOYPAD;
}

// --------------------------------------------------------------------------------------------- LCD

__device__ __forceinline__ uint32_t stat_set_mode(Machine &m, uint32_t mode) {
    if (m.stat_mode == mode) return 0;
    m.stat_mode = mode;
    m.stat = (m.stat & 0xFC) | mode;
    return (mode != 3 && ((m.stat >> (mode + 3)) & 1)) ? IRQ_STAT : 0;
}

__device__ __forceinline__ uint32_t stat_update_lyc(Machine &m) {
    if (m.lyc == m.ly) {
        m.stat |= 0x04;
        return (m.stat & 0x40) ? IRQ_STAT : 0;
    }
    m.stat &= 0xFB;
    return 0;
}

__device__ __forceinline__ void lcd_set_lcdc(Machine &m, uint32_t v) {
    if ((v ^ m.lcdc) & 0x10) m.lp_dirty = 144;
    m.lcdc = v;
    if (!(v & 0x80)) {
        m.clock = 0;
        m.target = FRAME_CYCLES;
        stat_set_mode(m, 0);
        m.next_mode = 2;
        m.ly = 0;
        m.lp_dirty = 144;  // line order restarts after an LCD-off period
    }
}

// 2-bit shade for colour index idx through palette register pal
__device__ __forceinline__ uint32_t pal_shade(uint32_t pal, uint32_t idx) { return (pal >> (idx * 2)) & 3; }

// spread the low 8 bits of x to the even bit positions of a 16-bit value
__device__ __forceinline__ uint32_t spread8(uint32_t x) {
    x = (x | (x << 4)) & 0x0F0Fu;
    x = (x | (x << 2)) & 0x3333u;
    x = (x | (x << 1)) & 0x5555u;
    return x;
}

// One tile row (two bit-planes, leftmost pixel = bit 7) -> 8 pixels x 2-bit colour index, leftmost at bit 0
__device__ __forceinline__ uint32_t tile_row_indices(uint32_t b1, uint32_t b2) {
    uint32_t r1 = __brev(b1) >> 24, r2 = __brev(b2) >> 24;
    return spread8(r1) | (spread8(r2) << 1);
}

// map 8 packed 2-bit colour indices through a palette register -> 8 packed 2-bit shades
__device__ __forceinline__ uint32_t apply_palette16(uint32_t idx16, uint32_t pal) {
    uint32_t lo = idx16 & 0x5555u, hi = (idx16 >> 1) & 0x5555u;
    uint32_t m0 = ~lo & ~hi & 0x5555u, m1 = lo & ~hi, m2 = ~lo & hi, m3 = lo & hi;  // one bit per pixel at even positions
    uint32_t s = 0;
    s |= m0 * (pal & 3);
    s |= m1 * ((pal >> 2) & 3);
    s |= m2 * ((pal >> 4) & 3);
    s |= m3 * ((pal >> 6) & 3);
    return s;
}

// Renderer.scanline + Renderer.scanline_sprites for line y.  `line` is this thread's 10-word
// scratch in shared memory (stride `ls` words), `keys` its 10-entry sprite sort scratch.
__device__ inline void render_line(Machine &m, uint32_t y, uint32_t *line, uint32_t *keys, uint32_t ls) {
    const uint32_t lcdc = m.lcdc;
    const int wx = (int)m.wx - 7, wy = (int)m.wy;
    const bool win_line = (lcdc & 0x20) && wy <= (int)y && wx < 160;
    if (win_line) m.ly_window += 1;
    const uint32_t tds = lcdc & 0x10;
    const int wstart = win_line ? (wx > 0 ? wx : 0) : 160;  // first screen x covered by the window

    // ---- background layer, x in [0, wstart)
    if (wstart > 0) {
        if (lcdc & 0x01) {
            const uint32_t map_base = (lcdc & 0x08) ? 0x1C00u : 0x1800u;
            const uint32_t row = map_base + ((((y + m.scy) >> 3) << 5) & 0x3FFu);
            const uint32_t fine_y = (y + m.scy) & 7;
            uint32_t col = m.scx >> 3;
            uint64_t acc = 0;
            int nbits = -(int)((m.scx & 7) * 2);  // drop the leftmost (scx & 7) pixels of the first tile
            uint32_t out = 0;
            const uint32_t nwords = ((uint32_t)wstart + 15) >> 4;
            while (out < nwords) {
                uint32_t t = mem_rd(m, MEM_VRAM + row + (col & 31));
                col++;
                if (!tds) t = (t ^ 0x80) + 128;
                uint32_t rowdata = vram_rd16(m, MEM_VRAM + t * 16 + fine_y * 2);
                uint32_t px = apply_palette16(tile_row_indices(rowdata & 0xFF, rowdata >> 8), m.bgp);
                if (nbits < 0) {
                    acc = (uint64_t)(px >> (uint32_t)(-nbits));
                    nbits += 16;
                } else {
                    acc |= (uint64_t)px << nbits;
                    nbits += 16;
                }
                if (nbits >= 32) {
                    line[out * ls] = (uint32_t)acc;
                    out++;
                    acc >>= 32;
                    nbits -= 32;
                }
            }
        } else {
            uint32_t fill = pal_shade(m.bgp, 0) * 0x55555555u;
            for (uint32_t k = 0; k < ((uint32_t)wstart + 15) >> 4; k++) line[k * ls] = fill;
        }
    }
    // ---- window layer, x in [wstart, 160)
    if (win_line) {
        const uint32_t map_base = (lcdc & 0x40) ? 0x1C00u : 0x1800u;
        const uint32_t lw = (uint32_t)m.ly_window;
        const uint32_t row = map_base + (((lw >> 3) << 5) & 0x3FFu);
        const uint32_t fine_y = lw & 7;
        uint32_t first = (uint32_t)(wstart - wx);  // window pixel index shown at x = wstart
        uint32_t col = first >> 3;
        uint32_t out = (uint32_t)wstart >> 4;
        uint32_t lead = ((uint32_t)wstart & 15) * 2;  // bits of word `out` that belong to the background
        uint64_t acc = lead ? (uint64_t)(line[out * ls] & ((1u << lead) - 1)) : 0;
        int nbits = (int)lead - (int)((first & 7) * 2);
        bool first_tile = true;
        while (out < FB_LINE_WORDS) {
            uint32_t t = mem_rd(m, MEM_VRAM + row + (col & 31));
            col++;
            if (!tds) t = (t ^ 0x80) + 128;
            uint32_t rowdata = vram_rd16(m, MEM_VRAM + t * 16 + fine_y * 2);
            uint32_t px = apply_palette16(tile_row_indices(rowdata & 0xFF, rowdata >> 8), m.bgp);
            if (first_tile) {
                px >>= (first & 7) * 2;
                acc |= (uint64_t)px << lead;
                nbits = (int)lead + 16 - (int)((first & 7) * 2);
                first_tile = false;
            } else {
                acc |= (uint64_t)px << nbits;
                nbits += 16;
            }
            if (nbits >= 32) {
                line[out * ls] = (uint32_t)acc;
                out++;
                acc >>= 32;
                nbits -= 32;
            }
        }
    }
    if (y == 143) m.ly_window = -1;

    // ---- sprites
    if (lcdc & 0x02) {
        const int height = (lcdc & 0x04) ? 16 : 8;
        int count = 0;
        for (uint32_t n = 0; n < 40 && count < 10; n++) {
            uint32_t e = mem_rd_word(m, (MEM_HI >> 2) + n);  // Y | X<<8 | tile<<16 | attr<<24
            int sy = (int)(e & 0xFF) - 16;
            if (sy <= (int)y && (int)y < sy + height) {
                int sx = (int)((e >> 8) & 0xFF) - 8;
                // insertion into ascending (x, n) order; key = (sx + 8) << 8 | n keeps it unsigned
                uint32_t key = ((uint32_t)(sx + 8) << 8) | n;
                int j = count - 1;
                while (j >= 0 && keys[j * ls] > key) {
                    keys[(j + 1) * ls] = keys[j * ls];
                    j--;
                }
                keys[(j + 1) * ls] = key;
                count++;
            }
        }
        for (int i = count - 1; i >= 0; i--) {  // lowest priority first
            uint32_t n = keys[i * ls] & 0xFF;
            uint32_t e = mem_rd_word(m, (MEM_HI >> 2) + n);
            int sy = (int)(e & 0xFF) - 16, sx = (int)((e >> 8) & 0xFF) - 8;
            uint32_t tile = (e >> 16) & 0xFF, attr = e >> 24;
            if (height == 16) tile &= 0xFE;
            int dy = (int)y - sy;
            uint32_t yy = (attr & 0x40) ? (uint32_t)(height - dy - 1) : (uint32_t)dy;
            uint32_t rowdata = vram_rd16(m, MEM_VRAM + tile * 16 + yy * 2);
            uint32_t b1 = rowdata & 0xFF, b2 = rowdata >> 8;
            if (attr & 0x20) {  // x flip: leftmost pixel = bit 0
                b1 = __brev(b1) >> 24;
                b2 = __brev(b2) >> 24;
            }
            uint32_t idx16 = tile_row_indices(b1, b2);
            uint32_t opaque = (idx16 | (idx16 >> 1)) & 0x5555u;  // 1 per non-transparent pixel (even bit)
            uint32_t shades = apply_palette16(idx16, (attr & 0x10) ? m.obp1 : m.obp0);
            // clip to the screen
            if (sx < 0) {
                uint32_t cut = (uint32_t)(-sx) * 2;
                opaque >>= cut;
                shades >>= cut;
                sx = 0;
            }
            if (sx >= 160 || opaque == 0) continue;
            uint32_t w = (uint32_t)sx >> 4, sh = ((uint32_t)sx & 15) * 2;
            uint64_t cur = line[w * ls];
            if (w + 1 < FB_LINE_WORDS) cur |= (uint64_t)line[(w + 1) * ls] << 32;
            uint64_t opq = (uint64_t)opaque << sh;
            if (w + 1 >= FB_LINE_WORDS) opq &= 0xFFFFFFFFull;  // pixels beyond x = 159
            if (attr & 0x80) {  // OBJ behind BG: only where the current pixel is white (flag == shade 0)
                uint64_t white = ~(cur | (cur >> 1)) & 0x5555555555555555ull;
                opq &= white;
            }
            uint64_t mask = opq | (opq << 1);
            cur = (cur & ~mask) | (((uint64_t)shades << sh) & mask);
            line[w * ls] = (uint32_t)cur;
            if (w + 1 < FB_LINE_WORDS) line[(w + 1) * ls] = (uint32_t)(cur >> 32);
        }
    }
    // ---- commit to the framebuffer (coalesced across the warp)
#pragma unroll
    for (uint32_t k = 0; k < FB_LINE_WORDS; k++) m.fb[(y * FB_LINE_WORDS + k) << 5] = line[k * ls];
    m.blank_shade = 0xFF;
}

__device__ inline void lcd_blank_screen(Machine &m) {
    uint32_t shade = pal_shade(m.bgp, 0);
    if (m.blank_shade == shade) return;  // already uniformly this shade: the refill would be a no-op
    uint32_t fill = shade * 0x55555555u;
    for (uint32_t k = 0; k < FB_WORDS; k++) m.fb[k << 5] = fill;
    m.blank_shade = shade;
}

// LCD.tick after `clock` has been advanced and found >= target (LCD on) or >= FRAME_CYCLES (LCD off).
__device__ inline void lcd_event(Machine &m, uint32_t *line, uint32_t *keys, uint32_t ls) {
    if (m.lcdc & 0x80) {
        uint32_t irq = stat_set_mode(m, m.next_mode);
        switch (m.stat_mode) {
        case 2:
            if (m.ly == 153) {
                m.ly = 0;
                m.clock %= FRAME_CYCLES;
                m.target %= FRAME_CYCLES;
            } else {
                m.ly = (m.ly + 1) & 0xFF;
            }
            m.target += 80;
            m.next_mode = 3;
            irq |= stat_update_lyc(m);
            break;
        case 3:
            m.target += 170;
            m.next_mode = 0;
            break;
        case 0:
            m.target += 206;
            if (m.ly < 144) {
                if (m.lp_dirty) {  // Renderer._scanlineparameters[y]
                    uint2 v = make_uint2(m.scx | (m.scy << 8) | (m.wx << 16) | (m.wy << 24), m.lcdc);
                    m.lp[m.ly << 5] = v;
                    m.lp_dirty--;
                }
                if (!m.disable_renderer) render_line(m, m.ly, line, keys, ls);
            }
            m.next_mode = (m.ly < 143) ? 2 : 1;
            break;
        default:
            m.target += 456;
            m.next_mode = 1;
            m.ly = (m.ly + 1) & 0xFF;
            irq |= stat_update_lyc(m);
            if (m.ly == 144) {
                irq |= IRQ_VBLANK;
                m.frame_done = 1;
            }
            if (m.ly == 153) m.next_mode = 2;
            break;
        }
        m.iflag |= irq;
    } else {
        m.frame_done = 1;
        m.clock %= FRAME_CYCLES;
        lcd_blank_screen(m);
    }
}

// ------------------------------------------------------------------------------------------- timer

__device__ __forceinline__ uint32_t timer_divider(uint32_t tac) {
    return (tac & 3) == 0 ? 1024u : (4u << ((tac & 3) * 2));  // 1024, 16, 64, 256
}

__device__ __forceinline__ void timer_tick(Machine &m, uint32_t cycles) {
    m.divc += cycles;
    m.div = (m.div + (m.divc >> 8)) & 0xFF;
    m.divc &= 0xFF;
    if (m.tac & 4) {
        m.timac += cycles;
        uint32_t dv = timer_divider(m.tac);
        if (m.timac >= dv) {
            m.timac -= dv;
            if (m.tima == 0xFF) {
                m.tima = m.tma;
                m.iflag |= IRQ_TIMER;
            } else {
                m.tima += 1;
            }
        }
    }
}

__device__ __forceinline__ int timer_cycles_to_interrupt(const Machine &m) {
    if (!(m.tac & 4)) return 1 << 16;
    return (int)((0x100 - m.tima) * timer_divider(m.tac)) - (int)m.timac;
}

// --------------------------------------------------------------------------------------------- bus
// Two flavours of every bus function:
//   *_full  : everything inline.  Used by the interpreter, which funnels ALL of an instruction's data
//             accesses through one read site and one write site (cpu_step), so the big IO switch is
//             instantiated once and `Machine` never has its address taken (it stays in registers).
//   bus_read / bus_write : small inline fast path + out-of-line slow path, for the wrapper kernels and
//             the debug bus access, where call-site count matters more than register residency.

#define BUS_READ_BODY                                                                        \
    if (a < 0xFF00) {                                                                        \
        if (a >= 0xFE00) return mem_rd(m, MEM_HI + (a - 0xFE00));                            \
        if (a >= 0xC000) return mem_rd(m, MEM_WRAM + (a & 0x1FFF)); /* WRAM + echo */       \
        if (a >= 0xA000) {                                                                   \
            if (!m.ram_en) return 0xFF;                                                      \
            uint32_t i = (m.rambank & 3) * 0x2000u + (a - 0xA000);                           \
            return m.cramb[((i >> 2) << 7) | (i & 3)];                                       \
        }                                                                                    \
        return mem_rd(m, MEM_VRAM + (a - 0x8000));                                           \
    }                                                                                        \
    if (a >= 0xFF80 && a < 0xFFFF) return mem_rd(m, MEM_HI + (a - 0xFE00)); /* HRAM */       \
    switch (a) {                                                                             \
    case 0xFF04: return m.div;                                                               \
    case 0xFF05: return m.tima;                                                              \
    case 0xFF06: return m.tma;                                                               \
    case 0xFF07: return m.tac;                                                               \
    case 0xFF0F: return m.iflag;                                                             \
    case 0xFF40: return m.lcdc;                                                              \
    case 0xFF41: return m.stat;                                                              \
    case 0xFF42: return m.scy;                                                               \
    case 0xFF43: return m.scx;                                                               \
    case 0xFF44: return m.ly;                                                                \
    case 0xFF45: return m.lyc;                                                               \
    case 0xFF46: return 0;                                                                   \
    case 0xFF47: return m.bgp;                                                               \
    case 0xFF48: return m.obp0;                                                              \
    case 0xFF49: return m.obp1;                                                              \
    case 0xFF4A: return m.wy;                                                                \
    case 0xFF4B: return m.wx;                                                                \
    case 0xFFFF: return m.ie;                                                                \
    default:                                                                                 \
        if (a >= 0xFF10 && a < 0xFF40) return 0; /* sound disabled (pokegym default) */      \
        return mem_rd(m, MEM_HI + (a - 0xFE00));                                             \
    }

__device__ __forceinline__ uint32_t bus_read_full(Machine &m, uint32_t a) {  // Motherboard.getitem
    if (a < 0x8000) return __ldg(m.rom + (a < 0x4000 ? a : a + m.rom_off));
    BUS_READ_BODY
}
__device__ __noinline__ uint32_t bus_read_slow(Machine &m, uint32_t a) { BUS_READ_BODY }
__device__ __forceinline__ uint32_t bus_read(Machine &m, uint32_t a) {
    if (a < 0x8000) return __ldg(m.rom + (a < 0x4000 ? a : a + m.rom_off));
    if (a >= 0xC000 && a < 0xE000) return mem_rd(m, MEM_WRAM + (a - 0xC000));
    return bus_read_slow(m, a);
}

// Motherboard.transfer_DMA: instantaneous copy of 0xA0 bytes to OAM
__device__ __forceinline__ void oam_dma(Machine &m, uint32_t page);
__device__ __forceinline__ uint32_t bus_read_full(Machine &m, uint32_t a);
__device__ __forceinline__ void oam_dma(Machine &m, uint32_t page) {
    uint32_t src = page << 8;
    bool plain = (page >= 0x80 && page < 0xA0) || (page >= 0xC0 && page < 0xFE);
    if (plain) {  // word copy inside the plain-RAM array (src is 256-byte aligned)
        uint32_t base = page < 0xA0 ? (MEM_VRAM + (src - 0x8000)) : (MEM_WRAM + (src & 0x1FFF));
        for (uint32_t k = 0; k < 40; k++) mem_wr_word(m, (MEM_HI >> 2) + k, mem_rd_word(m, (base >> 2) + k));
    } else {
        for (uint32_t k = 0; k < 160; k++) mem_wr(m, MEM_HI + k, bus_read_full(m, (src + k) & 0xFFFF));
    }
}

#define BUS_WRITE_BODY                                                                                    \
    v &= 0xFF;                                                                                            \
    if (a >= 0xC000 && a < 0xFE00) { mem_wr(m, MEM_WRAM + (a & 0x1FFF), v); return; }                     \
    if (a < 0x8000) { /* MBC3 registers */                                                                \
        if (a < 0x2000) {                                                                                 \
            if ((v & 0x0F) == 0x0A) m.ram_en = 1;                                                         \
            else if (v == 0) m.ram_en = 0; /* PyBoy: any other value leaves the latch untouched */        \
        } else if (a < 0x4000) {                                                                          \
            v &= 0x7F;                                                                                    \
            machine_set_rombank(m, v ? v : 1);                                                            \
        } else if (a < 0x6000) {                                                                          \
            m.rambank = v;                                                                                \
        }                                                                                                 \
        return;                                                                                           \
    }                                                                                                     \
    if (a < 0xA000) { mem_wr(m, MEM_VRAM + (a - 0x8000), v); return; }                                    \
    if (a < 0xC000) {                                                                                     \
        if (m.ram_en && m.rambank <= 3) {                                                                 \
            uint32_t i = m.rambank * 0x2000u + (a - 0xA000);                                              \
            m.cramb[((i >> 2) << 7) | (i & 3)] = (uint8_t)v;                                              \
        }                                                                                                 \
        return;                                                                                           \
    }                                                                                                     \
    if (a < 0xFF00 || (a >= 0xFF80 && a < 0xFFFF)) { mem_wr(m, MEM_HI + (a - 0xFE00), v); return; }       \
    switch (a) {                                                                                          \
    case 0xFF00: mem_wr(m, MEM_HI + 0x100, joypad_pull(m, v)); break;                                     \
    case 0xFF04: m.div = 0; m.divc = 0; m.timac = 0; break;                                               \
    case 0xFF05: m.tima = v; break;                                                                       \
    case 0xFF06: m.tma = v; break;                                                                        \
    case 0xFF07: m.tac = v & 7; break;                                                                    \
    case 0xFF0F: m.iflag = v; break;                                                                      \
    case 0xFF40: lcd_set_lcdc(m, v); break;                                                               \
    case 0xFF41: m.stat = (m.stat & 0x87) | (v & 0x78); break;                                            \
    case 0xFF42: if (v != m.scy) m.lp_dirty = 144; m.scy = v; break;                                      \
    case 0xFF43: if (v != m.scx) m.lp_dirty = 144; m.scx = v; break;                                      \
    case 0xFF44: m.ly = v; m.lp_dirty = 144; break; /* PyBoy lets LY be written */                        \
    case 0xFF45: m.lyc = v; break;                                                                        \
    case 0xFF46: oam_dma(m, v); break;                                                                    \
    case 0xFF47: m.bgp = v; break;                                                                        \
    case 0xFF48: m.obp0 = v; break;                                                                       \
    case 0xFF49: m.obp1 = v; break;                                                                       \
    case 0xFF4A: if (v != m.wy) m.lp_dirty = 144; m.wy = v; break;                                        \
    case 0xFF4B: if (v != m.wx) m.lp_dirty = 144; m.wx = v; break;                                        \
    case 0xFFFF: m.ie = v; break;                                                                         \
    default:                                                                                              \
        if (a >= 0xFF10 && a < 0xFF40) break; /* sound disabled: writes dropped */                        \
        mem_wr(m, MEM_HI + (a - 0xFE00), v);                                                              \
        break;                                                                                            \
    }

__device__ __forceinline__ void bus_write_full(Machine &m, uint32_t a, uint32_t v) { BUS_WRITE_BODY }  // Motherboard.setitem
__device__ __noinline__ void bus_write_slow(Machine &m, uint32_t a, uint32_t v) { BUS_WRITE_BODY }
__device__ __forceinline__ void bus_write(Machine &m, uint32_t a, uint32_t v) {
    if (a >= 0xC000 && a < 0xE000) mem_wr(m, MEM_WRAM + (a - 0xC000), v);
    else bus_write_slow(m, a, v);
}

// ------------------------------------------------------------------------------------------- SM83

__device__ __forceinline__ uint32_t reg8(const Machine &m, uint32_t idx) {  // B C D E H L - A
    uint32_t w = (idx & 4) ? m.hlaf : m.bcde;
    return (w >> (((idx ^ 1) & 3) * 8)) & 0xFF;
}
__device__ __forceinline__ void set_reg8(Machine &m, uint32_t idx, uint32_t v) {
    uint32_t sh = ((idx ^ 1) & 3) * 8, mask = 0xFFu << sh;
    if (idx & 4) m.hlaf = (m.hlaf & ~mask) | ((v & 0xFF) << sh);
    else m.bcde = (m.bcde & ~mask) | ((v & 0xFF) << sh);
}
__device__ __forceinline__ uint32_t reg_a(const Machine &m) { return (m.hlaf >> 16) & 0xFF; }
__device__ __forceinline__ uint32_t reg_f(const Machine &m) { return m.hlaf >> 24; }
__device__ __forceinline__ uint32_t reg_hl(const Machine &m) { return m.hlaf & 0xFFFF; }
__device__ __forceinline__ void set_a(Machine &m, uint32_t v) { m.hlaf = (m.hlaf & 0xFF00FFFFu) | ((v & 0xFF) << 16); }
__device__ __forceinline__ void set_f(Machine &m, uint32_t v) { m.hlaf = (m.hlaf & 0x00FFFFFFu) | (v << 24); }
__device__ __forceinline__ void set_af(Machine &m, uint32_t a, uint32_t f) { m.hlaf = (m.hlaf & 0xFFFFu) | ((a & 0xFF) << 16) | (f << 24); }
__device__ __forceinline__ void set_hl(Machine &m, uint32_t v) { m.hlaf = (m.hlaf & 0xFFFF0000u) | (v & 0xFFFF); }
__device__ __forceinline__ uint32_t reg_pair(const Machine &m, uint32_t p) {  // BC DE HL SP
    return p == 0 ? (m.bcde & 0xFFFF) : p == 1 ? (m.bcde >> 16) : p == 2 ? (m.hlaf & 0xFFFF) : m.sp;
}
__device__ __forceinline__ void set_reg_pair(Machine &m, uint32_t p, uint32_t v) {
    v &= 0xFFFF;
    if (p == 0) m.bcde = (m.bcde & 0xFFFF0000u) | v;
    else if (p == 1) m.bcde = (m.bcde & 0xFFFFu) | (v << 16);
    else if (p == 2) m.hlaf = (m.hlaf & 0xFFFF0000u) | v;
    else m.sp = v;
}
__device__ __forceinline__ bool condition(const Machine &m, uint32_t cc) {  // NZ Z NC C
    uint32_t f = reg_f(m);
    uint32_t bit = (cc & 2) ? (f & FLAG_C) : (f & FLAG_Z);
    return (bit != 0) == ((cc & 1) != 0);
}

// 8-bit ALU group (ADD ADC SUB SBC AND XOR OR CP) on A with operand v
__device__ __forceinline__ void alu8(Machine &m, uint32_t op, uint32_t v) {
    uint32_t a = reg_a(m), f = reg_f(m), carry = (f >> 4) & 1, res, nf;
    if (op < 4 || op == 7) {
        bool sub = (op & 2) || op == 7;
        uint32_t cin = (op & 1) && op != 7 ? carry : 0;
        if (!sub) {
            res = a + v + cin;
            nf = (((a & 0xF) + (v & 0xF) + cin) > 0xF ? FLAG_H : 0) | (res > 0xFF ? FLAG_C : 0);
        } else {
            res = a - v - cin;
            nf = FLAG_N | (((a & 0xF) < (v & 0xF) + cin) ? FLAG_H : 0) | ((a < v + cin) ? FLAG_C : 0);
        }
        res &= 0xFF;
        if (res == 0) nf |= FLAG_Z;
        if (op == 7) res = a;
    } else {
        res = op == 4 ? (a & v) : op == 5 ? (a ^ v) : (a | v);
        nf = (op == 4 ? FLAG_H : 0) | (res == 0 ? FLAG_Z : 0);
    }
    set_af(m, res, nf);
}

// Instruction fetch: opcode plus its two possible operand bytes as one little-endian word.  Code almost
// always runs from cartridge ROM, where the three bytes come from two aligned 32-bit read-only loads
// (the ROM image is padded, so the second load is always in bounds); RAM code takes the byte path.
__device__ __forceinline__ uint32_t fetch3(Machine &m, uint32_t pc) {
    if (pc < 0x7FFD && (pc & 0x3FFF) < 0x3FFD) {  // the whole instruction lies inside one ROM bank window
        uint32_t addr = pc < 0x4000 ? pc : pc + m.rom_off;
        const uint32_t *w = (const uint32_t *)(m.rom + (addr & ~3u));
        return __funnelshift_r(__ldg(w), __ldg(w + 1), (addr & 3) * 8);
    }
    uint32_t ins = 0;
    for (uint32_t i = 0; i < 3; i++) ins |= bus_read_full(m, (pc + i) & 0xFFFF) << (8 * i);
    return ins;
}

// CPU.tick: interrupt check, HALT handling, one instruction.  Returns T-cycles (pastraiser table as used
// by PyBoy; interrupt dispatch costs 0).  Every data access of the instruction goes through ONE read
// site (up to two consecutive bytes: operand or 16-bit pop) and ONE write site (up to two bytes: operand,
// 16-bit push or LD (nn),SP), so threads executing different opcodes still share the memory code.
__device__ __forceinline__ uint32_t cpu_step(Machine &m) {
    uint32_t wn = 0, w0a = 0, w0v = 0, w1a = 0, w1v = 0;  // deferred bus writes, issued in order w0, w1
    uint32_t cycles = 0;
#define PUSH16(val)                                  \
    do {                                             \
        uint32_t _v = (val);                         \
        w0a = (m.sp - 1) & 0xFFFF; w0v = _v >> 8;    \
        w1a = (m.sp - 2) & 0xFFFF; w1v = _v & 0xFF;  \
        wn = 2;                                      \
        m.sp = (m.sp - 2) & 0xFFFF;                  \
    } while (0)
#define WRITE8(addr, val) do { w0a = (addr) & 0xFFFF; w0v = (val); wn = 1; } while (0)

    bool execute = true;
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
            execute = false;  // PyBoy charges no cycles for the dispatch
        }
    } else if (m.halted) {  // debugger-only path in PyBoy: halted with a queued interrupt
        m.halted = 0;
        m.pc = (m.pc + 1) & 0xFFFF;
    }
    if (execute && m.halted) return 4;
    if (execute) {
        const uint32_t pc = m.pc;
        const uint32_t ins = fetch3(m, pc);
        const uint32_t op = ins & 0xFF, imm8 = (ins >> 8) & 0xFF, imm16 = (ins >> 8) & 0xFFFF;
        const bool cb = op == 0xCB;
        const uint32_t dop = cb ? imm8 : op;  // the byte whose x/y/z fields select the operation
        const uint32_t x = dop >> 6, y = (dop >> 3) & 7, z = dop & 7, p = y >> 1, q = y & 1;
        const uint32_t hl = reg_hl(m);
        // ---- read phase: which bytes does this instruction load?
        uint32_t rn = 0, ra = hl;
        if (cb) {
            rn = z == 6;
        } else if (x == 0) {
            if ((z == 4 || z == 5) && y == 6) rn = 1;
            else if (z == 2 && q == 1) { rn = 1; ra = p == 0 ? (m.bcde & 0xFFFF) : p == 1 ? (m.bcde >> 16) : hl; }
        } else if (x == 1 || x == 2) {
            rn = (z == 6 && op != 0x76);
        } else {
            if (z == 0) {
                if (y < 4) { if (condition(m, y)) { rn = 2; ra = m.sp; } }
                else if (y == 6) { rn = 1; ra = 0xFF00 + imm8; }
            } else if (z == 1) {
                if (q == 0 || p < 2) { rn = 2; ra = m.sp; }
            } else if (z == 2) {
                if (y == 6) { rn = 1; ra = 0xFF00 + (m.bcde & 0xFF); }
                else if (y == 7) { rn = 1; ra = imm16; }
            }
        }
        uint32_t rv = 0;
        for (uint32_t i = 0; i < rn; i++) rv |= bus_read_full(m, (ra + i) & 0xFFFF) << (8 * i);
        // ---- execute phase (registers only)
#define NEXT(n) m.pc = (pc + (n)) & 0xFFFF
        if (cb) {
            NEXT(2);
            uint32_t v = (z == 6) ? rv : reg8(m, z), f = reg_f(m), res;
            cycles = z == 6 ? 16 : 8;  // PyBoy's table charges 16 for BIT b,(HL) as well
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
                if (z == 6) WRITE8(hl, res);
                else set_reg8(m, z, res);
            }
        } else if (x == 1) {
            if (op == 0x76) {  // HALT: PC stays on the HALT byte, wake-up adds 1
                m.halted = 1;
                cycles = 4;
            } else {
                uint32_t v = (z == 6) ? rv : reg8(m, z);
                if (y == 6) WRITE8(hl, v);
                else set_reg8(m, y, v);
                NEXT(1);
                cycles = (y == 6 || z == 6) ? 8 : 4;
            }
        } else if (x == 2) {
            alu8(m, y, (z == 6) ? rv : reg8(m, z));
            NEXT(1);
            cycles = z == 6 ? 8 : 4;
        } else if (x == 0) {
            switch (z) {
            case 0:
                if (y == 0) { NEXT(1); cycles = 4; }
                else if (y == 1) {  // LD (nn),SP: low byte first
                    w0a = imm16; w0v = m.sp & 0xFF; w1a = (imm16 + 1) & 0xFFFF; w1v = m.sp >> 8; wn = 2;
                    NEXT(3);
                    cycles = 20;
                } else if (y == 2) { NEXT(2); cycles = 4; }  // STOP
                else if (y == 3 || condition(m, y - 4)) {  // JR
                    m.pc = (pc + 2 + ((imm8 ^ 0x80) - 0x80)) & 0xFFFF;
                    cycles = 12;
                } else { NEXT(2); cycles = 8; }
                break;
            case 1:
                if (q == 0) {
                    set_reg_pair(m, p, imm16);
                    NEXT(3);
                    cycles = 12;
                } else {  // ADD HL,rp
                    uint32_t v = reg_pair(m, p), t = hl + v;
                    set_f(m, (reg_f(m) & FLAG_Z) | (((hl & 0xFFF) + (v & 0xFFF)) > 0xFFF ? FLAG_H : 0) | (t > 0xFFFF ? FLAG_C : 0));
                    set_hl(m, t);
                    NEXT(1);
                    cycles = 8;
                }
                break;
            case 2: {
                uint32_t a = (p == 0) ? (m.bcde & 0xFFFF) : (p == 1) ? (m.bcde >> 16) : hl;
                if (q == 0) WRITE8(a, reg_a(m));
                else set_a(m, rv);
                if (p == 2) set_hl(m, a + 1);
                else if (p == 3) set_hl(m, a - 1);
                NEXT(1);
                cycles = 8;
                break;
            }
            case 3:
                set_reg_pair(m, p, reg_pair(m, p) + (q ? 0xFFFFu : 1u));
                NEXT(1);
                cycles = 8;
                break;
            case 4:
            case 5: {  // INC r / DEC r
                uint32_t v = (y == 6) ? rv : reg8(m, y), res, nf = reg_f(m) & FLAG_C;
                if (z == 4) {
                    res = (v + 1) & 0xFF;
                    nf |= ((v & 0xF) == 0xF ? FLAG_H : 0);
                } else {
                    res = (v - 1) & 0xFF;
                    nf |= FLAG_N | ((v & 0xF) == 0 ? FLAG_H : 0);
                }
                if (res == 0) nf |= FLAG_Z;
                set_f(m, nf);
                if (y == 6) WRITE8(hl, res);
                else set_reg8(m, y, res);
                NEXT(1);
                cycles = y == 6 ? 12 : 4;
                break;
            }
            case 6:
                if (y == 6) WRITE8(hl, imm8);
                else set_reg8(m, y, imm8);
                NEXT(2);
                cycles = y == 6 ? 12 : 8;
                break;
            default: {
                uint32_t a = reg_a(m), f = reg_f(m), c = (f >> 4) & 1;
                switch (y) {
                case 0: set_af(m, (a << 1) | (a >> 7), (a >> 7) ? FLAG_C : 0); break;  // RLCA
                case 1: set_af(m, (a >> 1) | (a << 7), (a & 1) ? FLAG_C : 0); break;   // RRCA
                case 2: set_af(m, (a << 1) | c, (a >> 7) ? FLAG_C : 0); break;         // RLA
                case 3: set_af(m, (a >> 1) | (c << 7), (a & 1) ? FLAG_C : 0); break;   // RRA
                case 4: {                                                              // DAA
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
                case 5: set_af(m, ~a, f | FLAG_N | FLAG_H); break;                     // CPL
                case 6: set_f(m, (f & FLAG_Z) | FLAG_C); break;                        // SCF
                default: set_f(m, (f & FLAG_Z) | ((f & FLAG_C) ^ FLAG_C)); break;      // CCF
                }
                NEXT(1);
                cycles = 4;
                break;
            }
            }
        } else {  // x == 3
            bool illegal = false;
            switch (z) {
            case 0:
                if (y < 4) {  // RET cc
                    if (rn) { m.pc = rv; m.sp = (m.sp + 2) & 0xFFFF; cycles = 20; }
                    else { NEXT(1); cycles = 8; }
                } else if (y == 4) { WRITE8(0xFF00 + imm8, reg_a(m)); NEXT(2); cycles = 12; }
                else if (y == 6) { set_a(m, rv); NEXT(2); cycles = 12; }
                else {  // ADD SP,e / LD HL,SP+e
                    uint32_t sp = m.sp;
                    uint32_t nf = (((sp & 0xF) + (imm8 & 0xF)) > 0xF ? FLAG_H : 0) | (((sp & 0xFF) + imm8) > 0xFF ? FLAG_C : 0);
                    uint32_t t = (sp + ((imm8 ^ 0x80) - 0x80)) & 0xFFFF;
                    set_f(m, nf);
                    NEXT(2);
                    if (y == 5) { m.sp = t; cycles = 16; }
                    else { set_hl(m, t); cycles = 12; }
                }
                break;
            case 1:
                if (q == 0) {  // POP
                    m.sp = (m.sp + 2) & 0xFFFF;
                    if (p == 3) set_af(m, rv >> 8, rv & 0xF0);
                    else set_reg_pair(m, p, rv);
                    NEXT(1);
                    cycles = 12;
                } else if (p < 2) {  // RET / RETI
                    if (p == 1) m.ime = 1;
                    m.pc = rv;
                    m.sp = (m.sp + 2) & 0xFFFF;
                    cycles = 16;
                } else if (p == 2) { m.pc = hl; cycles = 4; }
                else { m.sp = hl; NEXT(1); cycles = 8; }
                break;
            case 2:
                if (y < 4) {
                    if (condition(m, y)) { m.pc = imm16; cycles = 16; }
                    else { NEXT(3); cycles = 12; }
                } else if (y == 4) { WRITE8(0xFF00 + (m.bcde & 0xFF), reg_a(m)); NEXT(1); cycles = 8; }
                else if (y == 5) { WRITE8(imm16, reg_a(m)); NEXT(3); cycles = 16; }
                else if (y == 6) { set_a(m, rv); NEXT(1); cycles = 8; }
                else { set_a(m, rv); NEXT(3); cycles = 16; }
                break;
            case 3:
                if (y == 0) { m.pc = imm16; cycles = 16; }
                else if (y == 6) { m.ime = 0; NEXT(1); cycles = 4; }
                else if (y == 7) { m.ime = 1; NEXT(1); cycles = 4; }  // PyBoy: EI takes effect immediately
                else illegal = true;  // (y == 1 is the CB prefix, handled above)
                break;
            case 4:
                if (y < 4) {
                    NEXT(3);
                    if (condition(m, y)) { PUSH16(m.pc); m.pc = imm16; cycles = 24; }
                    else cycles = 12;
                } else illegal = true;
                break;
            case 5:
                if (q == 0) {  // PUSH
                    PUSH16((p == 3) ? ((reg_a(m) << 8) | reg_f(m)) : reg_pair(m, p));
                    NEXT(1);
                    cycles = 16;
                } else if (p == 0) {  // CALL nn
                    NEXT(3);
                    PUSH16(m.pc);
                    m.pc = imm16;
                    cycles = 24;
                } else illegal = true;
                break;
            case 6:
                alu8(m, y, imm8);
                NEXT(2);
                cycles = 8;
                break;
            default:  // RST
                NEXT(1);
                PUSH16(m.pc);
                m.pc = y * 8;
                cycles = 16;
                break;
            }
            if (illegal) {  // PyBoy raises; we latch a fault and behave as a 1-byte 4-cycle NOP
                m.fault = 1;
                NEXT(1);
                cycles = 4;
            }
        }
#undef NEXT
        m.n_instr++;
        m.iq = 0;
    }
    // ---- write phase
    for (uint32_t i = 0; i < wn; i++) bus_write_full(m, i ? w1a : w0a, i ? w1v : w0v);
#undef PUSH16
#undef WRITE8
    return cycles;
}
