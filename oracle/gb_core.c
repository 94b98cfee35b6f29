/*
 * oracle/gb_core.c -- TEST INFRASTRUCTURE (see gb_core.h header comment).
 *
 * Straight-line restatement of PyBoy 1.6.x (un-vendored dependency of the
 * reference, /root/reference/setup.py:12) as recalled in SURVEY.md Appendix A:
 *   Motherboard.tick / getitem / setitem / transfer_DMA   (pyboy/core/mb.py)
 *   CPU.tick / check_interrupts / handle_interrupt         (pyboy/core/cpu.py)
 *   opcodes (pastraiser cycle table)                       (pyboy/core/opcodes.py)
 *   Timer.tick / cycles_to_interrupt                       (pyboy/core/timer.py)
 *   LCD.tick / set_lcdc / STATRegister / Renderer.scanline (pyboy/core/lcd.py)
 *   MBC3.setitem / BaseMBC.getitem                         (pyboy/core/cartridge/)
 *   Interaction.key_event / pull                           (pyboy/core/interaction.py)
 * Call sites in the reference that reach this code:
 *   /root/reference/pokegym/pyboy_binding.py:44-91 (PyBoy(), send_input, _rendering, tick,
 *   load_state, screen_ndarray) and every get/set_memory_value in ram_map.py / environment.py.
 * One env at a time, byte-per-field state, a read/execute function per instruction group called
 * straight through `gb_read`/`gb_write`.  The CUDA product is organised differently on purpose (packed
 * register words, interleaved HBM arrays, one read site / one write site per instruction, LCD-event
 * main loop, tile-wise 2-bpp renderer), so the two implementations cross-check each other.
 */
#include "gb_core.h"

#include <string.h>

#define FZ 0x80
#define FN 0x40
#define FH 0x20
#define FC 0x10

#define INTR_VBLANK 0x01
#define INTR_LCDC 0x02
#define INTR_TIMER 0x04
#define INTR_SERIAL 0x08
#define INTR_HIGHTOLOW 0x10

static const uint32_t DMG_SHADES[4] = {0xFFFFFFu, 0x999999u, 0x555555u, 0x000000u};

/* ------------------------------------------------------------------ joypad */

static uint8_t joypad_pull(const GbCore *g, uint8_t v) {
    /* Interaction.pull */
    int p14 = (v >> 4) & 1, p15 = (v >> 5) & 1;
    uint8_t b = (uint8_t)(v | 0xCF);
    if (p14 && p15) {
    } else if (!p14 && !p15) {
    } else if (!p14) {
#if GBQ_JOYP_CLEARS_HIGH_NIBBLE
        b &= g->directional;
#else
        b &= (uint8_t)(0xF0 | g->directional);
#endif
    } else {
#if GBQ_JOYP_CLEARS_HIGH_NIBBLE
        b &= g->standard;
#else
        b &= (uint8_t)(0xF0 | g->standard);
#endif
    }
    return b;
}

void gb_button(GbCore *g, int button, int pressed) {
    /* Interaction.key_event + Motherboard.buttonevent */
    uint8_t d0 = g->directional, s0 = g->standard;
    uint8_t *reg = (button < 4) ? &g->directional : &g->standard;
    uint8_t bit = (uint8_t)(1u << (button & 3));
    if (pressed)
        *reg = (uint8_t)(*reg & ~bit);
    else
        *reg = (uint8_t)(*reg | bit);
    if (((d0 ^ g->directional) & d0) || ((s0 ^ g->standard) & s0)) g->IF |= INTR_HIGHTOLOW;
}

/* --------------------------------------------------------------------- lcd */

static uint8_t stat_set_mode(GbCore *g, uint8_t mode) {
    /* STATRegister.set_mode */
    if (g->stat_mode == mode) return 0;
    g->stat_mode = mode;
    g->STAT = (uint8_t)((g->STAT & 0xFC) | mode);
    if (mode != 3 && (g->STAT & (1u << (mode + 3)))) return INTR_LCDC;
    return 0;
}

static uint8_t stat_update_lyc(GbCore *g) {
    /* STATRegister.update_LYC */
    if (g->LYC == g->LY) {
        g->STAT |= 0x04;
        if (g->STAT & 0x40) return INTR_LCDC;
    } else {
        g->STAT &= 0xFB;
    }
    return 0;
}

static void lcd_set_lcdc(GbCore *g, uint8_t v) {
    /* LCD.set_lcdc */
    g->LCDC = v;
    if (!(v & 0x80)) {
        g->clock = 0;
        g->clock_target = GB_FRAME_CYCLES;
        stat_set_mode(g, 0);
        g->next_stat_mode = 2;
        g->LY = 0;
    }
}

static void lcd_set_stat(GbCore *g, uint8_t v) {
    /* STATRegister.set: bit 7 always set, bits 0-2 read-only */
    g->STAT = (uint8_t)((g->STAT & 0x87) | (v & 0x78));
}

static uint32_t pal_pixel(uint8_t pal, int idx) {
    /* PaletteRegister.getcolor: palette_mem_rgb[lookup[idx]]; entry 0 (white) carries COL0_FLAG in
     * every DMG palette (BGP, OBP0, OBP1) -- [STATE-EVIDENCE] white sprite pixels are 0xFFFFFF01. */
    int shade = (pal >> (idx * 2)) & 3;
    uint32_t px = DMG_SHADES[shade] << 8;
#if GBQ_COL0_FLAG_FOLLOWS_SHADE
    if (shade == 0) px |= 1;
#else
    if (idx == 0) px |= 1;
#endif
    return px;
}

static int tile_color(const GbCore *g, int tile, int yy, int xx) {
    /* utils.color_code(byte1, byte2, 7 - x) over VRAM tile data; tile in 0..383 */
    uint8_t b1 = g->vram[tile * 16 + yy * 2];
    uint8_t b2 = g->vram[tile * 16 + yy * 2 + 1];
    int s = 7 - xx;
    return (((b2 >> s) & 1) << 1) | ((b1 >> s) & 1);
}

void gb_render_scanline(GbCore *g, int y) {
    /* Renderer.scanline */
    int bx = g->SCX, by = g->SCY;
    int wx = (int)g->WX - 7, wy = g->WY;
    g->scanline_params[y][0] = (uint8_t)bx;
    g->scanline_params[y][1] = (uint8_t)by;
    g->scanline_params[y][2] = (uint8_t)wx;
    g->scanline_params[y][3] = (uint8_t)wy;
    g->scanline_params[y][4] = (g->LCDC >> 4) & 1;
    if (g->disable_renderer) return;

    int bg_off = (g->LCDC & 0x08) ? 0x1C00 : 0x1800;
    int win_off = (g->LCDC & 0x40) ? 0x1C00 : 0x1800;
    int win_en = (g->LCDC & 0x20) != 0;
    int unsigned_tiles = (g->LCDC & 0x10) != 0;
    int offset = bx & 7;
    uint32_t *row = &g->screen[y * 160];

    if (win_en && wy <= y && wx < 160) g->ly_window += 1;

    for (int x = 0; x < 160; x++) {
        if (win_en && wy <= y && wx <= x) {
            int lw = g->ly_window;
            int addr = win_off + ((lw / 8) * 32) % 0x400 + ((x - wx) / 8) % 32;
            int t = g->vram[addr];
            if (!unsigned_tiles) t = (t ^ 0x80) + 128;
            row[x] = pal_pixel(g->BGP, tile_color(g, t, lw % 8, (x - wx) % 8));
        } else if (g->LCDC & 0x01) {
            int addr = bg_off + (((y + by) / 8) * 32) % 0x400 + ((x + bx) / 8) % 32;
            int t = g->vram[addr];
            if (!unsigned_tiles) t = (t ^ 0x80) + 128;
            row[x] = pal_pixel(g->BGP, tile_color(g, t, (y + by) % 8, (x + offset) % 8));
        } else {
            row[x] = pal_pixel(g->BGP, 0);
        }
    }
    if (y == 143) g->ly_window = -1;

    /* Renderer.scanline_sprites */
    if (!(g->LCDC & 0x02)) return;
    int height = (g->LCDC & 0x04) ? 16 : 8;
    int32_t sel[10];
    int count = 0;
    for (int n = 0; n < 0xA0; n += 4) {
        int sy = (int)g->oam[n] - 16;
        int sx = (int)g->oam[n + 1] - 8;
        if (sy <= y && y < sy + height) {
            sel[count++] = sx * 65536 + n; /* x << 16 | n for the sort (n < 256) */
        }
        if (count == 10) break;
    }
    /* sort ascending by (x, n); insertion sort is stable and exact for <= 10 */
    for (int i = 1; i < count; i++) {
        int32_t k = sel[i];
        int j = i - 1;
        while (j >= 0 && sel[j] > k) {
            sel[j + 1] = sel[j];
            j--;
        }
        sel[j + 1] = k;
    }
    for (int i = count - 1; i >= 0; i--) {
        int n = sel[i] & 0xFF;
        int sy = (int)g->oam[n] - 16;
        int sx = (int)g->oam[n + 1] - 8;
        int tile = g->oam[n + 2];
        if (height == 16) tile &= 0xFE;
        uint8_t attr = g->oam[n + 3];
        int xflip = attr & 0x20, yflip = attr & 0x40, behind = attr & 0x80;
        uint8_t pal = (attr & 0x10) ? g->OBP1 : g->OBP0;
        int dy = y - sy;
        int yy = yflip ? height - dy - 1 : dy;
        int x = sx;
        for (int dx = 0; dx < 8; dx++, x++) {
            int xx = xflip ? 7 - dx : dx;
            int code = tile_color(g, tile, yy, xx); /* tile + 1 reached through yy >= 8 */
            if (x >= 0 && x < 160 && code != 0) {
                if (!behind || (row[x] & 1)) row[x] = pal_pixel(pal, code);
            }
        }
    }
}

static void lcd_blank_screen(GbCore *g) {
    uint32_t px = pal_pixel(g->BGP, 0);
    for (int i = 0; i < 144 * 160; i++) g->screen[i] = px;
}

static uint8_t lcd_tick(GbCore *g, int cycles) {
    /* LCD.tick */
    uint8_t intr = 0;
    g->clock += (uint64_t)cycles;
    if (g->LCDC & 0x80) {
        if (g->clock >= g->clock_target) {
            intr |= stat_set_mode(g, g->next_stat_mode);
            switch (g->stat_mode) {
            case 2:
                if (g->LY == 153) {
                    g->LY = 0;
                    g->clock %= GB_FRAME_CYCLES;
                    g->clock_target %= GB_FRAME_CYCLES;
                } else {
                    g->LY += 1;
                }
                g->clock_target += 80;
                g->next_stat_mode = 3;
                intr |= stat_update_lyc(g);
                break;
            case 3:
                g->clock_target += 170;
                g->next_stat_mode = 0;
                break;
            case 0:
                g->clock_target += 206;
                if (g->LY < 144) gb_render_scanline(g, g->LY);
                g->next_stat_mode = (g->LY < 143) ? 2 : 1;
                break;
            default: /* 1 */
                g->clock_target += 456;
                g->next_stat_mode = 1;
                g->LY += 1;
                intr |= stat_update_lyc(g);
                if (g->LY == 144) {
                    intr |= INTR_VBLANK;
                    g->frame_done = 1;
                }
                if (g->LY == 153) g->next_stat_mode = 2;
                break;
            }
        }
    } else {
        if (g->clock >= GB_FRAME_CYCLES) {
            g->frame_done = 1;
            g->clock %= GB_FRAME_CYCLES;
            lcd_blank_screen(g);
        }
    }
    return intr;
}

/* ------------------------------------------------------------------- timer */

static const int TIMER_DIVIDERS[4] = {1024, 16, 64, 256};

static int timer_tick(GbCore *g, int cycles) {
    g->DIV_counter += (uint32_t)cycles;
    g->DIV = (uint8_t)(g->DIV + (g->DIV_counter >> 8));
    g->DIV_counter &= 0xFF;
    if (!(g->TAC & 4)) return 0;
    g->TIMA_counter += (uint32_t)cycles;
    uint32_t div = (uint32_t)TIMER_DIVIDERS[g->TAC & 3];
    if (g->TIMA_counter >= div) {
        g->TIMA_counter -= div;
        if (g->TIMA == 0xFF) {
            g->TIMA = g->TMA;
            return 1;
        }
        g->TIMA += 1;
    }
    return 0;
}

static int64_t timer_cycles_to_interrupt(const GbCore *g) {
    if (!(g->TAC & 4)) return 1 << 16;
    return (int64_t)(0x100 - g->TIMA) * TIMER_DIVIDERS[g->TAC & 3] - (int64_t)g->TIMA_counter;
}

/* --------------------------------------------------------------------- bus */

uint8_t gb_read(GbCore *g, uint16_t a) {
    if (a < 0x4000) return g->rom[a];
    if (a < 0x8000) return g->rom[(uint32_t)(g->rombank % g->rom_banks) * 0x4000u + (a - 0x4000u)];
    if (a < 0xA000) return g->vram[a - 0x8000];
    if (a < 0xC000) {
        if (!g->ram_enabled) return 0xFF;
        return g->cart_ram[(uint32_t)(g->rambank & 3) * 0x2000u + (a - 0xA000u)];
    }
    if (a < 0xE000) return g->wram[a - 0xC000];
    if (a < 0xFE00) return g->wram[a - 0xE000];
    if (a < 0xFEA0) return g->oam[a - 0xFE00];
    if (a < 0xFF00) return g->nonio0[a - 0xFEA0];
    if (a < 0xFF4C) {
        switch (a) {
        case 0xFF04: return g->DIV;
        case 0xFF05: return g->TIMA;
        case 0xFF06: return g->TMA;
        case 0xFF07: return g->TAC;
        case 0xFF0F: return g->IF;
        case 0xFF40: return g->LCDC;
        case 0xFF41: return g->STAT;
        case 0xFF42: return g->SCY;
        case 0xFF43: return g->SCX;
        case 0xFF44: return g->LY;
        case 0xFF45: return g->LYC;
        case 0xFF46: return 0x00;
        case 0xFF47: return g->BGP;
        case 0xFF48: return g->OBP0;
        case 0xFF49: return g->OBP1;
        case 0xFF4A: return g->WY;
        case 0xFF4B: return g->WX;
        default:
#if GBQ_SOUND_DISABLED
            if (a >= 0xFF10 && a < 0xFF40) return 0;
#endif
            return g->io[a - 0xFF00];
        }
    }
    if (a < 0xFF80) return g->nonio1[a - 0xFF4C];
    if (a < 0xFFFF) return g->hram[a - 0xFF80];
    return g->IE;
}

static void mbc3_write(GbCore *g, uint16_t a, uint8_t v) {
    if (a < 0x2000) {
        if ((v & 0x0F) == 0x0A) {
            g->ram_enabled = 1;
        }
#if GBQ_MBC3_DISABLE_ONLY_ON_ZERO
        else if (v == 0) {
            g->ram_enabled = 0;
        }
#else
        else {
            g->ram_enabled = 0;
        }
#endif
    } else if (a < 0x4000) {
        v &= 0x7F;
        if (v == 0) v = 1;
        g->rombank = v;
    } else if (a < 0x6000) {
        g->rambank = v;
    } else {
        /* RTC latch: cartridge has no RTC, ignored */
    }
}

void gb_write(GbCore *g, uint16_t a, uint8_t v) {
    if (a < 0x8000) {
        mbc3_write(g, a, v);
    } else if (a < 0xA000) {
        g->vram[a - 0x8000] = v;
    } else if (a < 0xC000) {
        if (g->ram_enabled && g->rambank <= 3) g->cart_ram[(uint32_t)g->rambank * 0x2000u + (a - 0xA000u)] = v;
    } else if (a < 0xE000) {
        g->wram[a - 0xC000] = v;
    } else if (a < 0xFE00) {
        g->wram[a - 0xE000] = v;
    } else if (a < 0xFEA0) {
        g->oam[a - 0xFE00] = v;
    } else if (a < 0xFF00) {
        g->nonio0[a - 0xFEA0] = v;
    } else if (a < 0xFF4C) {
        switch (a) {
        case 0xFF00: g->io[0] = joypad_pull(g, v); break;
        case 0xFF04: g->DIV = 0; g->DIV_counter = 0; g->TIMA_counter = 0; break; /* Timer.reset */
        case 0xFF05: g->TIMA = v; break;
        case 0xFF06: g->TMA = v; break;
        case 0xFF07: g->TAC = v & 7; break;
        case 0xFF0F: g->IF = v; break;
        case 0xFF40: lcd_set_lcdc(g, v); break;
        case 0xFF41: lcd_set_stat(g, v); break;
        case 0xFF42: g->SCY = v; break;
        case 0xFF43: g->SCX = v; break;
        case 0xFF44:
#if GBQ_LY_WRITABLE
            g->LY = v;
#endif
            break;
        case 0xFF45: g->LYC = v; break;
        case 0xFF46: { /* Motherboard.transfer_DMA: instantaneous */
            uint16_t src = (uint16_t)(v << 8);
            for (int n = 0; n < 0xA0; n++) gb_write(g, (uint16_t)(0xFE00 + n), gb_read(g, (uint16_t)(src + n)));
            break;
        }
        case 0xFF47: g->BGP = v; break;
        case 0xFF48: g->OBP0 = v; break;
        case 0xFF49: g->OBP1 = v; break;
        case 0xFF4A: g->WY = v; break;
        case 0xFF4B: g->WX = v; break;
        default:
#if GBQ_SOUND_DISABLED
            if (a >= 0xFF10 && a < 0xFF40) break;
#endif
            g->io[a - 0xFF00] = v;
            break;
        }
    } else if (a < 0xFF80) {
        g->nonio1[a - 0xFF4C] = v;
    } else if (a < 0xFFFF) {
        g->hram[a - 0xFF80] = v;
    } else {
        g->IE = v;
    }
}

/* --------------------------------------------------------------------- cpu */

static uint8_t get_r8(GbCore *g, int r) {
    switch (r) {
    case 0: return g->B;
    case 1: return g->C;
    case 2: return g->D;
    case 3: return g->E;
    case 4: return (uint8_t)(g->HL >> 8);
    case 5: return (uint8_t)(g->HL & 0xFF);
    case 6: return gb_read(g, g->HL);
    default: return g->A;
    }
}

static void set_r8(GbCore *g, int r, uint8_t v) {
    switch (r) {
    case 0: g->B = v; break;
    case 1: g->C = v; break;
    case 2: g->D = v; break;
    case 3: g->E = v; break;
    case 4: g->HL = (uint16_t)((g->HL & 0x00FF) | (v << 8)); break;
    case 5: g->HL = (uint16_t)((g->HL & 0xFF00) | v); break;
    case 6: gb_write(g, g->HL, v); break;
    default: g->A = v; break;
    }
}

static uint16_t get_rp(const GbCore *g, int p) {
    switch (p) {
    case 0: return (uint16_t)((g->B << 8) | g->C);
    case 1: return (uint16_t)((g->D << 8) | g->E);
    case 2: return g->HL;
    default: return g->SP;
    }
}

static void set_rp(GbCore *g, int p, uint16_t v) {
    switch (p) {
    case 0: g->B = (uint8_t)(v >> 8); g->C = (uint8_t)v; break;
    case 1: g->D = (uint8_t)(v >> 8); g->E = (uint8_t)v; break;
    case 2: g->HL = v; break;
    default: g->SP = v; break;
    }
}

static void push16(GbCore *g, uint16_t v) {
    gb_write(g, (uint16_t)(g->SP - 1), (uint8_t)(v >> 8));
    gb_write(g, (uint16_t)(g->SP - 2), (uint8_t)(v & 0xFF));
    g->SP = (uint16_t)(g->SP - 2);
}

static uint16_t pop16(GbCore *g) {
    uint8_t lo = gb_read(g, g->SP);
    uint8_t hi = gb_read(g, (uint16_t)(g->SP + 1));
    g->SP = (uint16_t)(g->SP + 2);
    return (uint16_t)((hi << 8) | lo);
}

static int cond(const GbCore *g, int cc) {
    switch (cc) {
    case 0: return !(g->F & FZ);
    case 1: return (g->F & FZ) != 0;
    case 2: return !(g->F & FC);
    default: return (g->F & FC) != 0;
    }
}

static void alu(GbCore *g, int op, uint8_t v) {
    int a = g->A, c = (g->F & FC) ? 1 : 0, t;
    uint8_t f = 0;
    switch (op) {
    case 0: /* ADD */
        t = a + v;
        if ((t & 0xFF) == 0) f |= FZ;
        if ((a & 0xF) + (v & 0xF) > 0xF) f |= FH;
        if (t > 0xFF) f |= FC;
        g->A = (uint8_t)t;
        break;
    case 1: /* ADC */
        t = a + v + c;
        if ((t & 0xFF) == 0) f |= FZ;
        if ((a & 0xF) + (v & 0xF) + c > 0xF) f |= FH;
        if (t > 0xFF) f |= FC;
        g->A = (uint8_t)t;
        break;
    case 2: /* SUB */
    case 7: /* CP */
        t = a - v;
        f |= FN;
        if ((t & 0xFF) == 0) f |= FZ;
        if ((a & 0xF) - (v & 0xF) < 0) f |= FH;
        if (t < 0) f |= FC;
        if (op == 2) g->A = (uint8_t)t;
        break;
    case 3: /* SBC */
        t = a - v - c;
        f |= FN;
        if ((t & 0xFF) == 0) f |= FZ;
        if ((a & 0xF) - (v & 0xF) - c < 0) f |= FH;
        if (t < 0) f |= FC;
        g->A = (uint8_t)t;
        break;
    case 4: /* AND */
        g->A = (uint8_t)(a & v);
        f = FH;
        if (g->A == 0) f |= FZ;
        break;
    case 5: /* XOR */
        g->A = (uint8_t)(a ^ v);
        if (g->A == 0) f |= FZ;
        break;
    default: /* 6 OR */
        g->A = (uint8_t)(a | v);
        if (g->A == 0) f |= FZ;
        break;
    }
    g->F = f;
}

static int exec_cb(GbCore *g) {
    uint8_t op = gb_read(g, (uint16_t)(g->PC + 1));
    int r = op & 7, y = (op >> 3) & 7, x = op >> 6;
    int v = get_r8(g, r), t, c = (g->F & FC) ? 1 : 0;
    g->PC = (uint16_t)(g->PC + 2);
    if (x == 0) {
        switch (y) {
        case 0: t = (v << 1) + (v >> 7); break;                           /* RLC */
        case 1: t = (v >> 1) + ((v & 1) << 7) + ((v & 1) << 8); break;    /* RRC */
        case 2: t = (v << 1) + c; break;                                  /* RL  */
        case 3: t = (v >> 1) + (c << 7) + ((v & 1) << 8); break;          /* RR  */
        case 4: t = v << 1; break;                                        /* SLA */
        case 5: t = ((v >> 1) | (v & 0x80)) + ((v & 1) << 8); break;      /* SRA */
        case 6: t = ((v & 0xF0) >> 4) | ((v & 0x0F) << 4); break;         /* SWAP */
        default: t = (v >> 1) + ((v & 1) << 8); break;                    /* SRL */
        }
        uint8_t f = 0;
        if ((t & 0xFF) == 0) f |= FZ;
        if (t > 0xFF) f |= FC;
        g->F = f;
        set_r8(g, r, (uint8_t)t);
        return r == 6 ? 16 : 8;
    }
    if (x == 1) { /* BIT */
        uint8_t f = (uint8_t)((g->F & FC) | FH);
        if (!(v & (1 << y))) f |= FZ;
        g->F = f;
#if GBQ_BIT_HL_16_CYCLES
        return r == 6 ? 16 : 8;
#else
        return r == 6 ? 12 : 8;
#endif
    }
    if (x == 2)
        set_r8(g, r, (uint8_t)(v & ~(1 << y))); /* RES */
    else
        set_r8(g, r, (uint8_t)(v | (1 << y))); /* SET */
    return r == 6 ? 16 : 8;
}

static int fetch_and_execute(GbCore *g) {
    uint16_t pc = g->PC;
    uint8_t op = gb_read(g, pc);
    uint8_t n8 = gb_read(g, (uint16_t)(pc + 1));
#ifdef GB_OPCODE_HOOK /* optional instruction-mix profiling (tools only) */
    GB_OPCODE_HOOK(op, n8);
#endif
    uint16_t n16 = (uint16_t)(n8 | (gb_read(g, (uint16_t)(pc + 2)) << 8));
    int s8 = (int)((n8 ^ 0x80) - 0x80);
    int x = op >> 6, y = (op >> 3) & 7, z = op & 7, p = y >> 1, q = y & 1;
    int t;
    uint8_t f;

    if (x == 1) {
        if (op == 0x76) { /* HALT */
            g->halted = 1;
#if !GBQ_HALT_NO_PC_ADVANCE
            g->PC = (uint16_t)(pc + 1);
#endif
            return 4;
        }
        set_r8(g, y, get_r8(g, z));
        g->PC = (uint16_t)(pc + 1);
        return (y == 6 || z == 6) ? 8 : 4;
    }
    if (x == 2) {
        alu(g, y, get_r8(g, z));
        g->PC = (uint16_t)(pc + 1);
        return z == 6 ? 8 : 4;
    }
    if (x == 0) {
        switch (z) {
        case 0:
            if (y == 0) { g->PC = (uint16_t)(pc + 1); return 4; } /* NOP */
            if (y == 1) { /* LD (nn),SP */
                gb_write(g, n16, (uint8_t)(g->SP & 0xFF));
                gb_write(g, (uint16_t)(n16 + 1), (uint8_t)(g->SP >> 8));
                g->PC = (uint16_t)(pc + 3);
                return 20;
            }
            if (y == 2) { g->PC = (uint16_t)(pc + 2); return 4; } /* STOP (DMG: skip a byte) */
            if (y == 3 || cond(g, y - 4)) { /* JR */
                g->PC = (uint16_t)(pc + 2 + s8);
                return 12;
            }
            g->PC = (uint16_t)(pc + 2);
            return 8;
        case 1:
            if (q == 0) { set_rp(g, p, n16); g->PC = (uint16_t)(pc + 3); return 12; }
            {
                int hl = g->HL, v = get_rp(g, p);
                t = hl + v;
                f = g->F & FZ;
                if ((hl & 0xFFF) + (v & 0xFFF) > 0xFFF) f |= FH;
                if (t > 0xFFFF) f |= FC;
                g->F = f;
                g->HL = (uint16_t)t;
                g->PC = (uint16_t)(pc + 1);
                return 8;
            }
        case 2: {
            uint16_t addr = p == 0 ? get_rp(g, 0) : p == 1 ? get_rp(g, 1) : g->HL;
            if (q == 0) gb_write(g, addr, g->A); else g->A = gb_read(g, addr);
            if (p == 2) g->HL = (uint16_t)(g->HL + 1);
            if (p == 3) g->HL = (uint16_t)(g->HL - 1);
            g->PC = (uint16_t)(pc + 1);
            return 8;
        }
        case 3:
            set_rp(g, p, (uint16_t)(get_rp(g, p) + (q ? -1 : 1)));
            g->PC = (uint16_t)(pc + 1);
            return 8;
        case 4: { /* INC r */
            int v = get_r8(g, y);
            t = v + 1;
            f = g->F & FC;
            if ((t & 0xFF) == 0) f |= FZ;
            if ((v & 0xF) + 1 > 0xF) f |= FH;
            g->F = f;
            set_r8(g, y, (uint8_t)t);
            g->PC = (uint16_t)(pc + 1);
            return y == 6 ? 12 : 4;
        }
        case 5: { /* DEC r */
            int v = get_r8(g, y);
            t = v - 1;
            f = (uint8_t)((g->F & FC) | FN);
            if ((t & 0xFF) == 0) f |= FZ;
            if ((v & 0xF) - 1 < 0) f |= FH;
            g->F = f;
            set_r8(g, y, (uint8_t)t);
            g->PC = (uint16_t)(pc + 1);
            return y == 6 ? 12 : 4;
        }
        case 6:
            set_r8(g, y, n8);
            g->PC = (uint16_t)(pc + 2);
            return y == 6 ? 12 : 8;
        default: {
            int a = g->A, c = (g->F & FC) ? 1 : 0;
            switch (y) {
            case 0: t = (a << 1) + (a >> 7); g->F = t > 0xFF ? FC : 0; g->A = (uint8_t)t; break;               /* RLCA */
            case 1: t = (a >> 1) + ((a & 1) << 7) + ((a & 1) << 8); g->F = t > 0xFF ? FC : 0; g->A = (uint8_t)t; break; /* RRCA */
            case 2: t = (a << 1) + c; g->F = t > 0xFF ? FC : 0; g->A = (uint8_t)t; break;                    /* RLA */
            case 3: t = (a >> 1) + (c << 7) + ((a & 1) << 8); g->F = t > 0xFF ? FC : 0; g->A = (uint8_t)t; break; /* RRA */
            case 4: { /* DAA */
                int corr = 0;
                t = a;
                if (g->F & FH) corr |= 0x06;
                if (g->F & FC) corr |= 0x60;
                if (g->F & FN) {
                    t -= corr;
                } else {
                    if ((t & 0x0F) > 0x09) corr |= 0x06;
                    if (t > 0x99) corr |= 0x60;
                    t += corr;
                }
                f = g->F & FN;
                if ((t & 0xFF) == 0) f |= FZ;
                if (corr & 0x60) f |= FC;
                g->F = f;
                g->A = (uint8_t)t;
                break;
            }
            case 5: g->A = (uint8_t)~a; g->F |= (FN | FH); break;                              /* CPL */
            case 6: g->F = (uint8_t)((g->F & FZ) | FC); break;                                 /* SCF */
            default: g->F = (uint8_t)((g->F & FZ) | ((g->F & FC) ^ FC)); break;                /* CCF */
            }
            g->PC = (uint16_t)(pc + 1);
            return 4;
        }
        }
    }
    /* x == 3 */
    switch (z) {
    case 0:
        if (y < 4) { /* RET cc */
            if (cond(g, y)) { g->PC = pop16(g); return 20; }
            g->PC = (uint16_t)(pc + 1);
            return 8;
        }
        if (y == 4) { gb_write(g, (uint16_t)(0xFF00 + n8), g->A); g->PC = (uint16_t)(pc + 2); return 12; }
        if (y == 6) { g->A = gb_read(g, (uint16_t)(0xFF00 + n8)); g->PC = (uint16_t)(pc + 2); return 12; }
        { /* ADD SP,e / LD HL,SP+e */
            int sp = g->SP;
            f = 0;
            if ((sp & 0xF) + (n8 & 0xF) > 0xF) f |= FH;
            if ((sp & 0xFF) + (n8 & 0xFF) > 0xFF) f |= FC;
            g->F = f;
            g->PC = (uint16_t)(pc + 2);
            if (y == 5) { g->SP = (uint16_t)(sp + s8); return 16; }
            g->HL = (uint16_t)(sp + s8);
            return 12;
        }
    case 1:
        if (q == 0) { /* POP */
            uint16_t v = pop16(g);
            if (p == 3) { g->A = (uint8_t)(v >> 8); g->F = (uint8_t)(v & 0xF0); }
            else set_rp(g, p, v);
            g->PC = (uint16_t)(pc + 1);
            return 12;
        }
        if (p == 0) { g->PC = pop16(g); return 16; }                  /* RET  */
        if (p == 1) { g->ime = 1; g->PC = pop16(g); return 16; }      /* RETI */
        if (p == 2) { g->PC = g->HL; return 4; }                      /* JP HL */
        g->SP = g->HL; g->PC = (uint16_t)(pc + 1); return 8;          /* LD SP,HL */
    case 2:
        if (y < 4) {
            if (cond(g, y)) { g->PC = n16; return 16; }
            g->PC = (uint16_t)(pc + 3);
            return 12;
        }
        if (y == 4) { gb_write(g, (uint16_t)(0xFF00 + g->C), g->A); g->PC = (uint16_t)(pc + 1); return 8; }
        if (y == 5) { gb_write(g, n16, g->A); g->PC = (uint16_t)(pc + 3); return 16; }
        if (y == 6) { g->A = gb_read(g, (uint16_t)(0xFF00 + g->C)); g->PC = (uint16_t)(pc + 1); return 8; }
        g->A = gb_read(g, n16); g->PC = (uint16_t)(pc + 3); return 16;
    case 3:
        if (y == 0) { g->PC = n16; return 16; }
        if (y == 1) return exec_cb(g);
        if (y == 6) { g->ime = 0; g->PC = (uint16_t)(pc + 1); return 4; }
        if (y == 7) { g->ime = 1; g->PC = (uint16_t)(pc + 1); return 4; } /* GBQ_EI_IMMEDIATE */
        break;
    case 4:
        if (y < 4) {
            g->PC = (uint16_t)(pc + 3);
            if (cond(g, y)) { push16(g, g->PC); g->PC = n16; return 24; }
            return 12;
        }
        break;
    case 5:
        if (q == 0) { /* PUSH */
            uint16_t v = p == 3 ? (uint16_t)((g->A << 8) | g->F) : get_rp(g, p);
            push16(g, v);
            g->PC = (uint16_t)(pc + 1);
            return 16;
        }
        if (p == 0) { g->PC = (uint16_t)(pc + 3); push16(g, g->PC); g->PC = n16; return 24; }
        break;
    case 6:
        alu(g, y, n8);
        g->PC = (uint16_t)(pc + 2);
        return 8;
    default: /* RST */
        g->PC = (uint16_t)(pc + 1);
        push16(g, g->PC);
        g->PC = (uint16_t)(y * 8);
        return 16;
    }
    /* illegal opcode: PyBoy raises; we latch a fault and treat it as a 4-cycle NOP */
    g->fault = 1;
    g->PC = (uint16_t)(pc + 1);
    return 4;
}

static int handle_interrupt(GbCore *g, uint8_t flag, uint16_t vec) {
    if ((g->IE & flag) && (g->IF & flag)) {
        if (g->halted) g->PC = (uint16_t)(g->PC + 1);
        if (g->ime) {
            g->IF ^= flag;
            push16(g, g->PC);
            g->PC = vec;
            g->ime = 0;
        }
        return 1;
    }
    return 0;
}

static int check_interrupts(GbCore *g) {
    if (g->interrupt_queued) return 0;
    if ((g->IF & 0x1F) & (g->IE & 0x1F)) {
        if (handle_interrupt(g, INTR_VBLANK, 0x40)) g->interrupt_queued = 1;
        else if (handle_interrupt(g, INTR_LCDC, 0x48)) g->interrupt_queued = 1;
        else if (handle_interrupt(g, INTR_TIMER, 0x50)) g->interrupt_queued = 1;
        else if (handle_interrupt(g, INTR_SERIAL, 0x58)) g->interrupt_queued = 1;
        else if (handle_interrupt(g, INTR_HIGHTOLOW, 0x60)) g->interrupt_queued = 1;
        else g->interrupt_queued = 0;
        return 1;
    }
    g->interrupt_queued = 0;
    return 0;
}

static int cpu_tick(GbCore *g) {
    if (check_interrupts(g)) {
        g->halted = 0;
        return 0; /* GBQ_IRQ_DISPATCH_ZERO_CYCLES */
    }
    if (g->halted && g->interrupt_queued) {
        g->halted = 0;
        g->PC = (uint16_t)(g->PC + 1);
    } else if (g->halted) {
        return 4;
    }
    int cycles = fetch_and_execute(g);
    g->n_instr++;
    g->interrupt_queued = 0;
    return cycles;
}

int gb_step_once(GbCore *g) {
    int64_t cycles = cpu_tick(g);
    if (g->halted) {
        int64_t a = (int64_t)g->clock_target - (int64_t)g->clock;
        int64_t b = timer_cycles_to_interrupt(g);
        cycles = a < b ? a : b;
        if (cycles < 0) cycles = 0;
    }
    g->n_cycles += (uint64_t)cycles;
    if (timer_tick(g, (int)cycles)) g->IF |= INTR_TIMER;
    g->IF |= lcd_tick(g, (int)cycles);
    return (int)cycles;
}

void gb_tick(GbCore *g) {
    for (;;) {
        int processing = !g->frame_done;
        g->frame_done = 0;
        if (!processing) break;
        gb_step_once(g);
    }
}

/* pyboy_binding.ACTIONS (:40): Down, Left, Right, Up, A, B, Start, Select */
static const int ACTION_BUTTON[8] = {GB_BTN_DOWN, GB_BTN_LEFT, GB_BTN_RIGHT, GB_BTN_UP, GB_BTN_A, GB_BTN_B, GB_BTN_START, GB_BTN_SELECT};

void gb_run_action(GbCore *g, int action, int frame_skip) {
    /* pyboy_binding.run_action_on_emulator :71-91 with PyBoy.tick applying queued inputs first */
    int btn = ACTION_BUTTON[action & 7];
    g->disable_renderer = 1;
    for (int i = 0; i < frame_skip; i++) {
        if (i == 0) gb_button(g, btn, 1);
        if (i == 8) gb_button(g, btn, 0);
        if (i == frame_skip - 1) g->disable_renderer = 0;
        gb_tick(g);
    }
}

void gb_screen_obs_rgb(const GbCore *g, uint8_t *out) {
    for (int r = 0; r < 72; r++)
        for (int c = 0; c < 80; c++) {
            uint32_t px = g->screen[(2 * r) * 160 + 2 * c];
            uint8_t *o = out + (r * 80 + c) * 3;
            o[0] = (uint8_t)(px >> 8);
            o[1] = (uint8_t)(px >> 16);
            o[2] = (uint8_t)(px >> 24);
        }
}

/* ------------------------------------------------------------ state codec */

void gb_power_on(GbCore *g, const uint8_t *rom, size_t rom_len) {
    memset(g, 0, sizeof(*g));
    g->rom = rom;
    g->rom_banks = (uint32_t)(rom_len / 0x4000);
    if (g->rom_banks == 0) g->rom_banks = 1;
    g->A = 0x01; g->F = 0xB0; g->B = 0x00; g->C = 0x13; g->D = 0x00; g->E = 0xD8;
    g->HL = 0x014D; g->SP = 0xFFFE; g->PC = 0x0100;
    g->LCDC = 0x91; g->BGP = 0xFC; g->OBP0 = 0xFF; g->OBP1 = 0xFF; g->STAT = 0x80;
    g->next_stat_mode = 2;
    g->ly_window = -1;
    g->rombank = 1;
    g->directional = 0x0F; g->standard = 0x0F;
    g->io[0] = 0xFF;
    g->nonio1[4] = 1; /* FF50: boot ROM unmapped */
}

static uint16_t rd16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static uint64_t rd64(const uint8_t *p) {
    uint64_t v = 0;
    for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
    return v;
}
static void wr16(uint8_t *p, uint16_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void wr64(uint8_t *p, uint64_t v) { for (int i = 0; i < 8; i++) { p[i] = (uint8_t)v; v >>= 8; } }

int gb_load_state(GbCore *g, const uint8_t *b, size_t len) {
    /* Motherboard.load_state and the per-component load_state methods, in file order */
    if (len < 1) return -1;
    int ver = b[0];
    if (ver == 9 && len != GB_STATE_V9_LEN) return -2;
    if (ver == 7 && len != GB_STATE_V7_LEN) return -2;
    if (ver != 9 && ver != 7) return -3;
    const uint8_t *p = b + 1;
    g->bootrom_enabled = *p++;
    if (ver >= 8) {
        g->key1 = *p++;
        g->double_speed = *p++;
        g->cgb = *p++;
        if (g->cgb) return -4;
    }
    g->A = *p++; g->F = *p++; g->B = *p++; g->C = *p++; g->D = *p++; g->E = *p++;
    g->HL = rd16(p); p += 2;
    g->SP = rd16(p); p += 2;
    g->PC = rd16(p); p += 2;
    g->ime = *p++; g->halted = *p++; g->stopped = *p++; g->IE = *p++;
    if (ver >= 8) { g->interrupt_queued = *p++; g->IF = *p++; }
    memcpy(g->vram, p, 0x2000); p += 0x2000;
    memcpy(g->oam, p, 0xA0); p += 0xA0;
    lcd_set_lcdc(g, *p++);
    g->BGP = *p++; g->OBP0 = *p++; g->OBP1 = *p++;
#if GBQ_STAT_LOAD_KEEPS_MODE
    lcd_set_stat(g, *p++);
#else
    g->STAT = *p++; g->stat_mode = g->STAT & 3;
#endif
    g->LY = *p++; g->LYC = *p++;
    g->SCY = *p++; g->SCX = *p++; g->WY = *p++; g->WX = *p++;
    if (ver >= 8) {
        p++; /* cgb */
        p++; /* double_speed */
        g->clock = rd64(p); p += 8;
        g->clock_target = rd64(p); p += 8;
        g->next_stat_mode = *p++;
    }
    for (int y = 0; y < 144; y++) {
        g->scanline_params[y][0] = *p++;
        g->scanline_params[y][1] = *p++;
        g->scanline_params[y][2] = (uint8_t)(*p++ - 7);
        g->scanline_params[y][3] = *p++;
        g->scanline_params[y][4] = *p++;
    }
    for (int i = 0; i < 144 * 160; i++, p += 4) g->screen[i] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    memcpy(g->wram, p, 0x2000); p += 0x2000;
    memcpy(g->nonio0, p, 96); p += 96;
    memcpy(g->io, p, 76); p += 76;
    memcpy(g->hram, p, 127); p += 127;
    memcpy(g->nonio1, p, 52); p += 52;
    g->DIV = *p++; g->TIMA = *p++;
    g->DIV_counter = rd16(p); p += 2;
    g->TIMA_counter = rd16(p); p += 2;
    g->TMA = *p++; g->TAC = *p++;
    g->rombank = *p++; g->rambank = *p++; g->ram_enabled = *p++; g->memorymodel = *p++;
    memcpy(g->cart_ram, p, 0x8000); p += 0x8000;
    g->directional = *p++; g->standard = *p++;
    return ((size_t)(p - b) == len) ? 0 : -5;
}

void gb_save_state(const GbCore *g, uint8_t *b) {
    uint8_t *p = b;
    *p++ = 9; *p++ = g->bootrom_enabled; *p++ = g->key1; *p++ = g->double_speed; *p++ = g->cgb;
    *p++ = g->A; *p++ = g->F; *p++ = g->B; *p++ = g->C; *p++ = g->D; *p++ = g->E;
    wr16(p, g->HL); p += 2; wr16(p, g->SP); p += 2; wr16(p, g->PC); p += 2;
    *p++ = g->ime; *p++ = g->halted; *p++ = g->stopped; *p++ = g->IE; *p++ = g->interrupt_queued; *p++ = g->IF;
    memcpy(p, g->vram, 0x2000); p += 0x2000;
    memcpy(p, g->oam, 0xA0); p += 0xA0;
    *p++ = g->LCDC; *p++ = g->BGP; *p++ = g->OBP0; *p++ = g->OBP1; *p++ = g->STAT; *p++ = g->LY; *p++ = g->LYC;
    *p++ = g->SCY; *p++ = g->SCX; *p++ = g->WY; *p++ = g->WX;
    *p++ = g->cgb; *p++ = g->double_speed;
    wr64(p, g->clock); p += 8; wr64(p, g->clock_target); p += 8;
    *p++ = g->next_stat_mode;
    for (int y = 0; y < 144; y++) {
        *p++ = g->scanline_params[y][0];
        *p++ = g->scanline_params[y][1];
        *p++ = (uint8_t)(g->scanline_params[y][2] + 7);
        *p++ = g->scanline_params[y][3];
        *p++ = g->scanline_params[y][4];
    }
    for (int i = 0; i < 144 * 160; i++) { uint32_t v = g->screen[i]; *p++ = (uint8_t)v; *p++ = (uint8_t)(v >> 8); *p++ = (uint8_t)(v >> 16); *p++ = (uint8_t)(v >> 24); }
    memcpy(p, g->wram, 0x2000); p += 0x2000;
    memcpy(p, g->nonio0, 96); p += 96;
    memcpy(p, g->io, 76); p += 76;
    memcpy(p, g->hram, 127); p += 127;
    memcpy(p, g->nonio1, 52); p += 52;
    *p++ = g->DIV; *p++ = g->TIMA;
    wr16(p, (uint16_t)g->DIV_counter); p += 2; wr16(p, (uint16_t)g->TIMA_counter); p += 2;
    *p++ = g->TMA; *p++ = g->TAC;
    *p++ = g->rombank; *p++ = g->rambank; *p++ = g->ram_enabled; *p++ = g->memorymodel;
    memcpy(p, g->cart_ram, 0x8000); p += 0x8000;
    *p++ = g->directional; *p++ = g->standard;
}
