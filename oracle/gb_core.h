/*
 * oracle/gb_core.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99, one env at a time) of the emulator the reference
 * drives through `pokegym/pyboy_binding.py:42-91`: PyBoy 1.6.x (`pyboy<2.0.0`,
 * /root/reference/setup.py:12).  PyBoy itself is NOT vendored under
 * /root/reference and is not installable in this image, so this file restates
 * its published algorithm and is pinned against what the
 * reference tree does hold: the 264 v9 `.state` fixtures, each of which embeds
 * PyBoy's own rendered framebuffer (PPU known-answer vectors) and the exact
 * byte layout of every emulator field.
 *
 * PARITY STATUS: renderer + state codec pinned by the fixtures; the dynamic
 * core (SM83 timing, timer, LCD state machine, MBC3) is "parity unpinned"
 * versus real PyBoy (no PyBoy, no ROM in any environment we control) -- every
 * PyBoy-specific deviation from hardware sits behind a named GBQ_* switch.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may link or call this.  The product (pokegym_b200/csrc) shares no code
 * with it.
 */
#ifndef GB_ORACLE_CORE_H
#define GB_ORACLE_CORE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- PyBoy 1.6.x quirk switches (1 = behave like PyBoy as recalled) ---- */
#define GBQ_IRQ_DISPATCH_ZERO_CYCLES 1 /* cpu.tick returns 0 for an interrupt dispatch            */
#define GBQ_HALT_NO_PC_ADVANCE 1       /* HALT leaves PC on the HALT byte; wake-up does PC += 1    */
#define GBQ_EI_IMMEDIATE 1             /* EI enables IME at once (no one-instruction delay)        */
#define GBQ_BIT_HL_16_CYCLES 1         /* BIT b,(HL) costs 16 (pastraiser table), hardware is 12   */
#define GBQ_TIMA_ONE_INC_PER_TICK 1    /* timer.tick raises TIMA at most once per call             */
#define GBQ_LCD_ONE_TRANSITION 1       /* lcd.tick performs at most one mode change per call       */
#define GBQ_STAT_LOAD_KEEPS_MODE 1     /* load_state goes through STAT.set(): bits 0-2/_mode kept  */
#define GBQ_JOYP_CLEARS_HIGH_NIBBLE 1  /* Interaction.pull: byte &= 4-bit nibble                   */
#define GBQ_MBC3_DISABLE_ONLY_ON_ZERO 1 /* RAM disable only when the written value is exactly 0    */
#define GBQ_LY_WRITABLE 1              /* write to 0xFF44 stores into LY                           */
#define GBQ_COL0_FLAG_FOLLOWS_SHADE 1 /* framebuffer flag bit = "shade is white", for BG and OBJ pixels */
#define GBQ_SOUND_DISABLED 1           /* 0xFF10-3F: reads 0, writes dropped (sound=False)         */

#define GB_STATE_V9_LEN 142610u
#define GB_STATE_V7_LEN 142586u
#define GB_FRAME_CYCLES 70224

enum { GB_BTN_RIGHT = 0, GB_BTN_LEFT, GB_BTN_UP, GB_BTN_DOWN, GB_BTN_A, GB_BTN_B, GB_BTN_SELECT, GB_BTN_START };

typedef struct GbCore {
    /* cpu (pyboy/core/cpu.py) */
    uint8_t A, F, B, C, D, E;
    uint16_t HL, SP, PC;
    uint8_t ime, halted, stopped, IE, IF, interrupt_queued;
    /* lcd (pyboy/core/lcd.py) */
    uint8_t vram[0x2000];
    uint8_t oam[0xA0];
    uint8_t LCDC, BGP, OBP0, OBP1, STAT, LY, LYC, SCY, SCX, WY, WX;
    uint8_t stat_mode; /* STATRegister._mode: not serialised by PyBoy */
    uint64_t clock, clock_target;
    uint8_t next_stat_mode;
    uint8_t frame_done;
    uint8_t disable_renderer;
    int32_t ly_window;               /* Renderer.ly_window: not serialised */
    uint8_t scanline_params[144][5]; /* SCX, SCY, WX-7 (mod 256), WY, tiledata_select */
    uint32_t screen[144 * 160];      /* 0xRRGGBBff, ff bit0 = "BG colour 0" flag */
    /* ram (pyboy/core/ram.py) */
    uint8_t wram[0x2000];
    uint8_t nonio0[96]; /* FEA0-FEFF */
    uint8_t io[76];     /* FF00-FF4B */
    uint8_t hram[127];  /* FF80-FFFE */
    uint8_t nonio1[52]; /* FF4C-FF7F */
    /* timer */
    uint8_t DIV, TIMA, TMA, TAC;
    uint32_t DIV_counter, TIMA_counter;
    /* cartridge (MBC3, 4 x 8 KiB RAM) */
    uint8_t rombank, rambank, ram_enabled, memorymodel;
    uint8_t cart_ram[0x8000];
    /* joypad */
    uint8_t directional, standard;
    /* header bytes of the save-state */
    uint8_t bootrom_enabled, key1, double_speed, cgb;
    /* shared ROM */
    const uint8_t *rom;
    uint32_t rom_banks; /* number of 16 KiB banks */
    /* bookkeeping (ours) */
    uint8_t fault;       /* sticky: 1 = illegal opcode executed (PyBoy would raise) */
    uint64_t n_instr;    /* executed instructions (statistics only) */
    uint64_t n_cycles;   /* emulated T-cycles (statistics only) */
} GbCore;

/* Fresh machine in the post-boot-ROM DMG state (our convention for ROM-only runs). */
void gb_power_on(GbCore *g, const uint8_t *rom, size_t rom_len);
/* PyBoy `load_state` / `save_state` (v9 written; v9 and v7 read). Returns 0 on success. */
int gb_load_state(GbCore *g, const uint8_t *blob, size_t len);
void gb_save_state(const GbCore *g, uint8_t *blob /* GB_STATE_V9_LEN */);
/* PyBoy.get_memory_value / set_memory_value == Motherboard.getitem/setitem. */
uint8_t gb_read(GbCore *g, uint16_t addr);
void gb_write(GbCore *g, uint16_t addr, uint8_t v);
/* Interaction.key_event via Motherboard.buttonevent: pressed=1/0. */
void gb_button(GbCore *g, int button, int pressed);
/* Motherboard.tick(): run until lcd.frame_done. */
void gb_tick(GbCore *g);
/* One iteration of the Motherboard.tick loop body (returns cycles consumed). */
int gb_step_once(GbCore *g);
/* Renderer.scanline + scanline_sprites for line y into g->screen (honours disable_renderer). */
void gb_render_scanline(GbCore *g, int y);
/* pyboy_binding.run_action_on_emulator (:71-91) for action index 0..7, frame_skip frames. */
void gb_run_action(GbCore *g, int action, int frame_skip);
/* screen_ndarray()[::2, ::2] -> 72*80*3 bytes */
void gb_screen_obs_rgb(const GbCore *g, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
