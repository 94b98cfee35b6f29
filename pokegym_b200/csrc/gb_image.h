// gb_image.h -- PyBoy save-state blob <-> canonical per-env image (host code).
//
// The canonical image is the linear word layout of gb_layout.cuh (IMG_*): what state templates hold on the device and
// what k_scatter_image / k_gather_image move in and out of the interleaved arrays.  Shared by libgbenv.so (gbenv.cu)
// and the host-simulation test harness (tests/hostsim).
// Replaces pyboy_binding.open_state_file / load_pyboy_state (/root/reference/pokegym/pyboy_binding.py:59-69) and
// PyBoy.save_state; the field order is SURVEY.md section 8c.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/gbenv.h"
#include "gb_layout.cuh"

static inline void img_set_byte(std::vector<uint32_t> &img, uint32_t base_word, uint32_t byte_index, uint8_t v) {
    uint32_t &w = img[base_word + (byte_index >> 2)];
    uint32_t sh = (byte_index & 3) * 8;
    w = (w & ~(0xFFu << sh)) | ((uint32_t)v << sh);
}
static inline uint8_t img_get_byte(const std::vector<uint32_t> &img, uint32_t base_word, uint32_t byte_index) {
    return (uint8_t)(img[base_word + (byte_index >> 2)] >> ((byte_index & 3) * 8));
}

static const uint32_t SHADE_WORDS[4] = {0xFFFFFF01u, 0x99999900u, 0x55555500u, 0x00000000u};

// PyBoy v9 / v7 blob -> canonical image (layout: SURVEY.md 8c).  Returns 0 or a negative error.
static int blob_to_image(const uint8_t *b, size_t len, std::vector<uint32_t> &img, int *version_out, std::string &err) {
    if (len < 1) { err = "empty save-state"; return GBENV_E_STATE; }
    int ver = b[0];
    if (!((ver == 9 && len == 142610) || (ver == 7 && len == 142586))) {
        err = "unsupported PyBoy save-state (need v9/142610 B or v7/142586 B)";
        return GBENV_E_STATE;
    }
    img.assign(IMG_WORDS, 0);
    const uint8_t *p = b + 1;
    uint32_t hdr = *p++;  // bootrom_enabled
    if (ver >= 8) {
        hdr |= (uint32_t)p[0] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 24;
        if (p[2]) { err = "CGB save-states are not supported"; return GBENV_E_STATE; }
        p += 3;
    }
    uint32_t A = p[0], F = p[1], B = p[2], C = p[3], D = p[4], E = p[5];
    uint32_t HL = p[6] | (p[7] << 8), SP = p[8] | (p[9] << 8), PC = p[10] | (p[11] << 8);
    uint32_t ime = p[12], halted = p[13], stopped = p[14], IE = p[15], iq = 0, IF = 0;
    p += 16;
    if (ver >= 8) { iq = p[0]; IF = p[1]; p += 2; }
    uint32_t *regs = &img[IMG_REGS];
    regs[R_BCDE] = C | (B << 8) | (E << 16) | (D << 24);
    regs[R_HLAF] = HL | (A << 16) | (F << 24);
    regs[R_SPPC] = SP | (PC << 16);
    regs[R_INT] = (ime & 1) | ((halted & 1) << 1) | ((stopped & 1) << 2) | ((iq & 1) << 3) | (IE << 8) | (IF << 16);
    for (uint32_t i = 0; i < 0x2000; i++) img_set_byte(img, IMG_MEM, MEM_VRAM + i, *p++);
    for (uint32_t i = 0; i < 0xA0; i++) img_set_byte(img, IMG_MEM, MEM_HI + i, *p++);
    uint32_t LCDC = p[0], BGP = p[1], OBP0 = p[2], OBP1 = p[3], STAT = p[4], LY = p[5], LYC = p[6], SCY = p[7], SCX = p[8], WY = p[9], WX = p[10];
    p += 11;
    uint64_t clock = 0, target = 0;
    uint32_t next_mode = 2;
    if (ver >= 8) {
        p += 2;
        for (int i = 7; i >= 0; i--) clock = (clock << 8) | p[i];
        p += 8;
        for (int i = 7; i >= 0; i--) target = (target << 8) | p[i];
        p += 8;
        next_mode = *p++;
        if (clock > 0xFFFFFFFFull || target > 0xFFFFFFFFull) { err = "LCD clock out of range"; return GBENV_E_STATE; }
    }
    regs[R_LCD0] = LCDC | (STAT << 8) | (LY << 16) | (LYC << 24);
    regs[R_LCD1] = SCY | (SCX << 8) | (WY << 16) | (WX << 24);
    regs[R_LCD2] = BGP | (OBP0 << 8) | (OBP1 << 16) | (((STAT & 3) | ((next_mode & 3) << 2)) << 24);
    regs[R_CLOCK] = (uint32_t)clock;
    regs[R_TARGET] = (uint32_t)target;
    for (uint32_t y = 0; y < 144; y++, p += 5) {
        img[IMG_LP + 2 * y] = p[0] | (p[1] << 8) | (p[2] << 16) | (p[3] << 24);  // SCX SCY WX(raw) WY
        img[IMG_LP + 2 * y + 1] = p[4] ? 0x10u : 0u;                              // tile_data_select
    }
    for (uint32_t i = 0; i < 144 * 160; i++, p += 4) {
        uint32_t wv = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
        int shade = -1;
        for (int s = 0; s < 4; s++)
            if (wv == SHADE_WORDS[s]) shade = s;
        if (shade < 0 && wv == 0) shade = 3;
        if (shade < 0) { err = "framebuffer word is not one of PyBoy's four DMG values"; return GBENV_E_STATE; }
        img[IMG_FB + (i >> 4)] |= (uint32_t)shade << (2 * (i & 15));
    }
    for (uint32_t i = 0; i < 0x2000; i++) img_set_byte(img, IMG_MEM, MEM_WRAM + i, *p++);
    for (uint32_t i = 0; i < 96; i++) img_set_byte(img, IMG_MEM, MEM_HI + 0xA0 + i, *p++);
    for (uint32_t i = 0; i < 76; i++) img_set_byte(img, IMG_MEM, MEM_HI + 0x100 + i, *p++);
    for (uint32_t i = 0; i < 127; i++) img_set_byte(img, IMG_MEM, MEM_HI + 0x180 + i, *p++);
    for (uint32_t i = 0; i < 52; i++) img_set_byte(img, IMG_MEM, MEM_HI + 0x14C + i, *p++);
    uint32_t DIV = p[0], TIMA = p[1], DIVC = p[2] | (p[3] << 8), TIMAC = p[4] | (p[5] << 8), TMA = p[6], TAC = p[7];
    p += 8;
    regs[R_TIMER] = DIV | (TIMA << 8) | (TMA << 16) | (TAC << 24);
    regs[R_DIVC] = DIVC;
    regs[R_TIMAC] = TIMAC;
    regs[R_MBC] = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
    p += 4;
    for (uint32_t i = 0; i < 0x8000; i++) img_set_byte(img, IMG_CRAM, i, *p++);
    regs[R_JOY] = p[0] | (p[1] << 8) | (0xFFu << 16) | (144u << 24);  // ly_window kept by the merge, lp_dirty = 144
    p += 2;
    regs[R_HDR] = hdr;
    regs[R_MISC] = 0xFF;
    if ((size_t)(p - b) != len) { err = "save-state length mismatch"; return GBENV_E_STATE; }
    *version_out = ver;
    return GBENV_OK;
}

static void image_to_blob(const std::vector<uint32_t> &img, uint8_t *b) {
    const uint32_t *regs = &img[IMG_REGS];
    uint8_t *p = b;
    uint32_t hdr = regs[R_HDR];
    *p++ = 9; *p++ = hdr & 0xFF; *p++ = (hdr >> 8) & 0xFF; *p++ = (hdr >> 16) & 0xFF; *p++ = hdr >> 24;
    uint32_t bcde = regs[R_BCDE], hlaf = regs[R_HLAF], sppc = regs[R_SPPC], in = regs[R_INT];
    *p++ = (hlaf >> 16) & 0xFF; *p++ = hlaf >> 24; *p++ = (bcde >> 8) & 0xFF; *p++ = bcde & 0xFF; *p++ = bcde >> 24; *p++ = (bcde >> 16) & 0xFF;
    *p++ = hlaf & 0xFF; *p++ = (hlaf >> 8) & 0xFF;
    *p++ = sppc & 0xFF; *p++ = (sppc >> 8) & 0xFF; *p++ = (sppc >> 16) & 0xFF; *p++ = sppc >> 24;
    *p++ = in & 1; *p++ = (in >> 1) & 1; *p++ = (in >> 2) & 1; *p++ = (in >> 8) & 0xFF; *p++ = (in >> 3) & 1; *p++ = (in >> 16) & 0xFF;
    for (uint32_t i = 0; i < 0x2000; i++) *p++ = img_get_byte(img, IMG_MEM, MEM_VRAM + i);
    for (uint32_t i = 0; i < 0xA0; i++) *p++ = img_get_byte(img, IMG_MEM, MEM_HI + i);
    uint32_t l0 = regs[R_LCD0], l1 = regs[R_LCD1], l2 = regs[R_LCD2];
    *p++ = l0 & 0xFF; *p++ = l2 & 0xFF; *p++ = (l2 >> 8) & 0xFF; *p++ = (l2 >> 16) & 0xFF; *p++ = (l0 >> 8) & 0xFF; *p++ = (l0 >> 16) & 0xFF; *p++ = l0 >> 24;
    *p++ = l1 & 0xFF; *p++ = (l1 >> 8) & 0xFF; *p++ = (l1 >> 16) & 0xFF; *p++ = l1 >> 24;
    *p++ = hdr >> 24; *p++ = (hdr >> 16) & 0xFF;
    uint64_t clock = regs[R_CLOCK], target = regs[R_TARGET];
    for (int i = 0; i < 8; i++) { *p++ = (uint8_t)clock; clock >>= 8; }
    for (int i = 0; i < 8; i++) { *p++ = (uint8_t)target; target >>= 8; }
    *p++ = (l2 >> 26) & 3;
    for (uint32_t y = 0; y < 144; y++) {
        uint32_t w0 = img[IMG_LP + 2 * y], w1 = img[IMG_LP + 2 * y + 1];
        *p++ = w0 & 0xFF; *p++ = (w0 >> 8) & 0xFF; *p++ = (w0 >> 16) & 0xFF; *p++ = w0 >> 24; *p++ = (w1 >> 4) & 1;
    }
    for (uint32_t i = 0; i < 144 * 160; i++) {
        uint32_t wv = SHADE_WORDS[(img[IMG_FB + (i >> 4)] >> (2 * (i & 15))) & 3];
        *p++ = (uint8_t)wv; *p++ = (uint8_t)(wv >> 8); *p++ = (uint8_t)(wv >> 16); *p++ = (uint8_t)(wv >> 24);
    }
    for (uint32_t i = 0; i < 0x2000; i++) *p++ = img_get_byte(img, IMG_MEM, MEM_WRAM + i);
    for (uint32_t i = 0; i < 96; i++) *p++ = img_get_byte(img, IMG_MEM, MEM_HI + 0xA0 + i);
    for (uint32_t i = 0; i < 76; i++) *p++ = img_get_byte(img, IMG_MEM, MEM_HI + 0x100 + i);
    for (uint32_t i = 0; i < 127; i++) *p++ = img_get_byte(img, IMG_MEM, MEM_HI + 0x180 + i);
    for (uint32_t i = 0; i < 52; i++) *p++ = img_get_byte(img, IMG_MEM, MEM_HI + 0x14C + i);
    uint32_t t = regs[R_TIMER];
    *p++ = t & 0xFF; *p++ = (t >> 8) & 0xFF;
    *p++ = regs[R_DIVC] & 0xFF; *p++ = (regs[R_DIVC] >> 8) & 0xFF; *p++ = regs[R_TIMAC] & 0xFF; *p++ = (regs[R_TIMAC] >> 8) & 0xFF;
    *p++ = (t >> 16) & 0xFF; *p++ = t >> 24;
    uint32_t mb = regs[R_MBC];
    *p++ = mb & 0xFF; *p++ = (mb >> 8) & 0xFF; *p++ = (mb >> 16) & 0xFF; *p++ = mb >> 24;
    for (uint32_t i = 0; i < 0x8000; i++) *p++ = img_get_byte(img, IMG_CRAM, i);
    *p++ = regs[R_JOY] & 0xFF; *p++ = (regs[R_JOY] >> 8) & 0xFF;
}

// fresh post-boot DMG machine (our convention; mirrors oracle gb_power_on)
static void power_on_image(std::vector<uint32_t> &img) {
    img.assign(IMG_WORDS, 0);
    uint32_t *regs = &img[IMG_REGS];
    regs[R_BCDE] = 0x13 | (0x00 << 8) | (0xD8 << 16) | (0x00u << 24);
    regs[R_HLAF] = 0x014D | (0x01 << 16) | (0xB0u << 24);
    regs[R_SPPC] = 0xFFFE | (0x0100u << 16);
    regs[R_INT] = 0;
    regs[R_LCD0] = 0x91 | (0x80 << 8);
    regs[R_LCD1] = 0;
    regs[R_LCD2] = 0xFC | (0xFF << 8) | (0xFF << 16) | ((0u | (2u << 2)) << 24);
    regs[R_MBC] = 1;
    regs[R_JOY] = 0x0F | (0x0F << 8) | (0xFFu << 16) | (144u << 24);
    regs[R_MISC] = 0xFF;
    img_set_byte(img, IMG_MEM, MEM_HI + 0x100, 0xFF);      // P1
    img_set_byte(img, IMG_MEM, MEM_HI + 0x14C + 4, 0x01);  // FF50
    for (uint32_t i = 0; i < FB_WORDS; i++) img[IMG_FB + i] = 0xFFFFFFFFu;  // PyBoy's fresh screen buffer is all zero words = black
    for (uint32_t y = 0; y < 144; y++) img[IMG_LP + 2 * y] = 7u << 16;      // fresh _scanlineparameters hold WX - 7 = 0
}

