// hostsim.cpp -- TEST INFRASTRUCTURE ONLY: the device headers of pokegym_b200/csrc compiled by g++ (-DGB_HOSTSIM, see
// gb_hd.h) and stepped one env at a time on the CPU, so that the interpreter / LCD / bus logic of the CUDA kernels can be
// compared with the oracle without a GPU (`pytest -m "not gpu"`, tests/test_hostsim.py) and debugged with ordinary tools.
// Nothing under pokegym_b200/ loads this library and libgbenv.so is never built with GB_HOSTSIM: the product has no
// CPU execution path.  Only k_run_frames' per-env body (run_frames_env), the state image scatter / gather and the
// save-state codec are exposed.
#define GB_HOSTSIM 1
#include "../../pokegym_b200/csrc/gb_hd.h"
#include <stdio.h>
#if defined(GB_TRACE)
static FILE *g_trace = nullptr;
static void gb_trace(uint32_t pc, uint32_t bcde, uint32_t hlaf, uint32_t sp, uint32_t dx) {
    if (g_trace) fprintf(g_trace, "%04x %08x %08x %04x %08x\n", pc, bcde, hlaf, sp, dx);
}
extern "C" void hs_trace(const char *path) { if (g_trace) fclose(g_trace); g_trace = path ? fopen(path, "w") : nullptr; }
#endif
#if defined(GB_SLOT_TRACE)
// convergence study (tools/trace_convergence.py): every executed instruction of every env as one 64-bit record
#include <vector>
static std::vector<unsigned long long> g_slots;
#define GB_TRACE_SLOT(kind, phys, dx, dw) g_slots.push_back(((unsigned long long)(kind) << 48) | ((unsigned long long)((dw) & 0xFFFFu) << 32) | (((dx) & 0xFFull) << 24) | ((phys) & 0xFFFFFFull))
extern "C" size_t hs_slots(unsigned long long *out, size_t cap) {
    size_t n = g_slots.size();
    if (out) { memcpy(out, g_slots.data(), sizeof(unsigned long long) * (n < cap ? n : cap)); g_slots.clear(); }
    return n;
}
#endif
#include "../../pokegym_b200/csrc/gb_image.h"
#include "../../pokegym_b200/csrc/gb_kernels.cuh"

struct HostSim {
    int n = 0, n_tiles = 0;
    DevArrays d{};
    std::vector<uint32_t> mem, cram, fb, lp, regs, dl;
    std::vector<uint8_t> rom;
    std::vector<uint4> rom_dec;
    unsigned long long counters[4] = {0, 0, 0, 0};
    bool simt = true;  // which build of the fast loop hs_run steps (hs_set_simt)
    bool defer = true;  // deferred PPU (hs_set_defer): record lines in the frame loop, draw them afterwards like k_render_pending
};

static void scatter_image(HostSim *h, int env, const std::vector<uint32_t> &img, int version) {
    for (uint32_t j = 0; j < IMG_WORDS; j++) {
        uint32_t *slot = image_slot(h->d, env, j);
        uint32_t v = img[j];
        if (j >= IMG_REGS) v = merge_loaded_reg(j - IMG_REGS, *slot, v, version, img[IMG_REGS + R_LCD0] & 0xFF);
        *slot = v;
    }
}

extern "C" {

void *hs_create(int n, const uint8_t *rom, size_t rom_len) {
    HostSim *h = new HostSim();
    h->n = n;
    h->n_tiles = (n + GB_TILE - 1) / GB_TILE;
    size_t T = (size_t)h->n_tiles * GB_TILE;
    h->mem.assign(T * MEM_WORDS, 0); h->cram.assign(T * CRAM_WORDS, 0); h->fb.assign(T * FB_WORDS, 0);
    h->lp.assign(T * LP_WORDS, 0); h->regs.assign(T * R_WORDS, 0); h->dl.assign(T * DL_WORDS, 0);
    h->rom.assign(rom, rom + rom_len);
    h->rom.resize(rom_len + 16, 0);
    h->d.mem = h->mem.data(); h->d.cram = h->cram.data(); h->d.fb = h->fb.data(); h->d.lp = h->lp.data(); h->d.regs = h->regs.data(); h->d.dl = h->dl.data();
    h->d.rom = h->rom.data();
    h->d.rom_banks = (uint32_t)(rom_len / 0x4000);
    h->d.n_envs = n; h->d.n_tiles = h->n_tiles;
    pd_build_base(c_base_desc);
    h->rom_dec.resize(rom_len);
    blockDim.x = 1; threadIdx.x = 0;
    for (uint32_t o = 0; o < rom_len; o++) { blockIdx.x = (int)o; k_predecode_rom(h->rom.data(), (uint32_t)rom_len, h->rom_dec.data()); }
    blockIdx.x = 0;
    h->d.rom_dec = h->rom_dec.data();
    std::vector<uint32_t> img;
    power_on_image(img);
    for (int e = 0; e < n; e++) scatter_image(h, e, img, 0);
    return h;
}

void hs_destroy(void *p) { delete (HostSim *)p; }
// number of per-opcode base descriptors of the fast set without a class id (gb_classes.inc out of date), and the class count
int hs_class_coverage(int *n_classes) {
    pd_desc_t t[512];
    pd_build_base(t);
    int missing = 0;
    for (int i = 0; i < 512; i++)
        if ((t[i].x & 0xFF) < H_RARE && PD_CLASS(t[i].w) >= GB_CLS_COUNT) missing++;
    *n_classes = GB_CLS_COUNT;
    return missing;
}
void hs_set_simt(void *p, int simt) { ((HostSim *)p)->simt = simt != 0; }
void hs_flush_stats(unsigned long long *out) { out[0] = g_hs_flushes; out[1] = g_hs_flushed_lines; }
void hs_set_defer(void *p, int defer) { ((HostSim *)p)->defer = defer != 0; }

int hs_load_blob(void *p, int env, const uint8_t *blob, size_t len) {
    HostSim *h = (HostSim *)p;
    std::vector<uint32_t> img;
    int ver = 0;
    std::string err;
    int rc = blob_to_image(blob, len, img, &ver, err);
    if (rc) return rc;
    scatter_image(h, env, img, ver);
    return 0;
}

int hs_save_blob(void *p, int env, uint8_t *out) {
    HostSim *h = (HostSim *)p;
    std::vector<uint32_t> img(IMG_WORDS);
    for (uint32_t j = 0; j < IMG_WORDS; j++) img[j] = *image_slot(h->d, env, j);
    image_to_blob(img, out);
    return 0;
}

// actions: uint8[n] or null; render_mode as RunParams (0 off, 1 every frame, 2 last frame only)
int hs_run(void *p, const uint8_t *actions, int n_frames, int render_mode) {
    HostSim *h = (HostSim *)p;
    RunParams rp;
    rp.d = h->d; rp.actions = actions; rp.skip = nullptr; rp.n_frames = n_frames; rp.render_mode = render_mode; rp.release_frame = 8; rp.lanes = 1; rp.defer = h->defer && render_mode != 0;
    rp.bank_mask = (h->d.rom_banks & (h->d.rom_banks - 1)) == 0 ? h->d.rom_banks - 1 : 0;
    rp.counters = h->counters;
    for (int env = 0; env < h->n; env++) {
        EnvSlot slot;
        uint32_t line[FB_LINE_WORDS], keys[10];
        GB_TRACE_SLOT(3, 0, 0, 0);  // a new env's records start here
        machine_load(slot.m, rp.d, env >> 5, env & 31);
        const int button = actions ? c_action_button[actions[env] & 7] : -1;
        slot.m.rline = line; slot.m.rkeys = keys; slot.m.rls = 1;
        if (rp.defer) slot.m.dl = rp.d.dl + il_index(env >> 5, DL_WORDS, 0, env & 31);
        if (h->simt) run_frames_env<true>(slot.m, rp, button);
        else run_frames_env<false>(slot.m, rp, button);
        machine_store(slot.m, rp.d, env >> 5, env & 31);
        h->counters[0] += slot.m.n_instr; h->counters[1] += slot.m.n_cycles; h->counters[2] += n_frames;
    }
    if (rp.defer) {  // k_render_pending
        for (int env = 0; env < h->n; env++) {
            uint32_t line[FB_LINE_WORDS], keys[10];
            render_pending_lines(rp.d, env >> 5, env & 31, 0, 1, line, keys, 1);
        }
    }
    return 0;
}

void hs_counters(void *p, unsigned long long *out) { memcpy(out, ((HostSim *)p)->counters, sizeof(unsigned long long) * 4); }

int hs_core_extra(void *p, int env, int *stat_mode, int *ly_window, int *fault) {
    HostSim *h = (HostSim *)p;
    uint32_t lcd2 = h->regs[il_index(env >> 5, R_WORDS, R_LCD2, env & 31)], joy = h->regs[il_index(env >> 5, R_WORDS, R_JOY, env & 31)];
    uint32_t in = h->regs[il_index(env >> 5, R_WORDS, R_INT, env & 31)];
    *stat_mode = (lcd2 >> 24) & 3; *ly_window = (int)(int8_t)((joy >> 16) & 0xFF); *fault = (in >> 4) & 1;
    return 0;
}
}
