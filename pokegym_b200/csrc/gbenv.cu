// gbenv.cu -- host side of libgbenv.so: the C ABI of include/gbenv.h over the CUDA kernels.
// There is no CPU fallback anywhere in this file: every entry point either launches kernels on the
// handle's device or fails with GBENV_E_CUDA.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/gbenv.h"
#include "gb_image.h"
#include "gb_kernels.cuh"
#include "gb_wrap.cuh"

struct StateTemplate {
    uint32_t *d_image;  // IMG_WORDS words, canonical per-env image
    int version;
};

struct gbenv {
    int n = 0, n_tiles = 0, device = 0;
    DevArrays d{};
    WrapArrays w{};
    cudaStream_t stream = nullptr;
    std::vector<StateTemplate> templates;
    std::vector<int> initial_template;  // per env, -1 = none
    std::vector<int> reset_count;       // host mirror of WrapState.reset_count
    unsigned long long *d_counters = nullptr;
    uint32_t *d_stage_image = nullptr;
    uint8_t *d_stage_buf = nullptr;  // 64 KiB scratch for bus access
    double *d_info_rows = nullptr;
    int32_t *d_env_ids = nullptr;
    uint8_t *d_mask = nullptr;
    // staging for the *_host entry points
    uint8_t *d_actions = nullptr, *d_obs = nullptr, *d_done = nullptr;
    double *d_reward = nullptr;
    // CUDA-event ring: slot k holds {start, after k_run_frames, after k_wrap_*} of step k (mod EV_RING)
    static const int EV_RING = 64;
    cudaEvent_t ev[EV_RING][3] = {};
    unsigned long long ev_step = 0, ev_folded = 0;  // steps recorded / steps folded into the totals
    double ms_total[2] = {0.0, 0.0};
    bool ev_valid = false;
    unsigned long long launches = 0;
    int lanes = 32;  // envs per warp in k_run_frames
    size_t smem_opted = 48 * 1024;
    std::string err;
};

static std::string g_err;

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t _e = (call);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            char _b[512];                                                                              \
            snprintf(_b, sizeof(_b), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            if (h) h->err = _b; else g_err = _b;                                                       \
            return GBENV_E_CUDA;                                                                       \
        }                                                                                              \
    } while (0)

static int fail(gbenv *h, int code, const char *msg) {
    if (h) h->err = msg; else g_err = msg;
    return code;
}

// NULL means the legacy default stream -- the stream PyTorch uses unless told otherwise -- so a caller that
// passes nothing is ordered after its own tensor copies.  The handle's private stream is a blocking stream,
// i.e. implicitly ordered with the legacy default stream in both directions.
static cudaStream_t pick(gbenv *h, void *stream) { (void)h; return (cudaStream_t)stream; }

// ------------------------------------------------------------------------------- lifetime

extern "C" int gbenv_abi_version(void) { return GBENV_ABI_VERSION; }

extern "C" const char *gbenv_last_error(const gbenv *h) { return h ? h->err.c_str() : g_err.c_str(); }

extern "C" int gbenv_num_envs(const gbenv *h) { return h ? h->n : GBENV_E_ARG; }

static int scatter(gbenv *h, const uint32_t *d_image, int version, const int32_t *env_ids_host, int n, cudaStream_t st) {
    const int32_t *d_ids = nullptr;
    if (env_ids_host) {
        for (int i = 0; i < n; i++)
            if (env_ids_host[i] < 0 || env_ids_host[i] >= h->n) return fail(h, GBENV_E_ARG, "env id out of range");
        CK(cudaMemcpyAsync(h->d_env_ids, env_ids_host, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
        d_ids = h->d_env_ids;
    } else {
        n = h->n;
    }
    if (n <= 0) return GBENV_OK;
    dim3 block(32, 8), grid((n + 31) / 32, (IMG_WORDS + 7) / 8);
    k_scatter_image<<<grid, block, 0, st>>>(h->d, d_image, d_ids, n, version);
    h->launches++;
    CK(cudaGetLastError());
    if (env_ids_host) CK(cudaStreamSynchronize(st));  // d_env_ids is reused by the next call
    return GBENV_OK;
}

extern "C" int gbenv_destroy(gbenv *h) {
    if (!h) return GBENV_E_ARG;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (auto &t : h->templates) cudaFree(t.d_image);
    cudaFree(h->d.mem); cudaFree(h->d.cram); cudaFree(h->d.fb); cudaFree(h->d.lp); cudaFree(h->d.regs); cudaFree((void *)h->d.rom); cudaFree((void *)h->d.rom_dec);
    cudaFree(h->w.state); cudaFree(h->w.visited); cudaFree(h->w.counts_map); cudaFree(h->w.cm_hash);
    cudaFree(h->d_counters); cudaFree(h->d_stage_image); cudaFree(h->d_stage_buf); cudaFree(h->d_info_rows);
    cudaFree(h->d_env_ids); cudaFree(h->d_mask); cudaFree(h->d_actions); cudaFree(h->d_obs); cudaFree(h->d_done); cudaFree(h->d_reward);
    for (auto &slot : h->ev)
        for (auto &e : slot)
            if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return GBENV_OK;
}

extern "C" int gbenv_create(int n_envs, const uint8_t *rom_host, size_t rom_len, int device_id, gbenv **out) {
    gbenv *h = nullptr;
    if (n_envs <= 0 || !rom_host || rom_len < 0x8000 || (rom_len & 0x3FFF) || !out) return fail(nullptr, GBENV_E_ARG, "gbenv_create: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, GBENV_E_CUDA, "gbenv_create: no CUDA device (this library has no CPU fallback)");
    if (device_id < 0 || device_id >= ndev) return fail(nullptr, GBENV_E_ARG, "gbenv_create: bad device id");
    CK(cudaSetDevice(device_id));
    h = new gbenv();
    h->n = n_envs;
    h->n_tiles = (n_envs + GB_TILE - 1) / GB_TILE;
    h->device = device_id;
    h->initial_template.assign(n_envs, -1);
    h->reset_count.assign(n_envs, 0);
#define ALLOC(ptr, bytes)                                              \
    do {                                                               \
        if (cudaMalloc((void **)&(ptr), (bytes)) != cudaSuccess) {     \
            g_err = "gbenv_create: cudaMalloc failed for " #ptr;       \
            gbenv_destroy(h);                                          \
            return GBENV_E_NOMEM;                                      \
        }                                                              \
        cudaMemset((ptr), 0, (bytes));                                 \
    } while (0)
    size_t T = (size_t)h->n_tiles * GB_TILE * sizeof(uint32_t);
    ALLOC(h->d.mem, T * MEM_WORDS);
    ALLOC(h->d.cram, T * CRAM_WORDS);
    ALLOC(h->d.fb, T * FB_WORDS);
    ALLOC(h->d.lp, T * LP_WORDS);
    ALLOC(h->d.regs, T * R_WORDS);
    uint8_t *d_rom = nullptr;
    ALLOC(d_rom, rom_len + 16);  // padded: the instruction fetch reads two aligned words
    h->d.rom = d_rom;
    uint4 *d_rom_dec = nullptr;
    ALLOC(d_rom_dec, rom_len * sizeof(uint4));
    h->d.rom_dec = d_rom_dec;
    h->d.rom_banks = (uint32_t)(rom_len / 0x4000);
    h->d.n_envs = n_envs;
    h->d.n_tiles = h->n_tiles;
    // visited-bitmap slots per env: all 248 maps when that fits an 8 GiB budget, otherwise fewer
    // (an env that visits more maps than it has slots raises the `faults` counter; GBENV_VISITED_SLOTS overrides)
    int slots = WRAP_MAPS;
    size_t budget = (size_t)8 << 30, per_slot = (size_t)VIS_MAP_WORDS * 4;
    if ((size_t)n_envs * slots * per_slot > budget) slots = (int)(budget / ((size_t)n_envs * per_slot));
    if (slots < 8) slots = 8;
    if (const char *ev = getenv("GBENV_VISITED_SLOTS")) {
        int v = atoi(ev);
        if (v >= 1 && v <= WRAP_MAPS) slots = v;
    }
    h->w.slots = slots;
    {   // envs per warp: the interpreter is latency-bound, so spread the batch over as many warps as are resident
        // at once (one wave: SMs x STEP_MIN_BLOCKS blocks) and no more; GBENV_LANES / gbenv_set_lanes_per_warp override.
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_id);
        const int resident_warps = sms * STEP_MIN_BLOCKS * (STEP_THREADS / 32);
        // measured on B200 (tools/quick_bench.sh): the smallest power of two that keeps every warp resident wins; counts
        // that are not powers of two make warps straddle 32-env tiles (partial 128-byte lines) and lose 3-5 %
        int lanes = 1;
        while (lanes < 32 && (n_envs + lanes - 1) / lanes > resident_warps) lanes <<= 1;
        if (const char *ev = getenv("GBENV_LANES")) {
            int v = atoi(ev);
            if (v >= 1 && v <= 32) lanes = v;
        }
        h->lanes = lanes;
    }
    ALLOC(h->w.state, sizeof(WrapState) * (size_t)n_envs);
    ALLOC(h->w.visited, (size_t)n_envs * slots * per_slot);
    {   // heat maps (environment.py:648-679): dense 444x436 int32 per env while they fit in 4 GiB, else a per-env hash of the
        // cells actually touched (one new cell per step at most), sized to a 2 GiB budget; GBENV_COUNTS_MAP=dense|sparse|0
        bool dense = (size_t)n_envs * COUNTS_H * COUNTS_W * 4 <= ((size_t)4 << 30), sparse = !dense;
        if (const char *ev = getenv("GBENV_COUNTS_MAP")) {
            dense = !strcmp(ev, "dense") || !strcmp(ev, "1");
            sparse = !strcmp(ev, "sparse");
        }
        if (dense) ALLOC(h->w.counts_map, (size_t)n_envs * COUNTS_H * COUNTS_W * sizeof(int32_t));
        if (sparse) {
            int cap = 65536;
            while (cap > 1024 && (size_t)n_envs * cap * sizeof(uint2) > ((size_t)2 << 30)) cap >>= 1;
            if (const char *ev = getenv("GBENV_COUNTS_SLOTS")) {
                int v = atoi(ev);
                if (v >= 16 && v <= (1 << 18) && (v & (v - 1)) == 0) cap = v;
            }
            h->w.cm_cap = cap;
            ALLOC(h->w.cm_hash, (size_t)n_envs * cap * sizeof(uint2));
        }
    }
    ALLOC(h->d_counters, 8 * sizeof(unsigned long long));
    ALLOC(h->d_stage_image, IMG_WORDS * sizeof(uint32_t));
    ALLOC(h->d_stage_buf, 0x10000);
    ALLOC(h->d_info_rows, (size_t)n_envs * GBENV_INFO_SCALARS * sizeof(double));
    ALLOC(h->d_env_ids, (size_t)n_envs * sizeof(int32_t));
    ALLOC(h->d_mask, (size_t)n_envs);
#undef ALLOC
    CK(cudaStreamCreate(&h->stream));
    for (auto &slot : h->ev)
        for (auto &e : slot) CK(cudaEventCreate(&e));
    CK(cudaMemcpy(d_rom, rom_host, rom_len, cudaMemcpyHostToDevice));
    {   // decode the shared ROM once (gb_predecode.h): 16 bytes per ROM offset, L2-resident
        static uint4 base[512];
        pd_build_base(base);
        CK(cudaMemcpyToSymbol(c_base_desc, base, sizeof(base)));
        k_predecode_rom<<<(unsigned)((rom_len + 255) / 256), 256, 0, h->stream>>>(d_rom, (uint32_t)rom_len, d_rom_dec);
        CK(cudaGetLastError());
    }
    k_wrap_init<<<(n_envs + 127) / 128, 128, 0, h->stream>>>(h->w, n_envs);
    CK(cudaGetLastError());
    std::vector<uint32_t> img;
    power_on_image(img);
    CK(cudaMemcpyAsync(h->d_stage_image, img.data(), IMG_WORDS * 4, cudaMemcpyHostToDevice, h->stream));
    int rc = scatter(h, h->d_stage_image, 0 /* raw image, no load_state merge */, nullptr, 0, h->stream);
    if (rc) { g_err = h->err; gbenv_destroy(h); return rc; }
    CK(cudaStreamSynchronize(h->stream));
    *out = h;
    return GBENV_OK;
}

extern "C" int gbenv_set_lanes_per_warp(gbenv *h, int lanes) {
    if (!h || lanes < 1 || lanes > 32) return fail(h, GBENV_E_ARG, "gbenv_set_lanes_per_warp: lanes must be in 1..32");
    h->lanes = lanes;
    return GBENV_OK;
}

extern "C" int gbenv_sync(gbenv *h) {
    if (!h) return GBENV_E_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return GBENV_OK;
}

// ------------------------------------------------------------------------------- templates / state

extern "C" int gbenv_add_state_template(gbenv *h, const uint8_t *blob, size_t len, int *id_out) {
    if (!h || !blob || !id_out) return fail(h, GBENV_E_ARG, "gbenv_add_state_template: bad argument");
    CK(cudaSetDevice(h->device));
    std::vector<uint32_t> img;
    int ver = 0;
    std::string err;
    int rc = blob_to_image(blob, len, img, &ver, err);
    if (rc) return fail(h, rc, err.c_str());
    StateTemplate t{nullptr, ver};
    CK(cudaMalloc((void **)&t.d_image, IMG_WORDS * sizeof(uint32_t)));
    CK(cudaMemcpy(t.d_image, img.data(), IMG_WORDS * sizeof(uint32_t), cudaMemcpyHostToDevice));
    h->templates.push_back(t);
    *id_out = (int)h->templates.size() - 1;
    return GBENV_OK;
}

extern "C" int gbenv_load_template(gbenv *h, const int32_t *env_ids, int n, int tid) {
    if (!h || tid < 0 || tid >= (int)h->templates.size()) return fail(h, GBENV_E_ARG, "gbenv_load_template: bad template id");
    CK(cudaSetDevice(h->device));
    return scatter(h, h->templates[tid].d_image, h->templates[tid].version, env_ids, n, h->stream);
}

extern "C" int gbenv_set_initial_template(gbenv *h, const int32_t *env_ids, int n, int tid) {
    if (!h || tid < 0 || tid >= (int)h->templates.size()) return fail(h, GBENV_E_ARG, "gbenv_set_initial_template: bad template id");
    if (env_ids) {
        for (int i = 0; i < n; i++) {
            if (env_ids[i] < 0 || env_ids[i] >= h->n) return fail(h, GBENV_E_ARG, "env id out of range");
            h->initial_template[env_ids[i]] = tid;
        }
    } else {
        for (auto &t : h->initial_template) t = tid;
    }
    return GBENV_OK;
}

extern "C" int gbenv_power_on(gbenv *h, const int32_t *env_ids, int n) {
    if (!h) return GBENV_E_ARG;
    CK(cudaSetDevice(h->device));
    // a power-on is a complete re-initialisation: unlike a state load it also resets STAT mode / ly_window
    std::vector<uint32_t> img;
    power_on_image(img);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpyAsync(h->d_stage_image, img.data(), IMG_WORDS * 4, cudaMemcpyHostToDevice, h->stream));
    return scatter(h, h->d_stage_image, 0 /* version 0 = raw, no merge */, env_ids, n, h->stream);
}

extern "C" int gbenv_save_state(gbenv *h, int env, uint8_t *blob) {
    if (!h || env < 0 || env >= h->n || !blob) return fail(h, GBENV_E_ARG, "gbenv_save_state: bad argument");
    CK(cudaSetDevice(h->device));
    k_gather_image<<<(IMG_WORDS + 255) / 256, 256, 0, h->stream>>>(h->d, h->d_stage_image, env);
    h->launches++;
    CK(cudaGetLastError());
    std::vector<uint32_t> img(IMG_WORDS);
    CK(cudaMemcpyAsync(img.data(), h->d_stage_image, IMG_WORDS * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    image_to_blob(img, blob);
    return GBENV_OK;
}

// ------------------------------------------------------------------------------- emulator

static int launch_run(gbenv *h, const uint8_t *actions_dev, int n_frames, int render_mode, cudaStream_t st) {
    RunParams p;
    p.d = h->d;
    p.actions = actions_dev;
    p.n_frames = n_frames;
    p.render_mode = render_mode;
    p.release_frame = 8;
    p.counters = h->d_counters;
    p.lanes = h->lanes;
    p.bank_mask = (h->d.rom_banks & (h->d.rom_banks - 1)) == 0 ? h->d.rom_banks - 1 : 0;
    int warps = (h->n + h->lanes - 1) / h->lanes;
    int blocks = (warps * 32 + STEP_THREADS - 1) / STEP_THREADS;
    size_t smem = ENV_SMEM_BYTES((STEP_THREADS / 32) * h->lanes);
    if (smem > h->smem_opted) {  // beyond the 48 KiB default: opt in once per size
        CK(cudaFuncSetAttribute(k_run_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->smem_opted = smem;
    }
    k_run_frames<<<blocks, STEP_THREADS, smem, st>>>(p);
    h->launches++;
    CK(cudaGetLastError());
    return GBENV_OK;
}

extern "C" int gbenv_run_action(gbenv *h, const uint8_t *actions_dev, int frame_skip, void *stream) {
    if (!h || !actions_dev || frame_skip <= 0) return fail(h, GBENV_E_ARG, "gbenv_run_action: bad argument");
    CK(cudaSetDevice(h->device));
    return launch_run(h, actions_dev, frame_skip, 2, pick(h, stream));
}

extern "C" int gbenv_tick(gbenv *h, int n_frames, int render, void *stream) {
    if (!h || n_frames <= 0) return fail(h, GBENV_E_ARG, "gbenv_tick: bad argument");
    CK(cudaSetDevice(h->device));
    return launch_run(h, nullptr, n_frames, render ? 1 : 0, pick(h, stream));
}

extern "C" int gbenv_send_input(gbenv *h, int button, int pressed, void *stream) {
    if (!h || button < 0 || button > 7) return fail(h, GBENV_E_ARG, "gbenv_send_input: bad argument");
    CK(cudaSetDevice(h->device));
    k_send_input<<<(h->n + 127) / 128, 128, 0, pick(h, stream)>>>(h->d, button, pressed);
    h->launches++;
    CK(cudaGetLastError());
    return GBENV_OK;
}

static int bus_access(gbenv *h, int env, uint32_t addr, uint32_t n, uint8_t *host, int write) {
    if (!h || env < 0 || env >= h->n || !host || addr + n > 0x10000 || n == 0) return fail(h, GBENV_E_ARG, "bus access: bad argument");
    CK(cudaSetDevice(h->device));
    if (write) CK(cudaMemcpyAsync(h->d_stage_buf, host, n, cudaMemcpyHostToDevice, h->stream));
    k_bus_access<<<1, 1, 0, h->stream>>>(h->d, env, addr, n, h->d_stage_buf, write);
    h->launches++;
    CK(cudaGetLastError());
    if (!write) CK(cudaMemcpyAsync(host, h->d_stage_buf, n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return GBENV_OK;
}

extern "C" int gbenv_read_mem(gbenv *h, int env, uint32_t addr, uint32_t n, uint8_t *out) { return bus_access(h, env, addr, n, out, 0); }
extern "C" int gbenv_write_mem(gbenv *h, int env, uint32_t addr, uint32_t n, const uint8_t *in) {
    return bus_access(h, env, addr, n, (uint8_t *)in, 1);
}

extern "C" int gbenv_screen(gbenv *h, int env, uint8_t *rgb) {
    if (!h || env < 0 || env >= h->n || !rgb) return fail(h, GBENV_E_ARG, "gbenv_screen: bad argument");
    CK(cudaSetDevice(h->device));
    k_gather_image<<<(IMG_WORDS + 255) / 256, 256, 0, h->stream>>>(h->d, h->d_stage_image, env);
    h->launches++;
    CK(cudaGetLastError());
    std::vector<uint32_t> fb(FB_WORDS);
    CK(cudaMemcpyAsync(fb.data(), h->d_stage_image + IMG_FB, FB_WORDS * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    static const uint8_t grey[4] = {0xFF, 0x99, 0x55, 0x00};
    for (uint32_t i = 0; i < 144 * 160; i++) {
        uint8_t g = grey[(fb[i >> 4] >> (2 * (i & 15))) & 3];
        rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = g;
    }
    return GBENV_OK;
}

// ------------------------------------------------------------------------------- Environment API

extern "C" int gbenv_reset(gbenv *h, const uint8_t *mask_host, int max_episode_steps, double reward_scale, uint8_t *obs_dev, size_t obs_stride,
                           void *stream) {
    if (!h || !obs_dev || obs_stride < GBENV_OBS_BYTES || (obs_stride & 3)) return fail(h, GBENV_E_ARG, "gbenv_reset: bad argument");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = pick(h, stream);
    const uint8_t *d_mask = nullptr;
    if (mask_host) {
        CK(cudaMemcpyAsync(h->d_mask, mask_host, h->n, cudaMemcpyHostToDevice, st));
        d_mask = h->d_mask;
    }
    int blocks = (h->n + 127) / 128;
    k_wrap_reset_pre<<<blocks, 128, 0, st>>>(h->d, h->w, d_mask);  // environment.py:1239: D778 |= 0x10 before the load
    h->launches++;
    CK(cudaGetLastError());
    // environment.py:1241-1242: the save-state is loaded on an env's first reset only
    std::vector<std::vector<int32_t>> by_template(h->templates.size());
    for (int e = 0; e < h->n; e++) {
        if (mask_host && !mask_host[e]) continue;
        if (h->reset_count[e] == 0 && h->initial_template[e] >= 0) by_template[h->initial_template[e]].push_back(e);
        h->reset_count[e]++;
    }
    for (size_t t = 0; t < by_template.size(); t++) {
        auto &ids = by_template[t];
        if (ids.empty()) continue;
        bool all = (int)ids.size() == h->n;
        int rc = scatter(h, h->templates[t].d_image, h->templates[t].version, all ? nullptr : ids.data(), (int)ids.size(), st);
        if (rc) return rc;
    }
    k_wrap_reset_post<<<blocks, 128, 0, st>>>(h->d, h->w, d_mask, max_episode_steps, reward_scale);
    k_wrap_obs<<<h->n_tiles, 256, 0, st>>>(h->d, h->w, d_mask, obs_dev, obs_stride);
    h->launches += 2;
    CK(cudaGetLastError());
    if (mask_host) CK(cudaStreamSynchronize(st));  // d_mask is reused
    return GBENV_OK;
}

// accumulate the device time of recorded steps [ev_folded, upto) into ms_total (blocks on their events)
static int fold_events(gbenv *h, unsigned long long upto) {
    while (h->ev_folded < upto && h->ev_folded < h->ev_step) {
        cudaEvent_t *ev = h->ev[h->ev_folded % gbenv::EV_RING];
        CK(cudaEventSynchronize(ev[2]));
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, ev[0], ev[1]));
        CK(cudaEventElapsedTime(&b, ev[1], ev[2]));
        h->ms_total[0] += a;
        h->ms_total[1] += b;
        h->ev_folded++;
    }
    return GBENV_OK;
}

extern "C" int gbenv_step(gbenv *h, const uint8_t *actions_dev, uint8_t *obs_dev, size_t obs_stride, double *reward_dev, uint8_t *done_dev,
                          void *stream) {
    if (!h || !actions_dev || !obs_dev || !reward_dev || !done_dev || obs_stride < GBENV_OBS_BYTES || (obs_stride & 3))
        return fail(h, GBENV_E_ARG, "gbenv_step: bad argument");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = pick(h, stream);
    if (h->ev_step - h->ev_folded >= (unsigned long long)gbenv::EV_RING) {  // slot about to be reused: fold it first
        int rcf = fold_events(h, h->ev_folded + 1);
        if (rcf) return rcf;
    }
    cudaEvent_t *ev = h->ev[h->ev_step % gbenv::EV_RING];
    CK(cudaEventRecord(ev[0], st));
    int rc = launch_run(h, actions_dev, GBENV_ACT_FREQ, 2, st);
    if (rc) return rc;
    CK(cudaEventRecord(ev[1], st));
    int blocks = (h->n_tiles * GB_TILE + 127) / 128;
    k_wrap_step<<<blocks, 128, 0, st>>>(h->d, h->w, reward_dev, done_dev, h->d_info_rows);
    k_wrap_obs<<<h->n_tiles, 256, 0, st>>>(h->d, h->w, nullptr, obs_dev, obs_stride);
    h->launches += 2;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev[2], st));
    h->ev_step++;
    h->ev_valid = true;
    return GBENV_OK;
}

static int ensure_host_staging(gbenv *h) {
    if (h->d_actions) return GBENV_OK;
    CK(cudaMalloc((void **)&h->d_actions, h->n));
    CK(cudaMalloc((void **)&h->d_obs, (size_t)h->n * GBENV_OBS_BYTES));
    CK(cudaMalloc((void **)&h->d_done, h->n));
    CK(cudaMalloc((void **)&h->d_reward, (size_t)h->n * sizeof(double)));
    return GBENV_OK;
}

extern "C" int gbenv_step_host(gbenv *h, const uint8_t *actions, uint8_t *obs, double *reward, uint8_t *done) {
    if (!h || !actions || !obs || !reward || !done) return fail(h, GBENV_E_ARG, "gbenv_step_host: bad argument");
    CK(cudaSetDevice(h->device));
    int rc = ensure_host_staging(h);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->d_actions, actions, h->n, cudaMemcpyHostToDevice, h->stream));
    rc = gbenv_step(h, h->d_actions, h->d_obs, GBENV_OBS_BYTES, h->d_reward, h->d_done, h->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(obs, h->d_obs, (size_t)h->n * GBENV_OBS_BYTES, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(reward, h->d_reward, (size_t)h->n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(done, h->d_done, h->n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return GBENV_OK;
}

extern "C" int gbenv_reset_host(gbenv *h, const uint8_t *mask, int max_episode_steps, double reward_scale, uint8_t *obs) {
    if (!h || !obs) return fail(h, GBENV_E_ARG, "gbenv_reset_host: bad argument");
    CK(cudaSetDevice(h->device));
    int rc = ensure_host_staging(h);
    if (rc) return rc;
    // rows of envs that are not reset keep the caller's bytes: stage the caller's buffer in first
    if (mask) CK(cudaMemcpyAsync(h->d_obs, obs, (size_t)h->n * GBENV_OBS_BYTES, cudaMemcpyHostToDevice, h->stream));
    rc = gbenv_reset(h, mask, max_episode_steps, reward_scale, h->d_obs, GBENV_OBS_BYTES, h->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(obs, h->d_obs, (size_t)h->n * GBENV_OBS_BYTES, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return GBENV_OK;
}

extern "C" int gbenv_get_info(gbenv *h, double *info_dev, void *stream) {
    if (!h || !info_dev) return fail(h, GBENV_E_ARG, "gbenv_get_info: bad argument");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(info_dev, h->d_info_rows, (size_t)h->n * GBENV_INFO_SCALARS * sizeof(double), cudaMemcpyDeviceToDevice, pick(h, stream)));
    return GBENV_OK;
}

extern "C" int gbenv_reduce_info(gbenv *h, double *sum_dev, void *stream) {
    if (!h || !sum_dev) return fail(h, GBENV_E_ARG, "gbenv_reduce_info: bad argument");
    CK(cudaSetDevice(h->device));
    k_reduce_info<<<GBENV_INFO_SCALARS, 256, 0, pick(h, stream)>>>(h->d_info_rows, h->n, sum_dev);
    h->launches++;
    CK(cudaGetLastError());
    return GBENV_OK;
}

extern "C" int gbenv_counts_map(gbenv *h, int env, int32_t *map_host) {
    if (!h || env < 0 || env >= h->n || !map_host) return fail(h, GBENV_E_ARG, "gbenv_counts_map: bad argument");
    if (!h->w.counts_map && !h->w.cm_hash) return fail(h, GBENV_E_ARG, "gbenv_counts_map: exploration map tracking is disabled (GBENV_COUNTS_MAP=0)");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    if (!h->w.counts_map) {  // sparse: rebuild the dense image from this env's hash entries
        std::vector<uint2> tab((size_t)h->w.cm_cap);
        CK(cudaMemcpy(tab.data(), h->w.cm_hash + (size_t)env * h->w.cm_cap, tab.size() * sizeof(uint2), cudaMemcpyDeviceToHost));
        memset(map_host, 0, (size_t)COUNTS_H * COUNTS_W * 4);
        for (const uint2 &e : tab)
            if (e.x) map_host[e.x - 1] = (int32_t)e.y;
        return GBENV_OK;
    }
    CK(cudaMemcpy(map_host, h->w.counts_map + (size_t)env * COUNTS_H * COUNTS_W, (size_t)COUNTS_H * COUNTS_W * 4, cudaMemcpyDeviceToHost));
    return GBENV_OK;
}

// ------------------------------------------------------------------------------- diagnostics

__global__ void k_count_faults(DevArrays d, WrapArrays w, unsigned long long *out) {
    int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= d.n_envs) return;
    uint32_t in = d.regs[il_index(env >> 5, R_WORDS, R_INT, env & 31)];
    if (((in >> 4) & 1) || w.state[env].overflow) atomicAdd(out, 1ull);
}

extern "C" int gbenv_get_counters(gbenv *h, gbenv_counters_t *out) {
    if (!h || !out) return fail(h, GBENV_E_ARG, "gbenv_get_counters: bad argument");
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(h->d_counters + 3, 0, sizeof(unsigned long long), h->stream));
    k_count_faults<<<(h->n + 127) / 128, 128, 0, h->stream>>>(h->d, h->w, h->d_counters + 3);
    CK(cudaGetLastError());
    unsigned long long c[4];
    CK(cudaMemcpyAsync(c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    out->instructions = c[0];
    out->cycles = c[1];
    out->frames = c[2];
    out->faults = c[3];
    out->kernel_launches = h->launches;
    return GBENV_OK;
}

extern "C" int gbenv_last_kernel_ms(gbenv *h, int which, float *ms) {
    if (!h || !ms || which < 0 || which > 1) return fail(h, GBENV_E_ARG, "gbenv_last_kernel_ms: bad argument");
    if (!h->ev_valid || h->ev_step == 0) return fail(h, GBENV_E_ARG, "gbenv_last_kernel_ms: no step recorded yet");
    CK(cudaSetDevice(h->device));
    cudaEvent_t *ev = h->ev[(h->ev_step - 1) % gbenv::EV_RING];
    CK(cudaEventSynchronize(ev[2]));
    CK(cudaEventElapsedTime(ms, ev[which], ev[which + 1]));
    return GBENV_OK;
}

extern "C" int gbenv_kernel_time_total(gbenv *h, int which, double *ms_total, uint64_t *steps) {
    if (!h || which < 0 || which > 1 || !ms_total || !steps) return fail(h, GBENV_E_ARG, "gbenv_kernel_time_total: bad argument");
    CK(cudaSetDevice(h->device));
    int rc = fold_events(h, h->ev_step);
    if (rc) return rc;
    *ms_total = h->ms_total[which];
    *steps = h->ev_folded;
    return GBENV_OK;
}

extern "C" int gbenv_get_core_extra(gbenv *h, int env, gbenv_core_extra_t *out) {
    if (!h || env < 0 || env >= h->n || !out) return fail(h, GBENV_E_ARG, "gbenv_get_core_extra: bad argument");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    uint32_t lcd2 = 0, joy = 0, in = 0;
    int tile = env >> 5, lane = env & 31;
    CK(cudaMemcpy(&lcd2, h->d.regs + il_index(tile, R_WORDS, R_LCD2, lane), 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&joy, h->d.regs + il_index(tile, R_WORDS, R_JOY, lane), 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&in, h->d.regs + il_index(tile, R_WORDS, R_INT, lane), 4, cudaMemcpyDeviceToHost));
    out->stat_mode = (lcd2 >> 24) & 3;
    out->ly_window = (int)(int8_t)((joy >> 16) & 0xFF);
    out->fault = (in >> 4) & 1;
    out->reserved = 0;
    return GBENV_OK;
}

extern "C" int gbenv_debug_render_frame(gbenv *h, int env) {
    if (!h || env < 0 || env >= h->n) return fail(h, GBENV_E_ARG, "gbenv_debug_render_frame: bad argument");
    CK(cudaSetDevice(h->device));
    k_debug_render_frame<<<1, 1, 0, h->stream>>>(h->d, env);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return GBENV_OK;
}
