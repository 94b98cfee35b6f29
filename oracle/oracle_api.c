/*
 * oracle/oracle_api.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The CPU oracle behind the same entry points as include/gbenv.h, prefixed `oracle_`, all pointers
 * host memory.  Envs are independent GbCore instances (gb_core.c) plus the wrapper restatement
 * (pokegym_wrapper.c); loops over envs are OpenMP-parallel so bench.py's cpu_baseline / reference
 * legs can use every host core.  Only tests/, smoke() and bench.py's baseline legs may load this.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gbenv.h"
#include "gb_core.h"
#include "pokegym_wrapper.h"

#include <pthread.h>
#include <unistd.h>

typedef struct oracle_env {
    GbCore core;
    PgWrapper wrap;
    int initial_template;
} oracle_env;

typedef struct par_pool par_pool;
typedef struct oracle {
    int n;
    uint8_t *rom;
    size_t rom_len;
    oracle_env *envs;
    uint8_t **templates;
    size_t *template_len;
    int n_templates;
    int threads;
    struct par_pool *pool; /* persistent worker threads (par_for) */
    char err[256];
} oracle;

static char g_err[256] = "";
static void par_pool_stop(oracle *h);

static int fail(oracle *h, int code, const char *msg) {
    snprintf(h ? h->err : g_err, 256, "%s", msg);
    return code;
}

int oracle_abi_version(void) { return GBENV_ABI_VERSION; }

int oracle_create(int n_envs, const uint8_t *rom, size_t rom_len, int device_id, oracle **out) {
    (void)device_id;
    if (n_envs <= 0 || !rom || rom_len < 0x8000 || !out) return fail(NULL, GBENV_E_ARG, "oracle_create: bad argument");
    oracle *h = (oracle *)calloc(1, sizeof(oracle));
    if (!h) return GBENV_E_NOMEM;
    h->n = n_envs;
    h->rom = (uint8_t *)malloc(rom_len);
    h->rom_len = rom_len;
    h->envs = (oracle_env *)calloc((size_t)n_envs, sizeof(oracle_env));
    if (!h->rom || !h->envs) return fail(NULL, GBENV_E_NOMEM, "oracle_create: out of memory");
    memcpy(h->rom, rom, rom_len);
    for (int e = 0; e < n_envs; e++) {
        gb_power_on(&h->envs[e].core, h->rom, h->rom_len);
        pg_wrapper_init(&h->envs[e].wrap);
        h->envs[e].initial_template = -1;
    }
    h->threads = 0;
    h->pool = NULL;
    *out = h;
    return GBENV_OK;
}

int oracle_set_threads(oracle *h, int threads) {
    if (!h) return GBENV_E_ARG;
    h->threads = threads;
    return GBENV_OK;
}

int oracle_destroy(oracle *h) {
    if (!h) return GBENV_E_ARG;
    par_pool_stop(h);
    for (int e = 0; e < h->n; e++) pg_wrapper_free(&h->envs[e].wrap);
    for (int t = 0; t < h->n_templates; t++) free(h->templates[t]);
    free(h->templates);
    free(h->template_len);
    free(h->envs);
    free(h->rom);
    free(h);
    return GBENV_OK;
}

const char *oracle_last_error(const oracle *h) { return h ? h->err : g_err; }
int oracle_num_envs(const oracle *h) { return h ? h->n : GBENV_E_ARG; }
int oracle_sync(oracle *h) { return h ? GBENV_OK : GBENV_E_ARG; }
int oracle_check(oracle *h) { return h ? GBENV_OK : GBENV_E_ARG; } /* the oracle keeps whole maps per env: nothing to exhaust */

int oracle_add_state_template(oracle *h, const uint8_t *blob, size_t len, int *id_out) {
    if (!h || !blob || !id_out) return GBENV_E_ARG;
    GbCore *tmp = (GbCore *)malloc(sizeof(GbCore));
    gb_power_on(tmp, h->rom, h->rom_len);
    int rc = gb_load_state(tmp, blob, len);
    free(tmp);
    if (rc) return fail(h, GBENV_E_STATE, "oracle_add_state_template: unsupported save-state");
    h->templates = (uint8_t **)realloc(h->templates, sizeof(uint8_t *) * (size_t)(h->n_templates + 1));
    h->template_len = (size_t *)realloc(h->template_len, sizeof(size_t) * (size_t)(h->n_templates + 1));
    h->templates[h->n_templates] = (uint8_t *)malloc(len);
    memcpy(h->templates[h->n_templates], blob, len);
    h->template_len[h->n_templates] = len;
    *id_out = h->n_templates++;
    return GBENV_OK;
}

#define FOR_ENVS(h, ids, n, e)                                                   \
    for (int _i = 0, e = 0; _i < ((ids) ? (n) : (h)->n) && ((e = (ids) ? (ids)[_i] : _i), 1); _i++)

int oracle_load_template(oracle *h, const int32_t *ids, int n, int t) {
    if (!h || t < 0 || t >= h->n_templates) return GBENV_E_ARG;
    FOR_ENVS(h, ids, n, e) {
        if (e < 0 || e >= h->n) return GBENV_E_ARG;
        gb_load_state(&h->envs[e].core, h->templates[t], h->template_len[t]);
    }
    return GBENV_OK;
}

int oracle_set_initial_template(oracle *h, const int32_t *ids, int n, int t) {
    if (!h || t < 0 || t >= h->n_templates) return GBENV_E_ARG;
    FOR_ENVS(h, ids, n, e) {
        if (e < 0 || e >= h->n) return GBENV_E_ARG;
        h->envs[e].initial_template = t;
    }
    return GBENV_OK;
}

int oracle_power_on(oracle *h, const int32_t *ids, int n) {
    if (!h) return GBENV_E_ARG;
    FOR_ENVS(h, ids, n, e) {
        if (e < 0 || e >= h->n) return GBENV_E_ARG;
        gb_power_on(&h->envs[e].core, h->rom, h->rom_len);
    }
    return GBENV_OK;
}

int oracle_save_state(oracle *h, int env, uint8_t *blob) {
    if (!h || env < 0 || env >= h->n || !blob) return GBENV_E_ARG;
    gb_save_state(&h->envs[env].core, blob);
    return GBENV_OK;
}

static int nthreads(const oracle *h) {
    if (h->threads > 0) return h->threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* Persistent worker pool: a parallel-for over envs (OpenMP is not usable with every gcc in this image).  The workers are
 * created on first use and sleep on a condition variable between jobs, so a step costs no thread start-up; envs are handed
 * out dynamically through an atomic counter. */
typedef void (*env_fn)(oracle *h, int e, void *ctx);
struct par_pool {
    pthread_t th[256];
    int n_threads;
    pthread_mutex_t mu;
    pthread_cond_t go, done;
    unsigned long generation; /* bumped per job */
    int running;              /* workers still inside the current job */
    int quit;
    oracle *h;
    env_fn fn;
    void *ctx;
    int next; /* next env to hand out (atomic) */
};

static void par_drain(par_pool *p) {
    for (;;) {
        int e = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
        if (e >= p->h->n) break;
        p->fn(p->h, e, p->ctx);
    }
}

static void *par_worker(void *arg) {
    par_pool *p = (par_pool *)arg;
    unsigned long seen = 0;
    pthread_mutex_lock(&p->mu);
    for (;;) {
        while (!p->quit && p->generation == seen) pthread_cond_wait(&p->go, &p->mu);
        if (p->quit) break;
        seen = p->generation;
        pthread_mutex_unlock(&p->mu);
        par_drain(p);
        pthread_mutex_lock(&p->mu);
        if (--p->running == 0) pthread_cond_signal(&p->done);
    }
    pthread_mutex_unlock(&p->mu);
    return NULL;
}

static void par_pool_stop(oracle *h) {
    par_pool *p = h->pool;
    if (!p) return;
    pthread_mutex_lock(&p->mu);
    p->quit = 1;
    pthread_cond_broadcast(&p->go);
    pthread_mutex_unlock(&p->mu);
    for (int i = 0; i < p->n_threads; i++) pthread_join(p->th[i], NULL);
    pthread_mutex_destroy(&p->mu);
    pthread_cond_destroy(&p->go);
    pthread_cond_destroy(&p->done);
    free(p);
    h->pool = NULL;
}

static void par_for(oracle *h, env_fn fn, void *ctx) {
    int t = nthreads(h);
    if (t > h->n) t = h->n;
    if (t > 256) t = 256;
    if (t <= 1) {
        for (int e = 0; e < h->n; e++) fn(h, e, ctx);
        return;
    }
    if (h->pool && h->pool->n_threads != t - 1) par_pool_stop(h); /* oracle_set_threads changed the width */
    if (!h->pool) {
        par_pool *p = (par_pool *)calloc(1, sizeof(par_pool));
        pthread_mutex_init(&p->mu, NULL);
        pthread_cond_init(&p->go, NULL);
        pthread_cond_init(&p->done, NULL);
        p->h = h;
        p->n_threads = t - 1; /* the calling thread works too */
        for (int i = 0; i < p->n_threads; i++) pthread_create(&p->th[i], NULL, par_worker, p);
        h->pool = p;
    }
    par_pool *p = h->pool;
    pthread_mutex_lock(&p->mu);
    p->fn = fn;
    p->ctx = ctx;
    __atomic_store_n(&p->next, 0, __ATOMIC_RELAXED);
    p->running = p->n_threads;
    p->generation++;
    pthread_cond_broadcast(&p->go);
    pthread_mutex_unlock(&p->mu);
    par_drain(p);
    pthread_mutex_lock(&p->mu);
    while (p->running) pthread_cond_wait(&p->done, &p->mu);
    pthread_mutex_unlock(&p->mu);
}

typedef struct run_ctx {
    const uint8_t *actions;
    int frame_skip, render;
    uint8_t *obs;
    size_t obs_stride;
    double *reward;
    uint8_t *done;
    const uint8_t *skip; /* oracle_step_masked: envs that sit this step out */
} run_ctx;

static void fn_run_action(oracle *h, int e, void *ctx) {
    run_ctx *c = (run_ctx *)ctx;
    gb_run_action(&h->envs[e].core, c->actions[e], c->frame_skip);
}
static void fn_tick(oracle *h, int e, void *ctx) {
    run_ctx *c = (run_ctx *)ctx;
    h->envs[e].core.disable_renderer = !c->render;
    for (int f = 0; f < c->frame_skip; f++) gb_tick(&h->envs[e].core);
}
static void fn_step(oracle *h, int e, void *ctx) {
    run_ctx *c = (run_ctx *)ctx;
    oracle_env *E = &h->envs[e];
    int d = 0;
    if (c->skip && c->skip[e]) { /* not stepped: reward 0, done 0, observation row untouched */
        c->reward[e] = 0.0;
        c->done[e] = 0;
        return;
    }
    c->reward[e] = pg_step(&E->wrap, &E->core, c->actions[e], c->obs + (size_t)e * c->obs_stride, &d);
    c->done[e] = (uint8_t)d;
}

int oracle_run_action(oracle *h, const uint8_t *actions, int frame_skip, void *stream) {
    (void)stream;
    if (!h || !actions) return GBENV_E_ARG;
    run_ctx c = {actions, frame_skip, 0, NULL, 0, NULL, NULL, NULL};
    par_for(h, fn_run_action, &c);
    return GBENV_OK;
}

int oracle_tick(oracle *h, int n_frames, int render, void *stream) {
    (void)stream;
    if (!h) return GBENV_E_ARG;
    run_ctx c = {NULL, n_frames, render, NULL, 0, NULL, NULL, NULL};
    par_for(h, fn_tick, &c);
    return GBENV_OK;
}

int oracle_send_input(oracle *h, int button, int pressed, void *stream) {
    (void)stream;
    if (!h || button < 0 || button > 7) return GBENV_E_ARG;
    for (int e = 0; e < h->n; e++) gb_button(&h->envs[e].core, button, pressed);
    return GBENV_OK;
}

int oracle_read_mem(oracle *h, int env, uint32_t addr, uint32_t n, uint8_t *out) {
    if (!h || env < 0 || env >= h->n || addr + n > 0x10000) return GBENV_E_ARG;
    for (uint32_t i = 0; i < n; i++) out[i] = gb_read(&h->envs[env].core, (uint16_t)(addr + i));
    return GBENV_OK;
}

int oracle_write_mem(oracle *h, int env, uint32_t addr, uint32_t n, const uint8_t *in) {
    if (!h || env < 0 || env >= h->n || addr + n > 0x10000) return GBENV_E_ARG;
    for (uint32_t i = 0; i < n; i++) gb_write(&h->envs[env].core, (uint16_t)(addr + i), in[i]);
    return GBENV_OK;
}

int oracle_screen(oracle *h, int env, uint8_t *rgb) {
    if (!h || env < 0 || env >= h->n) return GBENV_E_ARG;
    const GbCore *g = &h->envs[env].core;
    for (int i = 0; i < 144 * 160; i++) {
        uint32_t px = g->screen[i];
        rgb[3 * i + 0] = (uint8_t)(px >> 8);
        rgb[3 * i + 1] = (uint8_t)(px >> 16);
        rgb[3 * i + 2] = (uint8_t)(px >> 24);
    }
    return GBENV_OK;
}

int oracle_reset(oracle *h, const uint8_t *mask, int max_episode_steps, double reward_scale, uint8_t *obs, size_t obs_stride,
                 void *stream) {
    (void)stream;
    if (!h || !obs || obs_stride < GBENV_OBS_BYTES) return GBENV_E_ARG;
    for (int e = 0; e < h->n; e++) {
        if (mask && !mask[e]) continue;
        oracle_env *E = &h->envs[e];
        const uint8_t *blob = E->initial_template >= 0 ? h->templates[E->initial_template] : NULL;
        size_t len = E->initial_template >= 0 ? h->template_len[E->initial_template] : 0;
        pg_reset(&E->wrap, &E->core, blob, len, max_episode_steps, reward_scale, obs + (size_t)e * obs_stride);
    }
    return GBENV_OK;
}

int oracle_step_masked(oracle *h, const uint8_t *actions, const uint8_t *skip, uint8_t *obs, size_t obs_stride, double *reward, uint8_t *done,
                       void *stream) {
    (void)stream;
    if (!h || !actions || !obs || !reward || !done || obs_stride < GBENV_OBS_BYTES) return GBENV_E_ARG;
    run_ctx c = {actions, 24, 0, obs, obs_stride, reward, done, skip};
    par_for(h, fn_step, &c);
    return GBENV_OK;
}

int oracle_step(oracle *h, const uint8_t *actions, uint8_t *obs, size_t obs_stride, double *reward, uint8_t *done, void *stream) {
    return oracle_step_masked(h, actions, NULL, obs, obs_stride, reward, done, stream);
}

int oracle_step_host(oracle *h, const uint8_t *actions, uint8_t *obs, double *reward, uint8_t *done) {
    return oracle_step(h, actions, obs, GBENV_OBS_BYTES, reward, done, NULL);
}

/* host memory is the oracle's only memory: the "device mask" form and the two-call step are the plain calls */
int oracle_reset_dev(oracle *h, const uint8_t *mask, int max_episode_steps, double reward_scale, uint8_t *obs, size_t obs_stride, void *stream) {
    return oracle_reset(h, mask, max_episode_steps, reward_scale, obs, obs_stride, stream);
}

int oracle_submit_host(oracle *h, const uint8_t *actions, uint8_t *obs, double *reward, uint8_t *done) {
    return oracle_step(h, actions, obs, GBENV_OBS_BYTES, reward, done, NULL);
}

int oracle_fetch_host(oracle *h) { return h ? GBENV_OK : GBENV_E_ARG; }

int oracle_reset_host(oracle *h, const uint8_t *mask, int max_episode_steps, double reward_scale, uint8_t *obs) {
    return oracle_reset(h, mask, max_episode_steps, reward_scale, obs, GBENV_OBS_BYTES, NULL);
}

int oracle_get_info(oracle *h, double *info, void *stream) {
    (void)stream;
    if (!h || !info) return GBENV_E_ARG;
    for (int e = 0; e < h->n; e++) pg_info(&h->envs[e].wrap, &h->envs[e].core, info + (size_t)e * GBENV_INFO_SCALARS);
    return GBENV_OK;
}

int oracle_reduce_info(oracle *h, double *sum, void *stream) {
    (void)stream;
    if (!h || !sum) return GBENV_E_ARG;
    double row[GBENV_INFO_SCALARS];
    memset(sum, 0, sizeof(double) * GBENV_INFO_SCALARS);
    for (int e = 0; e < h->n; e++) {
        pg_info(&h->envs[e].wrap, &h->envs[e].core, row);
        for (int k = 0; k < GBENV_INFO_SCALARS; k++) sum[k] += row[k];
    }
    return GBENV_OK;
}

int oracle_counts_map(oracle *h, int env, int32_t *map) {
    if (!h || env < 0 || env >= h->n || !map) return GBENV_E_ARG;
    pg_counts_map(&h->envs[env].wrap, map);
    return GBENV_OK;
}

int oracle_get_counters(oracle *h, gbenv_counters_t *out) {
    if (!h || !out) return GBENV_E_ARG;
    memset(out, 0, sizeof(*out));
    for (int e = 0; e < h->n; e++) {
        out->instructions += h->envs[e].core.n_instr;
        out->cycles += h->envs[e].core.n_cycles;
        out->faults += h->envs[e].core.fault;
    }
    out->frames = out->cycles / GB_FRAME_CYCLES;
    return GBENV_OK;
}

int oracle_last_kernel_ms(oracle *h, int which, float *ms) {
    (void)h;
    (void)which;
    if (ms) *ms = 0.0f;
    return GBENV_OK;
}

int oracle_get_lanes_per_warp(const oracle *h) { return h ? 1 : GBENV_E_ARG; }
int oracle_set_lanes_per_warp(oracle *h, int lanes) {
    (void)lanes;
    return h ? GBENV_OK : GBENV_E_ARG;
}

int oracle_kernel_time_total(oracle *h, int which, double *ms, uint64_t *steps) {
    (void)h;
    (void)which;
    if (ms) *ms = 0.0;
    if (steps) *steps = 0;
    return GBENV_OK;
}

int oracle_get_core_extra(oracle *h, int env, gbenv_core_extra_t *out) {
    if (!h || env < 0 || env >= h->n || !out) return GBENV_E_ARG;
    out->stat_mode = h->envs[env].core.stat_mode;
    out->ly_window = h->envs[env].core.ly_window;
    out->fault = h->envs[env].core.fault;
    out->reserved = 0;
    return GBENV_OK;
}

/* wrapper-state digest used by parity tests (see pokegym_wrapper.h) */
int oracle_wrapper_digest(oracle *h, int env, double *out, int n) {
    if (!h || env < 0 || env >= h->n || !out) return GBENV_E_ARG;
    return pg_digest(&h->envs[env].wrap, out, n);
}

/* Render one scanline of a freshly loaded state (PPU known-answer tests) */
int oracle_debug_render_frame(oracle *h, int env) {
    if (!h || env < 0 || env >= h->n) return GBENV_E_ARG;
    GbCore *g = &h->envs[env].core;
    uint8_t scx = g->SCX, scy = g->SCY, wx = g->WX, wy = g->WY, lcdc = g->LCDC;
    g->disable_renderer = 0;
    g->ly_window = -1;
    for (int y = 0; y < 144; y++) {
        g->SCX = g->scanline_params[y][0];
        g->SCY = g->scanline_params[y][1];
        g->WX = (uint8_t)(g->scanline_params[y][2] + 7);
        g->WY = g->scanline_params[y][3];
        g->LCDC = (uint8_t)((lcdc & ~0x10) | (g->scanline_params[y][4] ? 0x10 : 0));
        gb_render_scanline(g, y);
    }
    g->SCX = scx; g->SCY = scy; g->WX = wx; g->WY = wy; g->LCDC = lcdc;
    return GBENV_OK;
}
