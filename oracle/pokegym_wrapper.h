/*
 * oracle/pokegym_wrapper.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's Python wrapper around the emulator:
 *   /root/reference/pokegym/environment.py  reset :1233-1334, step :1336-1812, render :256-274
 *   /root/reference/pokegym/ram_map.py :1441-1898, ram_map_leanke.py :793-1098, game_map.py :11-18
 *   /root/reference/pokegym/bin/ram_reader/red_ram_api.py :59-73,149-225,404-422,542-602,816-820
 * PINNED: tests/test_wrapper_vs_reference.py runs the UNMODIFIED reference Environment on a PyBoy
 * shim over gb_core.c and compares reward / obs / RAM side effects with this file step by step;
 * the resulting vectors are committed under tests/golden/ for the GPU tests.
 */
#ifndef PG_WRAPPER_H
#define PG_WRAPPER_H

#include <stddef.h>
#include <stdint.h>

#include "gb_core.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PG_MAPS 248
#define PG_CUT_COORDS_MAX 512
#define PG_INFO_SCALARS 72

typedef struct PgCutCoord {
    int32_t x, y, map;
    double value;
} PgCutCoord;

typedef struct PgWrapper {
    /* env-lifetime state (never cleared by reset) */
    int reset_count;
    int is_dead;
    int last_map; /* -1 initially */
    double item_reward[5]; /* lemonade, silph scope, lift key, poke doll, bicycle */
    int32_t *counts_map;   /* 444 x 436 */
    /* per-episode state */
    int time, max_episode_steps;
    double reward_scale;
    int have_last_reward;
    double last_reward;
    int max_events, max_level_sum, max_opponent_level;
    uint8_t *seen_coords;   /* bitmap [248][256][256 bits] */
    uint8_t *screen_memory; /* bitmap [248][256][256 bits] */
    int n_seen_coords;
    uint8_t seen_maps[PG_MAPS];
    int n_seen_maps;
    int prev_map_n; /* -2 = None */
    int death_count;
    double total_healing, last_hp;
    int last_party_size;
    int hm_count_latch, cut, used_cut, used_cut_on_map_n;
    PgCutCoord cut_coords[PG_CUT_COORDS_MAX];
    int n_cut_coords;
    uint8_t cut_tiles[256];
    int n_cut_tiles;
    int32_t cut_state[3][6];
    int n_cut_state;
    int seen_start_menu, seen_pokemon_menu, seen_stats_menu, seen_bag_menu;
    uint8_t seen_pokemon[152], caught_pokemon[152], moves_obtained[0xA5];
    int last_map_id_plus1; /* last_10_map_ids[0][0]; 0 after reset */
    /* last-step scalars kept for the info dict */
    double info[PG_INFO_SCALARS];
    int overflow; /* 1 if a bounded container overflowed (cut_coords) */
} PgWrapper;

void pg_wrapper_init(PgWrapper *w);
void pg_wrapper_free(PgWrapper *w);
/* Environment.reset: state blob may be NULL (no initial state registered). obs: 72*80*4 bytes. */
void pg_reset(PgWrapper *w, GbCore *g, const uint8_t *state_blob, size_t state_len, int max_episode_steps, double reward_scale,
              uint8_t *obs);
/* Environment.step: returns reward, writes obs and *done. */
double pg_step(PgWrapper *w, GbCore *g, int action, uint8_t *obs, int *done);
/* wrapper half of step only (everything after run_action_on_emulator) */
double pg_after_emulation(PgWrapper *w, GbCore *g, int action, uint8_t *obs, int *done);
void pg_info(const PgWrapper *w, const GbCore *g, double *out /* PG_INFO_SCALARS */);
void pg_counts_map(const PgWrapper *w, int32_t *out /* 444*436 */);
/* flat digest of the wrapper state for parity tests; returns number of doubles written */
int pg_digest(const PgWrapper *w, double *out, int n);

#ifdef __cplusplus
}
#endif
#endif
