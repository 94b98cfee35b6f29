/*
 * oracle/pokegym_wrapper.c -- TEST INFRASTRUCTURE (see pokegym_wrapper.h).
 * Straight restatement of the reference wrapper, one env, plain C, float64 where Python uses float.
 * Citations are /root/reference/pokegym/<file>:<line>.
 */
#include "pokegym_wrapper.h"

#include <stdlib.h>
#include <string.h>

#include "../include/gbenv_info.h"

static const struct { int x, y, known; } MAP_OFFSETS[PG_MAPS] = {
#include "map_offsets.inc"
};

/* environment.py:58-78 TREE_POSITIONS_PIXELS (first, second, map) */
static const int TREES[19][3] = {
    {3184, 3584, 6},  {3375, 3391, 6},  {2528, 3616, 134}, {2480, 3568, 134}, {2560, 3584, 134}, {1104, 2944, 13}, {1264, 3136, 13},
    {1216, 3616, 13}, {1216, 3744, 13}, {1216, 3872, 13},  {1088, 4000, 1},   {992, 4288, 1},    {3984, 4512, 5},  {4640, 1392, 36},
    {4464, 2176, 20}, {5488, 2336, 21}, {5488, 2368, 21},  {5488, 2400, 21},  {5488, 2432, 21}};

/* red_memory_menus.py:237-299 keys of TEXT_MENU_CURSOR_LOCATIONS, as (CC30, CC31) */
static const uint8_t CURSOR_KEYS[][2] = {
    {0xD3, 0xC3}, {0xFB, 0xC3}, {0x23, 0xC4}, {0x4B, 0xC4}, {0x73, 0xC4}, {0x9B, 0xC4}, {0xC3, 0xC4}, /* start menu */
    {0x4C, 0xC4}, {0x74, 0xC4},                                                                     /* pokecenter */
    {0xB5, 0xC3}, {0xDD, 0xC3}, {0x05, 0xC4},                                                       /* pokemart */
    {0xC9, 0xC3}, {0xF1, 0xC3}, {0x19, 0xC4}, {0x41, 0xC4},                                         /* pc */
    {0x9A, 0xC4}, {0xC2, 0xC4}, {0xEA, 0xC4},                                                       /* pc someone */
    {0xC1, 0xC4}, {0xA9, 0xC4}, {0xBD, 0xC4}, {0xD1, 0xC4}, {0xE5, 0xC4},                           /* battle fight/moves */
    {0xC7, 0xC4}, {0xB4, 0xC3}, {0xDC, 0xC3}, {0x04, 0xC4}, {0x2C, 0xC4}, {0x54, 0xC4}, {0x7C, 0xC4}, /* roster */
    {0x9C, 0xC4}, {0xC4, 0xC4}, {0xEC, 0xC4},                                                       /* sub select / cancel */
    {0xE9, 0xC4}, {0x8A, 0xC4}, {0xB2, 0xC4},                                                       /* battle item */
    {0xF5, 0xC3}, {0x1D, 0xC4}, {0x45, 0xC4},                                                       /* mart/pc item 1,2,n */
    {0x69, 0x01}, {0xEF, 0xC4},                                                                     /* item cancel, run */
    {0x4F, 0xC4}, {0x77, 0xC4}, {0x69, 0xC4},                                                       /* yes, no, no-hack */
    {0x59, 0xC4}, {0x6D, 0xC4}, {0x81, 0xC4}};                                                      /* overwrite move 2-4 */

/* ram_map_leanke.py monitor_* tables as (address, bit, weight), in dict order */
typedef struct { uint16_t addr; uint8_t bit; int8_t w; } EvBit;
#define T 1
#define Q 5
static const EvBit SILPH[] = { /* :107-164 */
    {0xD825, 2, T}, {0xD825, 3, T}, {0xD825, 4, T}, {0xD825, 5, T}, {0xD826, 5, Q}, {0xD826, 6, Q}, {0xD827, 2, T}, {0xD827, 3, T},
    {0xD828, 0, Q}, {0xD828, 1, Q}, {0xD829, 2, T}, {0xD829, 3, T}, {0xD829, 4, T}, {0xD82A, 0, Q}, {0xD82A, 1, Q}, {0xD82B, 2, T},
    {0xD82B, 3, T}, {0xD82B, 4, T}, {0xD82B, 5, T}, {0xD82C, 0, Q}, {0xD82C, 1, Q}, {0xD82C, 2, Q}, {0xD82D, 6, T}, {0xD82D, 7, T},
    {0xD82E, 0, T}, {0xD82E, 7, Q}, {0xD82F, 5, T}, {0xD82F, 6, T}, {0xD82F, 7, T}, {0xD830, 0, T}, {0xD830, 4, Q}, {0xD830, 5, Q},
    {0xD830, 6, Q}, {0xD831, 2, T}, {0xD831, 3, T}, {0xD831, 4, T}, {0xD832, 0, Q}, {0xD833, 2, T}, {0xD833, 3, T}, {0xD833, 4, T},
    {0xD834, 0, Q}, {0xD834, 1, Q}, {0xD834, 2, Q}, {0xD834, 3, Q}, {0xD835, 1, T}, {0xD835, 2, T}, {0xD836, 0, Q}, {0xD837, 4, T},
    {0xD837, 5, T}, {0xD838, 0, Q}, {0xD838, 5, 5 /* ITEM */}, {0xD838, 7, 5 /* GYM_LEADER */}, {0xD7B9, 7, 2 /* TASK */}};
static const EvBit DOJO[] = { /* :816-827 */
    {0xD7B1, 0, -1}, {0xD7B1, 1, 5}, {0xD7B1, 2, 1}, {0xD7B1, 3, 1}, {0xD7B1, 4, 1}, {0xD7B1, 5, 1}, {0xD7B1, 6, 3}, {0xD7B1, 7, 3}};
static const EvBit HIDEOUT[] = { /* :867-889, every weight overridden to 1 */
    {0xD815, 1, 1}, {0xD815, 2, 1}, {0xD815, 3, 1}, {0xD815, 4, 1}, {0xD815, 5, 1}, {0xD817, 1, 1}, {0xD819, 1, 1}, {0xD819, 2, 1},
    {0xD81B, 2, 1}, {0xD81B, 3, 1}, {0xD81B, 4, 1}, {0xD81B, 5, 1}, {0xD81B, 6, 1}, {0xD81B, 7, 1}, {0xD77E, 1, 1}};
static const EvBit TOWER[] = { /* :936-957 */
    {0xD765, 1, T}, {0xD765, 2, T}, {0xD765, 3, T}, {0xD766, 1, T}, {0xD766, 2, T}, {0xD766, 3, T}, {0xD767, 2, T}, {0xD767, 3, T}, {0xD767, 4, T},
    {0xD767, 5, T}, {0xD768, 1, T}, {0xD768, 2, T}, {0xD768, 3, T}, {0xD768, 7, Q}, {0xD769, 1, T}, {0xD769, 2, T}, {0xD769, 3, T}};
static const EvBit GYM3[] = {{0xD773, 1, 2}, {0xD773, 0, 2}, {0xD773, 7, 5}, {0xD773, 2, 2}, {0xD773, 3, 2}, {0xD773, 4, 2}}; /* :1038-1047 */
static const EvBit GYM4[] = {{0xD792, 1, 5}, {0xD77C, 2, 2}, {0xD77C, 3, 2}, {0xD77C, 4, 2}, {0xD77C, 5, 2}, {0xD77C, 6, 2}, {0xD77C, 7, 2}, {0xD77D, 0, 2}};
static const EvBit GYM5[] = {{0xD7B3, 1, 5}, {0xD792, 2, 2}, {0xD792, 3, 2}, {0xD792, 4, 2}, {0xD792, 5, 2}, {0xD792, 6, 2}, {0xD792, 7, 2}};
static const EvBit GYM6[] = {{0xD7B3, 1, 5}, {0xD7B3, 2, 2}, {0xD7B3, 3, 2}, {0xD7B3, 4, 2}, {0xD7B3, 5, 2}, {0xD7B3, 6, 2}, {0xD7B3, 7, 2}, {0xD7B4, 0, 2}};
static const EvBit GYM7[] = {{0xD79A, 1, 5}, {0xD79A, 2, 2}, {0xD79A, 3, 2}, {0xD79A, 4, 2}, {0xD79A, 5, 2}, {0xD79A, 6, 2}, {0xD79A, 7, 2}, {0xD79B, 0, 2}};
#undef T
#undef Q
#define COUNT(a) ((int)(sizeof(a) / sizeof(a[0])))

static int M(GbCore *g, int addr) { return gb_read(g, (uint16_t)addr); }
static int BITOF(GbCore *g, int addr, int bit) { return (M(g, addr) >> bit) & 1; }
static int popcount8(int v) {
    int c = 0;
    for (; v; v &= v - 1) c++;
    return c;
}

/* environment.py:1201-1219 calculate_event_rewards(base 10, increment 2, multiplier 1) */
static int event_rewards(GbCore *g, const EvBit *t, int n) {
    int total = 0, cur = 10;
    for (int i = 0; i < n; i++) {
        int points = t[i].w * BITOF(g, t[i].addr, t[i].bit);
        if (points > 0) {
            total += cur * points;
            cur += 2;
        }
    }
    return total;
}

void pg_wrapper_init(PgWrapper *w) {
    memset(w, 0, sizeof(*w));
    w->counts_map = (int32_t *)calloc(444 * 436, sizeof(int32_t));
    w->seen_coords = (uint8_t *)calloc((size_t)PG_MAPS * 256 * 32, 1);
    w->screen_memory = (uint8_t *)calloc((size_t)PG_MAPS * 256 * 32, 1);
    w->last_map = -1;
    w->prev_map_n = -2;
    w->last_hp = 1.0;
    w->last_party_size = 1;
    w->reward_scale = 1.0;
    w->max_episode_steps = 20480;
}

void pg_wrapper_free(PgWrapper *w) {
    free(w->counts_map);
    free(w->seen_coords);
    free(w->screen_memory);
    w->counts_map = NULL;
    w->seen_coords = w->screen_memory = NULL;
}

static size_t bit_index(int map, int r, int c) { return ((size_t)map * 256 + (size_t)r) * 256 + (size_t)c; }
static int bm_get(const uint8_t *bm, int map, int r, int c) {
    size_t i = bit_index(map, r, c);
    return (bm[i >> 3] >> (i & 7)) & 1;
}
static void bm_set(uint8_t *bm, int map, int r, int c) {
    size_t i = bit_index(map, r, c);
    bm[i >> 3] |= (uint8_t)(1u << (i & 7));
}

/* ram_map.position :1522-1538 (bytes, so only the map clamp can fire) */
static void position(GbCore *g, int *r, int *c, int *map_n) {
    *r = M(g, 0xD361);
    *c = M(g, 0xD362);
    *map_n = M(g, 0xD35E);
    if (*map_n > 247) *map_n = 247;
}

/* environment.py:256-274 render + :233-254 get_fixed_window */
static void render(PgWrapper *w, GbCore *g, uint8_t *obs) {
    int r, c, map_n;
    position(g, &r, &c, &map_n);
    if (r <= 254 && c <= 254) bm_set(w->screen_memory, map_n, r, c);
    for (int i = 0; i < 72; i++)
        for (int j = 0; j < 80; j++) {
            uint32_t px = g->screen[(2 * i) * 160 + 2 * j];
            uint8_t *o = obs + (i * 80 + j) * 4;
            o[0] = (uint8_t)(px >> 8);
            o[1] = (uint8_t)(px >> 16);
            o[2] = (uint8_t)(px >> 24);
            int rr = r - 36 + i, cc = c - 40 + j;
            o[3] = (rr >= 0 && rr < 255 && cc >= 0 && cc < 255 && bm_get(w->screen_memory, map_n, rr, cc)) ? 255 : 0;
        }
}

/* environment.py:1014-1025 */
static void minor_patch_victory_road(GbCore *g) {
    static const int ab[5][2] = {{0xD7EE, 0}, {0xD7EE, 7}, {0xD813, 0}, {0xD813, 6}, {0xD869, 7}};
    for (int i = 0; i < 5; i++) gb_write(g, (uint16_t)ab[i][0], (uint8_t)(M(g, ab[i][0]) | (1 << ab[i][1])));
}

/* environment.py:1027-1052 (observable part: the RAM patch on a map change) */
static void update_last_10_map_ids(PgWrapper *w, GbCore *g) {
    int cur = M(g, 0xD35E) + 1;
    if (cur == w->last_map_id_plus1) return;
    w->last_map_id_plus1 = cur;
    int map_id = cur - 1;
    if (map_id == 0x6C || map_id == 0xC2 || map_id == 0xC6 || map_id == 0x22) minor_patch_victory_road(g);
}

void pg_reset(PgWrapper *w, GbCore *g, const uint8_t *blob, size_t len, int max_episode_steps, double reward_scale, uint8_t *obs) {
    /* :1236-1239  all_events_string is read first, then get_base_event_flags writes D778 |= 0x10 */
    gb_write(g, 0xD778, (uint8_t)(M(g, 0xD778) | 0x10));
    /* :1241-1242  the state is only loaded on the first reset */
    if (w->reset_count == 0 && blob) gb_load_state(g, blob, len);
    /* :1251-1331 */
    memset(w->screen_memory, 0, (size_t)PG_MAPS * 256 * 32);
    w->reset_count += 1;
    w->time = 0;
    w->max_episode_steps = max_episode_steps;
    w->reward_scale = reward_scale;
    w->have_last_reward = 0;
    w->last_reward = 0.0;
    w->prev_map_n = -2;
    w->max_events = 0;
    w->max_level_sum = 0;
    w->max_opponent_level = 0;
    memset(w->seen_coords, 0, (size_t)PG_MAPS * 256 * 32);
    w->n_seen_coords = 0;
    memset(w->seen_maps, 0, sizeof(w->seen_maps));
    w->n_seen_maps = 0;
    w->death_count = 0;
    w->total_healing = 0.0;
    w->last_hp = 1.0;
    w->last_party_size = 1;
    w->hm_count_latch = 0;
    w->cut = 0;
    w->used_cut = 0;
    w->n_cut_coords = 0;
    memset(w->cut_tiles, 0, sizeof(w->cut_tiles));
    w->n_cut_tiles = 0;
    w->n_cut_state = 0;
    w->seen_start_menu = w->seen_pokemon_menu = w->seen_stats_menu = w->seen_bag_menu = 0;
    memset(w->seen_pokemon, 0, sizeof(w->seen_pokemon));
    memset(w->caught_pokemon, 0, sizeof(w->caught_pokemon));
    memset(w->moves_obtained, 0, sizeof(w->moves_obtained));
    w->last_map_id_plus1 = 0;
    memset(w->info, 0, sizeof(w->info));
    update_last_10_map_ids(w, g); /* :1327 */
    render(w, g, obs);            /* :1334 */
}

/* red_ram_api.py:59-73 process_game_states: only its RAM side effect (:596-600) is observable */
static void process_game_states(GbCore *g) {
    if (M(g, 0xCFC4) != 0) return; /* pre-battle needs text; get_menu_state writes only when no text */
    int battle_type = M(g, 0xD057); /* 255 -> DIED(4): still truthy */
    int pre_battle = M(g, 0xD059);
    if (battle_type || pre_battle) {
        /* Battle._get_battle_menu_state :176-201 returns GAME_STATE_UNKNOWN only on this path */
        int c0 = M(g, 0xCC30), c1 = M(g, 0xCC31);
        for (int i = 0; i < COUNT(CURSOR_KEYS); i++)
            if (CURSOR_KEYS[i][0] == c0 && CURSOR_KEYS[i][1] == c1) return; /* a known menu state */
        if ((c0 == 0 && c1 == 0) || !battle_type) return;                      /* BATTLE_ANIMATION */
        if ((M(g, 0xD125) == 0x01 && M(g, 0xD730) != 0x40) || M(g, 0xCC52) == 0x00) return; /* BATTLE_TEXT */
    }
    if (M(g, 0xCD38) != 0) return; /* FOLLOWING_NPC :816-820 */
    gb_write(g, 0xCC30, 0);
    gb_write(g, 0xCC31, 0);
    for (int i = 0; i < 10; i++) gb_write(g, (uint16_t)(0xCF7C + i), 0);
}

static void local_to_global(int r, int c, int map_n, int *gr, int *gc) {
    /* game_map.py:11-18 */
    *gr = r + MAP_OFFSETS[map_n].y;
    *gc = c + MAP_OFFSETS[map_n].x;
}

/* environment.py:277-312 */
static double detect_and_reward_trees(int player_x, int player_y, int map_n) {
    double total = 0.0;
    for (int i = 0; i < 19; i++) {
        if (TREES[i][2] != map_n) continue;
        int tree_x = TREES[i][1] / 16, tree_y = TREES[i][0] / 16;
        int cy = (tree_x == 212 && tree_y == 210) ? 211 : tree_y;
        int d = abs(player_x - tree_x) + abs(player_y - cy);
        if (d <= 5) total += 1.0 / (double)(d > 1 ? d : 1);
    }
    return total;
}

static const int32_t CUT_SEQ[2][2][6] = {{{0x3D, 1, 1, 0, 4, 1}, {0x3D, 1, 1, 0, 1, 1}}, {{0x50, 1, 1, 0, 4, 1}, {0x50, 1, 1, 0, 1, 1}}};
static const int32_t CUT_GRASS_SEQ[3][6] = {{0x52, 255, 1, 0, 1, 1}, {0x52, 255, 1, 0, 1, 1}, {0x52, 1, 1, 0, 1, 1}};
static const int32_t CUT_FAIL_SEQ[3][6] = {{-1, 255, 0, 0, 4, 1}, {-1, 255, 0, 0, 1, 1}, {-1, 255, 0, 0, 1, 1}};

static void cut_coords_set(PgWrapper *w, int x, int y, int map, double v) {
    for (int i = 0; i < w->n_cut_coords; i++)
        if (w->cut_coords[i].x == x && w->cut_coords[i].y == y && w->cut_coords[i].map == map) {
            w->cut_coords[i].value = v;
            return;
        }
    if (w->n_cut_coords == PG_CUT_COORDS_MAX) {
        w->overflow = 1;
        return;
    }
    PgCutCoord *e = &w->cut_coords[w->n_cut_coords++];
    e->x = x; e->y = y; e->map = map; e->value = v;
}

double pg_after_emulation(PgWrapper *w, GbCore *g, int action, uint8_t *obs, int *done) {
    (void)action;
    w->time += 1; /* :1338 */
    int r, c, map_n;
    position(g, &r, &c, &map_n); /* :1344-1345 */
    if (!bm_get(w->seen_coords, map_n, r, c)) {
        bm_set(w->seen_coords, map_n, r, c);
        w->n_seen_coords++;
    }
    process_game_states(g); /* :1348 */
    /* :1349,1358-1372  20 bag slots, no early stop */
    static const int ITEM_IDS[5] = {0x3E, 0x48, 0x4A, 0x33, 0x06};
    for (int i = 0; i < 20; i++) {
        int id = M(g, 0xD31E + 2 * i);
        for (int k = 0; k < 5; k++)
            if (id == ITEM_IDS[k]) w->item_reward[k] = 20.0;
    }
    update_last_10_map_ids(w, g); /* :1352 */
    /* :1375 (used_cut as of before this step's update) */
    double exploration_reward = (w->used_cut < 1 ? 0.02 : 0.1) * (double)w->n_seen_coords;
    /* :1377 update_heat_map :648-679 */
    int glob_r, glob_c;
    local_to_global(r, c, map_n, &glob_r, &glob_c);
    if (glob_r < 444 && glob_c < 436) {
        if (w->last_map == map_n || w->last_map == -1)
            w->counts_map[glob_r * 436 + glob_c] += 1;
        else
            w->counts_map[glob_r * 436 + glob_c] = -1;
    }
    w->last_map = map_n;
    if (map_n != w->prev_map_n) { /* :1378-1382 */
        w->used_cut_on_map_n = 0;
        w->prev_map_n = map_n;
        if (!w->seen_maps[map_n]) {
            w->seen_maps[map_n] = 1;
            w->n_seen_maps++;
        }
    }
    /* :1386-1391 level reward */
    int party_size = M(g, 0xD163), level_sum = 0, max_level = 0;
    for (int k = 0; k < 6; k++) {
        int lv = M(g, 0xD18C + 44 * k);
        level_sum += lv;
        if (lv > max_level) max_level = lv;
    }
    if (level_sum > w->max_level_sum) w->max_level_sum = level_sum;
    double level_reward = w->max_level_sum < 50 ? (double)w->max_level_sum : 50.0 + (double)(w->max_level_sum - 50) / 4.0;
    /* :1394-1408 healing / death */
    int hp_sum = 0, max_hp_sum = 0;
    for (int k = 0; k < 6; k++) {
        hp_sum += 256 * M(g, 0xD16C + 44 * k) + M(g, 0xD16D + 44 * k);
        max_hp_sum += 256 * M(g, 0xD18D + 44 * k) + M(g, 0xD18E + 44 * k);
    }
    double hp = max_hp_sum == 0 ? 1.0 : (double)hp_sum / (double)max_hp_sum;
    double hp_delta = hp - w->last_hp;
    if (hp_delta > 0.2 && party_size == w->last_party_size && !w->is_dead) w->total_healing += hp_delta;
    if (hp <= 0 && w->last_hp > 0) {
        w->death_count += 1;
        w->is_dead = 1;
    } else if (hp > 0.01) {
        w->is_dead = 0;
    }
    w->last_hp = hp;
    w->last_party_size = party_size;
    /* :1411-1426 */
    int badges = popcount8(M(g, 0xD356));
    int bill_state = BITOF(g, 0xD7F2, 3);
    int hm_mask = 0;
    for (int i = 0; i < 10; i++) { /* ram_map.get_items_in_bag :1867-1875 */
        int id = M(g, 0xD31E + 2 * i);
        if (id == 0 || id == 0xFF) break;
        if (id >= 0xC4 && id <= 0xC8) hm_mask |= 1 << (id - 0xC4);
    }
    int hm_count = popcount8(hm_mask);
    if (hm_count >= 1 && w->hm_count_latch == 0) w->hm_count_latch = 1;
    int cut_rew = w->cut * 8;
    double tree_distance_reward = detect_and_reward_trees(glob_r, glob_c, map_n); /* :1429-1431 */
    /* :1435-1440 money / opponent level (info only) */
    int money = 0;
    {
        int b0 = M(g, 0xD347), b1 = M(g, 0xD348), b2 = M(g, 0xD349);
        money = 10000 * (10 * (b0 >> 4) + (b0 & 15)) + 100 * (10 * (b1 >> 4) + (b1 & 15)) + (10 * (b2 >> 4) + (b2 & 15));
    }
    int max_opp = 0;
    for (int k = 0; k < 6; k++) {
        int lv = M(g, 0xD8C5 + 44 * k);
        if (lv > max_opp) max_opp = lv;
    }
    if (max_opp > w->max_opponent_level) w->max_opponent_level = max_opp;
    /* :1443-1445 events */
    int num_events = 0;
    for (int a = 0xD747; a < 0xD886; a++) num_events += popcount8(M(g, a));
    int events = num_events - 13 - BITOF(g, 0xD754, 0);
    if (events < 0) events = 0;
    if (events > w->max_events) w->max_events = events;
    /* :1448 dojo (ram_map_leanke.py:793-814) */
    int dojo_reward = 0;
    for (int i = 0; i < COUNT(DOJO); i++) dojo_reward += DOJO[i].w * BITOF(g, DOJO[i].addr, DOJO[i].bit);
    /* :1457-1491 */
    int silph_ev = event_rewards(g, SILPH, COUNT(SILPH));
    int dojo_ev = event_rewards(g, DOJO, COUNT(DOJO));
    int hideout_ev = event_rewards(g, HIDEOUT, COUNT(HIDEOUT));
    int tower_ev = event_rewards(g, TOWER, COUNT(TOWER));
    int g3 = event_rewards(g, GYM3, COUNT(GYM3)), g4 = event_rewards(g, GYM4, COUNT(GYM4)), g5 = event_rewards(g, GYM5, COUNT(GYM5));
    int g6 = event_rewards(g, GYM6, COUNT(GYM6)), g7 = event_rewards(g, GYM7, COUNT(GYM7));
    /* :1496-1538 cut state machine */
    if (M(g, 0xD057) == 0 && w->cut == 1) {
        int dir = M(g, 0xC109);
        int x = M(g, 0xD362), y = M(g, 0xD361), map_id = M(g, 0xD35E);
        int have_coords = 1, cx = x, cy = y;
        if (dir == 0) cy = y + 1;
        else if (dir == 4) cy = y - 1;
        else if (dir == 8) cx = x - 1;
        else if (dir == 0xC) cx = x + 1;
        else have_coords = 0; /* the reference would raise UnboundLocalError on use */
        if (w->n_cut_state == 3) {
            memmove(w->cut_state[0], w->cut_state[1], sizeof(int32_t) * 12);
            w->n_cut_state = 2;
        }
        int32_t *s = w->cut_state[w->n_cut_state++];
        s[0] = M(g, 0xCFC6); s[1] = M(g, 0xCFCB); s[2] = M(g, 0xCD6A); s[3] = M(g, 0xD367); s[4] = M(g, 0xD125); s[5] = M(g, 0xCD3D);
        int hit = 0;
        double val = 0.0;
        if (w->n_cut_state == 3) {
            for (int q = 0; q < 2 && !hit; q++)
                if (!memcmp(w->cut_state[1], CUT_SEQ[q][0], 24) && !memcmp(w->cut_state[2], CUT_SEQ[q][1], 24)) { hit = 1; val = 10.0; }
            if (!hit && !memcmp(w->cut_state, CUT_GRASS_SEQ, 72)) { hit = 1; val = 0.001; }
            if (!hit) {
                int same = 1;
                for (int i = 0; i < 3 && same; i++)
                    for (int k = 1; k < 6; k++)
                        if (w->cut_state[i][k] != CUT_FAIL_SEQ[i][k]) { same = 0; break; }
                if (same) { hit = 1; val = 0.001; }
            }
        }
        if (hit && have_coords) {
            cut_coords_set(w, cx, cy, map_id, val);
            int tile = w->cut_state[w->n_cut_state - 1][0];
            if (!w->cut_tiles[tile]) { w->cut_tiles[tile] = 1; w->n_cut_tiles++; }
        }
        if (BITOF(g, 0xD803, 0)) { /* :1527-1538, D057 == 0 already known */
            int cf13 = M(g, 0xCF13), ff8c = M(g, 0xFF8C), cf94 = M(g, 0xCF94);
            if (cf13 == 0 && ff8c == 6 && cf94 == 0) w->seen_start_menu = 1;
            if (cf13 == 0 && ff8c == 6 && cf94 == 2) w->seen_pokemon_menu = 1;
            if (cf13 == 0) w->seen_stats_menu = 1;
            if (cf13 == 0 && cf94 == 3) w->seen_bag_menu = 1;
        }
    }
    /* :1541 update_pokedex :552-558 */
    int n_seen = 0, n_caught = 0;
    for (int i = 0; i < 19; i++) {
        int cm = M(g, 0xD2F7 + i), sm = M(g, 0xD30A + i);
        for (int j = 0; j < 8; j++) {
            w->caught_pokemon[8 * i + j] = (cm >> j) & 1;
            w->seen_pokemon[8 * i + j] = (sm >> j) & 1;
        }
        n_seen += popcount8(sm);
        n_caught += popcount8(cm);
    }
    /* :1542 update_moves_obtained :560-580 */
    for (int k = 0; k < 6; k++) {
        int base = 0xD16B + 44 * k;
        if (M(g, base) != 0)
            for (int j = 0; j < 4; j++) {
                int mv = M(g, base + j + 8);
                if (mv != 0) {
                    if (mv < 0xA5) w->moves_obtained[mv] = 1; /* the reference would raise IndexError beyond */
                    if (mv == 15) w->cut = 1;
                }
            }
    }
    int box_n = M(g, 0xDA80);
    for (int i = 0; i < box_n; i++) {
        int off = i * 200 + 0xDA96;
        if (off + 11 > 0xFFFF) break; /* the reference would raise past the address space */
        if (M(g, off) != 0)
            for (int j = 0; j < 4; j++) {
                int mv = M(g, off + j + 8);
                if (mv != 0 && mv < 0xA5) w->moves_obtained[mv] = 1;
            }
    }
    int n_moves = 0;
    for (int i = 0; i < 0xA5; i++) n_moves += w->moves_obtained[i];
    /* :1544 bill_capt ram_map.py:1889-1898 */
    int bill_capt_rew = 5 * (BITOF(g, 0xD7F1, 0) + BITOF(g, 0xD7F2, 3) + BITOF(g, 0xD7F2, 4) + BITOF(g, 0xD7F2, 5) + BITOF(g, 0xD7F2, 6) +
                             BITOF(g, 0xD7F2, 7) + BITOF(g, 0xD803, 0) + BITOF(g, 0xD803, 1));
    /* :1547-1552 */
    if (M(g, 0xCD4D) == 61) {
        gb_write(g, 0xCD4D, 0);
        w->used_cut += 1;
    }
    /* :1554-1600 reward, same association order as the Python expression */
    double S = w->reward_scale;
    double start_menu = w->seen_start_menu * 0.01, pokemon_menu = w->seen_pokemon_menu * 0.1;
    double stats_menu = w->seen_stats_menu * 0.1, bag_menu = w->seen_bag_menu * 0.1;
    double cut_coords = 0.0;
    for (int i = 0; i < w->n_cut_coords; i++) cut_coords += w->cut_coords[i].value;
    cut_coords = cut_coords * 1.0;
    double cut_tiles = w->n_cut_tiles * 1.0;
    double that_guy = ((start_menu + pokemon_menu) + stats_menu) + bag_menu;
    double seen_pokemon_reward = S * (double)n_seen, caught_pokemon_reward = S * (double)n_caught, moves_obtained_reward = S * (double)n_moves;
    int bill_reward = 5 * bill_state, hm_reward = hm_count * 10, badges_reward = 10 * badges;
    double acc = (double)(w->max_events + bill_capt_rew);
    acc += seen_pokemon_reward;
    acc += caught_pokemon_reward;
    acc += moves_obtained_reward;
    acc += (double)bill_reward;
    acc += (double)hm_reward;
    acc += level_reward;
    acc += 0.0; /* death_reward */
    acc += (double)badges_reward;
    acc += w->total_healing;
    acc += exploration_reward;
    acc += (double)cut_rew;
    acc += that_guy / 2;
    acc += cut_coords;
    acc += cut_tiles;
    acc += tree_distance_reward * 0.6;
    acc += (double)(dojo_reward * 5);
    for (int k = 0; k < 5; k++) acc += w->item_reward[k];
    acc += (double)(dojo_ev + silph_ev + hideout_ev + tower_ev + g3 + g4 + g5 + g6 + g7);
    acc += (double)(g3 + g4 + g5 + g6 + g7);
    double reward_abs = S * acc;
    double reward;
    if (!w->have_last_reward) { /* :1604-1610 */
        reward = 0.0;
        w->last_reward = 0.0;
        w->have_last_reward = 1;
    } else {
        reward = reward_abs - w->last_reward;
        w->last_reward = reward_abs;
    }
    *done = w->time >= w->max_episode_steps; /* :1613 */
    /* scalars of the info dict :1621-1703 */
    double *I = w->info;
    I[GBI_COUNT] = 1; I[GBI_STEP] = w->time; I[GBI_X] = c; I[GBI_Y] = r; I[GBI_MAP] = map_n; I[GBI_PCOUNT] = party_size;
    for (int k = 0; k < 6; k++) I[GBI_LEVEL0 + k] = M(g, 0xD18C + 44 * k);
    I[GBI_LEVELS_SUM] = level_sum;
    I[GBI_DEATHS] = w->death_count; I[GBI_BADGES] = badges; I[GBI_OPPONENT_LEVEL] = w->max_opponent_level;
    I[GBI_MET_BILL] = BITOF(g, 0xD7F1, 0); I[GBI_USED_CELL_SEPARATOR] = BITOF(g, 0xD7F2, 3); I[GBI_SS_TICKET] = BITOF(g, 0xD7F2, 4);
    I[GBI_MET_BILL_2] = BITOF(g, 0xD7F2, 5); I[GBI_BILL_SAID] = BITOF(g, 0xD7F2, 6); I[GBI_LEFT_BILLS_HOUSE] = BITOF(g, 0xD7F2, 7);
    I[GBI_GOT_HM01] = BITOF(g, 0xD803, 0); I[GBI_RUBBED_CAPTAINS_BACK] = BITOF(g, 0xD803, 1);
    I[GBI_MAPS_EXPLORED] = w->n_seen_maps; I[GBI_PARTY_SIZE] = party_size; I[GBI_HIGHEST_LEVEL] = max_level; I[GBI_TOTAL_PARTY_LEVEL] = level_sum;
    I[GBI_EVENT] = events; I[GBI_MONEY] = money; I[GBI_SEEN_POKEMON] = n_seen; I[GBI_CAUGHT_POKEMON] = n_caught; I[GBI_MOVES_OBTAINED] = n_moves;
    I[GBI_BILL_SAVED] = bill_state; I[GBI_HM_COUNT] = hm_count; I[GBI_CUT_TAUGHT] = w->cut; I[GBI_BILL_CAPT] = bill_capt_rew / 5.0;
    I[GBI_CUT_COORDS] = cut_coords; I[GBI_CUT_TILES] = cut_tiles; I[GBI_BAG_MENU] = bag_menu; I[GBI_STATS_MENU] = stats_menu;
    I[GBI_POKEMON_MENU] = pokemon_menu; I[GBI_START_MENU] = start_menu; I[GBI_USED_CUT] = w->used_cut;
    I[GBI_DEFEATED_DOJO] = BITOF(g, 0xD7B1, 0); I[GBI_GOT_HITMONLEE] = 3 * BITOF(g, 0xD7B1, 6); I[GBI_GOT_HITMONCHAN] = 3 * BITOF(g, 0xD7B1, 7);
    I[GBI_R_DELTA] = reward; I[GBI_R_EVENT] = w->max_events; I[GBI_R_LEVEL] = level_reward; I[GBI_R_OPPONENT_LEVEL] = 0.006 * w->max_opponent_level;
    I[GBI_R_BADGES] = badges_reward; I[GBI_R_BILL_SAVED] = bill_reward; I[GBI_R_HM_COUNT] = hm_reward; I[GBI_R_HEALING] = w->total_healing;
    I[GBI_R_EXPLORATION] = exploration_reward; I[GBI_R_TREE_DISTANCE] = tree_distance_reward; I[GBI_R_DOJO_OLD] = dojo_reward;
    I[GBI_R_ITEMS] = w->item_reward[0] + w->item_reward[1] + w->item_reward[2] + w->item_reward[3] + w->item_reward[4];
    I[GBI_R_USED_CUT] = cut_rew; I[GBI_R_ABS] = reward_abs; I[GBI_SEEN_COORDS] = w->n_seen_coords; I[GBI_DONE] = *done;
    for (int k = 0; k < 5; k++) I[GBI_R_LEMONADE + k] = w->item_reward[k];
    render(w, g, obs); /* :1812 */
    return reward;
}

double pg_step(PgWrapper *w, GbCore *g, int action, uint8_t *obs, int *done) {
    gb_run_action(g, action, 24); /* :1337 */
    return pg_after_emulation(w, g, action, obs, done);
}

void pg_info(const PgWrapper *w, const GbCore *g, double *out) {
    (void)g;
    memcpy(out, w->info, sizeof(double) * PG_INFO_SCALARS);
    long long s = 0;
    for (int i = 0; i < 444 * 436; i++) s += w->counts_map[i];
    out[GBI_COORD_SUM] = (double)s;
}

void pg_counts_map(const PgWrapper *w, int32_t *out) { memcpy(out, w->counts_map, sizeof(int32_t) * 444 * 436); }

int pg_digest(const PgWrapper *w, double *out, int n) {
    double v[] = {(double)w->reset_count, (double)w->is_dead, (double)w->last_map, (double)w->time, w->last_reward, (double)w->max_events,
                  (double)w->max_level_sum, (double)w->n_seen_coords, (double)w->n_seen_maps, (double)w->death_count, w->total_healing, w->last_hp,
                  (double)w->last_party_size, (double)w->cut, (double)w->used_cut, (double)w->n_cut_coords, (double)w->n_cut_tiles,
                  (double)w->n_cut_state, (double)w->seen_start_menu, (double)w->seen_pokemon_menu, (double)w->seen_stats_menu,
                  (double)w->seen_bag_menu, (double)w->last_map_id_plus1, w->item_reward[0], w->item_reward[1], w->item_reward[2],
                  w->item_reward[3], w->item_reward[4], (double)w->max_opponent_level, (double)w->have_last_reward};
    int m = (int)(sizeof(v) / sizeof(v[0]));
    if (m > n) m = n;
    memcpy(out, v, sizeof(double) * (size_t)m);
    return m;
}
