/*
 * gbenv_info.h -- slot layout of the per-env episode-info row (GBENV_INFO_SCALARS doubles).
 *
 * Numeric entries of the reference's `info["stats"]` and `info["reward"]` dicts
 * (/root/reference/pokegym/environment.py:1621-1703), refreshed by every gbenv_step.  Multi-GPU runs
 * sum these rows over envs on each rank and all-reduce the GBENV_INFO_SCALARS-double vector over NCCL
 * (SURVEY.md section 8e); slot 0 carries the env count so means can be formed afterwards.
 */
#ifndef GBENV_INFO_H
#define GBENV_INFO_H

enum {
    GBI_COUNT = 0,
    GBI_STEP,
    GBI_X,
    GBI_Y,
    GBI_MAP,
    GBI_PCOUNT,
    GBI_LEVEL0, /* 6 slots */
    GBI_LEVELS_SUM = GBI_LEVEL0 + 6,
    GBI_COORD_SUM, /* np.sum(counts_map) */
    GBI_DEATHS,
    GBI_BADGES,
    GBI_OPPONENT_LEVEL,
    GBI_MET_BILL,
    GBI_USED_CELL_SEPARATOR,
    GBI_SS_TICKET,
    GBI_MET_BILL_2,
    GBI_BILL_SAID,
    GBI_LEFT_BILLS_HOUSE,
    GBI_GOT_HM01,
    GBI_RUBBED_CAPTAINS_BACK,
    GBI_MAPS_EXPLORED,
    GBI_PARTY_SIZE,
    GBI_HIGHEST_LEVEL,
    GBI_TOTAL_PARTY_LEVEL,
    GBI_EVENT,
    GBI_MONEY,
    GBI_SEEN_POKEMON,
    GBI_CAUGHT_POKEMON,
    GBI_MOVES_OBTAINED,
    GBI_BILL_SAVED,
    GBI_HM_COUNT,
    GBI_CUT_TAUGHT,
    GBI_BILL_CAPT,
    GBI_CUT_COORDS,
    GBI_CUT_TILES,
    GBI_BAG_MENU,
    GBI_STATS_MENU,
    GBI_POKEMON_MENU,
    GBI_START_MENU,
    GBI_USED_CUT,
    GBI_DEFEATED_DOJO,
    GBI_GOT_HITMONLEE,
    GBI_GOT_HITMONCHAN,
    GBI_R_DELTA,
    GBI_R_EVENT,
    GBI_R_LEVEL,
    GBI_R_OPPONENT_LEVEL,
    GBI_R_BADGES,
    GBI_R_BILL_SAVED,
    GBI_R_HM_COUNT,
    GBI_R_HEALING,
    GBI_R_EXPLORATION,
    GBI_R_TREE_DISTANCE,
    GBI_R_DOJO_OLD,
    GBI_R_ITEMS,
    GBI_R_USED_CUT,
    GBI_R_ABS,
    GBI_SEEN_COORDS,
    GBI_DONE,
    GBI_R_LEMONADE, /* has_<item>_in_bag_reward: 0 or 20, sticky for the env lifetime (environment.py:1358-1372) */
    GBI_R_SILPH_SCOPE,
    GBI_R_LIFT_KEY,
    GBI_R_POKEDOLL,
    GBI_R_BICYCLE,
    GBI__END
};

#if defined(__cplusplus)
static_assert(GBI__END <= 72, "info row overflows GBENV_INFO_SCALARS");
#else
typedef char gbi_fits_in_row[(GBI__END <= 72) ? 1 : -1];
#endif

#endif
