"""Names of the per-env info row (include/gbenv_info.h) and its conversion to the reference's info dict
layout (`info["stats"]`, `info["reward"]`; /root/reference/pokegym/environment.py:1621-1703)."""
from __future__ import annotations

import re
from pathlib import Path
from typing import Dict, List

import numpy as np


def _names_from_header() -> List[str]:
    text = (Path(__file__).resolve().parent.parent / "include" / "gbenv_info.h").read_text()
    body = text[text.index("enum {") : text.index("GBI__END")]
    names: List[str] = []
    for tok in re.findall(r"GBI_([A-Z0-9_]+)(?:\s*=\s*([^,]+))?,", body):
        name, init = tok
        if name == "LEVEL0":
            names += [f"level{k}" for k in range(6)]
        elif name == "LEVELS_SUM" and init:
            names.append("levels_sum")
        elif name != "COUNT" or not names:
            names.append(name.lower())
    return names


INFO_NAMES: List[str] = _names_from_header()
INFO_INDEX: Dict[str, int] = {n: i for i, n in enumerate(INFO_NAMES)}
assert len(INFO_NAMES) <= 64 and INFO_NAMES[0] == "count", INFO_NAMES[:3]

_REWARD_KEYS = {"r_delta": "delta", "r_event": "event", "r_level": "level", "r_opponent_level": "opponent_level", "r_badges": "badges",
                "r_bill_saved": "bill_saved_reward", "r_hm_count": "hm_count_reward", "r_healing": "healing", "r_exploration": "exploration",
                "r_tree_distance": "tree_distance_reward", "r_dojo_old": "dojo_reward_old", "r_used_cut": "used_cut_reward"}


def info_row_to_dict(row: np.ndarray) -> dict:
    """One info row -> {"stats": {...}, "reward": {...}} with the reference's key names where they exist."""
    stats, reward = {}, {}
    for i, name in enumerate(INFO_NAMES):
        v = float(row[i])
        if name in _REWARD_KEYS:
            reward[_REWARD_KEYS[name]] = v
        elif name.startswith("r_"):
            reward[name[2:]] = v
        elif name not in ("count",):
            stats[name] = v
    stats["levels"] = [stats.pop(f"level{k}") for k in range(6)]
    for k in range(1, 7):
        stats[f"badge_{k}"] = float(stats["badges"] >= k)
    return {"stats": stats, "reward": reward}


def info_sum_to_means(total: np.ndarray) -> Dict[str, float]:
    """All-reduced sum vector -> per-env means (slot 0 carries the env count)."""
    n = max(float(total[0]), 1.0)
    return {name: float(total[i]) / n for i, name in enumerate(INFO_NAMES) if name != "count"}
