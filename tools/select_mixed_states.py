#!/usr/bin/env python
"""Pick a small, diverse set of the reference's PyBoy save-states for BASELINE.json config 5 (divergence stress:
"envs reset from mixed overworld/battle/menu save states") and write them to tests/golden/red_states_mixed.npz.

The reference ships 264 v9 states (bin/checkpoints_{battles,bill,pallet}/**, unused_states/, current_state/, backups/).
Selection is greedy over a feature tuple that decides which code an env runs next and which branch of the wrapper's
Game.process_game_states classifier (red_ram_api.py:59-73, :149-225, :542-602) it takes: ROM bank, PC, HALT flag, battle
type D057, pre-battle flag D059, text/sprite flag CFC4, whether the menu cursor CC30/CC31 is one of the 48 known
locations, CD38, map id, party size, LCDC window/sprite bits.  A state is taken when it shows a feature VALUE not seen
yet; then the set is topped up to `--count` with the states farthest (Hamming, over WRAM) from those already chosen.

Run in the build container (needs /root/reference).  The output travels with the repo; the GPU box never reads
/root/reference."""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pokegym_b200.state_file import parse_state  # noqa: E402

REF = Path("/root/reference/pokegym")
CURSOR_KEYS = {0xC3D3, 0xC3FB, 0xC423, 0xC44B, 0xC473, 0xC49B, 0xC4C3, 0xC44C, 0xC474, 0xC3B5, 0xC3DD, 0xC405, 0xC3C9, 0xC3F1, 0xC419, 0xC441,
               0xC49A, 0xC4C2, 0xC4EA, 0xC4C1, 0xC4A9, 0xC4BD, 0xC4D1, 0xC4E5, 0xC4C7, 0xC3B4, 0xC3DC, 0xC404, 0xC42C, 0xC454, 0xC47C, 0xC49C,
               0xC4C4, 0xC4EC, 0xC4E9, 0xC48A, 0xC4B2, 0xC3F5, 0xC41D, 0xC445, 0x0169, 0xC4EF, 0xC44F, 0xC477, 0xC469, 0xC459, 0xC46D, 0xC481}


def features(st):
    c, w = st.cpu, st.raw["wram"]
    m = lambda a: int(w[a - 0xC000])
    cur = m(0xCC30) | (m(0xCC31) << 8)
    return {
        "bank": int(st.raw["mbc"][0]), "pc": c["PC"], "halted": c["halted"], "ime": c["IME"], "battle": m(0xD057), "prebattle": m(0xD059),
        "cfc4": int(m(0xCFC4) != 0), "cursor_known": int(cur in CURSOR_KEYS), "cursor_zero": int(cur == 0), "cd38": int(m(0xCD38) != 0),
        "map": m(0xD35E), "party": m(0xD163), "lcdc": int(st.raw["lcd_regs"][0]) & 0x63, "d125": m(0xD125), "cc52": int(m(0xCC52) == 0),
        "ly": int(st.raw["lcd_regs"][5]) // 16, "ram_en": int(st.raw["mbc"][2]),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=40)
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden" / "red_states_mixed.npz"))
    args = ap.parse_args()
    paths = sorted(p for p in REF.rglob("*") if p.is_file() and p.stat().st_size == 142_610)
    blobs = [p.read_bytes() for p in paths]
    feats = [features(parse_state(b)) for b in blobs]
    seen, chosen = {}, []
    for i, f in enumerate(feats):
        new = [k for k, v in f.items() if v not in seen.setdefault(k, set())]
        if new:
            chosen.append(i)
            for k, v in f.items():
                seen[k].add(v)
    wr = np.stack([parse_state(b).raw["wram"] for b in blobs])
    if len(chosen) > args.count:  # keep the ones that contributed most: re-run greedily by number of new values
        order, seen2, keep = list(chosen), {}, []
        while order and len(keep) < args.count:
            best = max(order, key=lambda i: sum(v not in seen2.setdefault(k, set()) for k, v in feats[i].items()))
            keep.append(best)
            order.remove(best)
            for k, v in feats[best].items():
                seen2[k].add(v)
        chosen = sorted(keep)
    while len(chosen) < args.count:
        d = np.min(np.stack([(wr != wr[j]).sum(axis=1) for j in chosen]), axis=0)
        d[chosen] = -1
        chosen.append(int(np.argmax(d)))
    chosen = sorted(chosen)
    names = [str(paths[i].relative_to(REF)) for i in chosen]
    np.savez_compressed(args.out, states=np.stack([np.frombuffer(blobs[i], dtype=np.uint8) for i in chosen]), names=np.array(names))
    for i in chosen:
        f = feats[i]
        print(f"{str(paths[i].relative_to(REF)):55s} bank {f['bank']:2d} pc {f['pc']:04x} halt {f['halted']} battle {f['battle']} pre {f['prebattle']} map {f['map']:3d} party {f['party']}")
    print(len(chosen), "states ->", args.out, Path(args.out).stat().st_size // 1024, "KiB")


if __name__ == "__main__":
    main()
