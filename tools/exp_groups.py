#!/usr/bin/env python
"""Tuning experiment: E envs as G independent groups (handles) stepped on G CUDA streams, so that the tail of one group's
emulation kernel (its few slowest envs) overlaps the other groups' work.  usage: exp_groups.py E G [steps] [lanes]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from pokegym_b200 import _capi
import __graft_entry__ as g

E, G = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 0
import os
lib = _capi.GbEnvLib(os.environ.get("GBENV_LIB") or g.build_cuda())
rom = bench.build_rom("pokelike")
dev = torch.device("cuda", 0)
n = E // G
hs = [_capi.Handle(lib, n, rom, device_id=0) for _ in range(G)]
streams = [torch.cuda.Stream() for _ in range(G)]
for h in hs:
    if lanes:
        h.set_lanes_per_warp(lanes)
    h.tick(60, True)
ring = [torch.zeros((4, n, _capi.OBS_BYTES), dtype=torch.uint8, device=dev) for _ in range(G)]
rew = [torch.zeros(n, dtype=torch.float64, device=dev) for _ in range(G)]
done = [torch.zeros(n, dtype=torch.uint8, device=dev) for _ in range(G)]
gen = torch.Generator(device=dev); gen.manual_seed(7)
acts = [torch.randint(0, 8, (256, n), generator=gen, device=dev, dtype=torch.uint8) for _ in range(G)]
torch.cuda.synchronize()
for k in range(G):
    hs[k].reset(ring[k][0], stream=streams[k].cuda_stream)
pre = 150
def run(a, b):
    for i in range(a, b):
        for k in range(G):
            hs[k].step(acts[k][i % 256], ring[k][i % 4], rew[k], done[k], stream=streams[k].cuda_stream)
run(0, pre)
torch.cuda.synchronize()
t0 = time.perf_counter()
run(pre, pre + steps)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"E={E} G={G} lanes={hs[0].lanes_per_warp()} env-steps/s {E * steps / dt:.0f}  ms/step {1000 * dt / steps:.2f}")
