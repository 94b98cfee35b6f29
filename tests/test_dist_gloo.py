"""N > 1 host logic on CPU: world_size-2 gloo group, env sharding with no data-path collective, sum
all-reduce of the episode-info vector, max-over-ranks timing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pokegym_b200.dist import all_reduce_info, max_over_ranks, shard_range


def test_shard_range_partitions_exactly():
    for n, w in ((32768, 8), (4096, 2), (10, 4), (3, 8), (0, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, n_total, steps, lib_path, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pokegym_b200 import _capi
    from pokegym_b200.tools import synth_rom

    lo, hi = shard_range(n_total, rank, world)
    n = hi - lo
    lib = _capi.GbEnvLib(lib_path, "oracle_")  # tests may use the oracle; the product path is CUDA-only
    h = _capi.Handle(lib, n, synth_rom.build_pokelike_rom())
    h.tick(30, True)
    obs = np.zeros((n, _capi.OBS_BYTES), np.uint8)
    rew = np.zeros(n)
    done = np.zeros(n, np.uint8)
    h.reset(obs)
    actions = np.random.default_rng(0).integers(0, 8, (steps, n_total)).astype(np.uint8)  # global action table, sliced per rank
    for s in range(steps):
        h.step(np.ascontiguousarray(actions[s, lo:hi]), obs, rew, done)
    local = np.zeros(_capi.INFO_SCALARS)
    h.reduce_info(local)
    total = torch.from_numpy(local.copy())
    all_reduce_info(total)
    t = max_over_ranks(0.5 + rank, None)
    q.put((rank, local, total.numpy(), t, float(rew.sum())))
    dist.destroy_process_group()


def test_two_rank_sharded_run_matches_single_process(built):
    n_total, steps, world = 6, 6, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, steps, str(built["oracle"]), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process run over all envs
    from pokegym_b200 import _capi
    from pokegym_b200.tools import synth_rom

    lib = _capi.GbEnvLib(built["oracle"], "oracle_")
    h = _capi.Handle(lib, n_total, synth_rom.build_pokelike_rom())
    h.tick(30, True)
    obs = np.zeros((n_total, _capi.OBS_BYTES), np.uint8)
    rew = np.zeros(n_total)
    done = np.zeros(n_total, np.uint8)
    h.reset(obs)
    actions = np.random.default_rng(0).integers(0, 8, (steps, n_total)).astype(np.uint8)
    for s in range(steps):
        h.step(actions[s], obs, rew, done)
    ref = np.zeros(_capi.INFO_SCALARS)
    h.reduce_info(ref)
    assert np.array_equal(res[0][2], res[1][2]), "ranks disagree after the all-reduce"
    assert np.allclose(res[0][2], ref, rtol=0, atol=1e-9), "sharded sum differs from the single-process sum"
    assert res[0][2][0] == n_total  # slot 0 carries the env count
    assert res[0][3] == res[1][3] == 1.5  # max over ranks
