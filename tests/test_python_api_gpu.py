"""The Python drop-in boundary on the GPU: Environment (single env, NumPy out) and VecEnvironment."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_environment_facade_matches_pokegym_signatures(cuda_lib, oracle_lib, roms):
    from pokegym_b200 import Environment, _capi

    rom = roms("pokelike")
    env = Environment(rom_path=rom)
    assert env.observation_space.shape == (72, 80, 4) and env.observation_space.dtype == np.uint8 and env.action_space.n == 8
    obs, info = env.reset(max_episode_steps=5)
    assert obs.shape == (72, 80, 4) and obs.dtype == np.uint8 and info == {}
    cpu = _capi.Handle(oracle_lib, 1, rom)
    cpu.tick(60, True)
    o = np.zeros((1, _capi.OBS_BYTES), np.uint8)
    r = np.zeros(1)
    d = np.zeros(1, np.uint8)
    cpu.reset(o, max_episode_steps=5)
    assert np.array_equal(obs, o.reshape(72, 80, 4))
    for s, a in enumerate([4, 0, 3, 6, 1]):
        obs, rew, term, trunc, info = env.step(a)
        cpu.step(np.array([a], np.uint8), o, r, d)
        assert isinstance(rew, float) and rew == r[0] and term == trunc == bool(d[0])
        assert np.array_equal(obs, o.reshape(72, 80, 4)) and np.array_equal(env.render(), obs)
        assert bool(info) == (s == 4)
    assert info["stats"]["step"] == 5 and "pokemon_exploration_map" in info and info["pokemon_exploration_map"].shape == (444, 436)
    # the complete reference layout (environment.py:1621-1810): 11 top-level sections, 130 named event flags
    assert list(info) == ["pokemon_exploration_map", "stats", "reward", "detailed_rewards_silph_co", "detailed_rewards_dojo", "detailed_rewards_hideout",
                          "detailed_rewards_poke_tower", "detailed_rewards_gyms", "silph_co_events_aggregate", "dojo_events_aggregate",
                          "hideout_events_aggregate", "poke_tower_events_aggregate", "gym_events"]
    n_flags = sum(len(info[k]) for k in info if k.endswith("_aggregate")) + sum(len(v) for v in info["gym_events"].values())
    assert n_flags == 130 and info["reward"]["delta"] == rew
    env.close()


def test_vec_environment_rollout_and_auto_reset(cuda_lib, roms):
    import torch

    from pokegym_b200 import VecEnvironment
    from pokegym_b200.puffer import PufferVecAdaptor

    n, T = 96, 4
    rollout = torch.zeros((T, n, 72, 80, 4), dtype=torch.uint8, device="cuda")
    vec = VecEnvironment(n, roms("pokelike"), rollout=rollout, max_episode_steps=6, auto_reset=True)
    obs, _ = vec.reset()
    assert obs.data_ptr() == rollout[0].data_ptr()  # observations land in the rollout tensor, no copy
    g = torch.Generator(device="cuda").manual_seed(0)
    seen_done = False
    for t in range(1, 9):
        a = torch.randint(0, 8, (n,), generator=g, device="cuda", dtype=torch.uint8)
        obs, rew, done, trunc, _ = vec.step(a)
        assert obs.data_ptr() == rollout[t % T].data_ptr() and rew.dtype == torch.float64 and done.dtype == torch.bool
        seen_done |= bool(done.all())
    assert seen_done
    s = vec.info_sum().cpu().numpy()
    assert s[0] == n
    c = VecEnvironment.compact_view(obs)  # (grey, visited): channels 0-2 of the reference layout are identical
    assert c.shape == (n, 72, 80, 2) and c.data_ptr() == obs.data_ptr() + 2
    assert torch.equal(obs[..., 0], c[..., 0]) and torch.equal(obs[..., 1], c[..., 0]) and torch.equal(obs[..., 3], c[..., 1])
    full = vec.full_info(5)  # complete reference-shaped info dict for one env of the batch
    assert full["stats"]["step"] >= 1 and len(full["silph_co_events_aggregate"]) == 53
    vec.close()


def test_puffer_adaptor_contract(cuda_lib, oracle_lib, roms):
    """PufferLib's vector contract (pufferlib.vector.Serial.send): an env that reported done is RESET on the next send -- its
    action is ignored, it is not stepped, reward 0 / terminal False and the reset observation are delivered -- while the other
    envs step.  Checked against the oracle driven the same way, env by env."""
    import torch

    from pokegym_b200 import VecEnvironment, _capi
    from pokegym_b200.puffer import PufferVecAdaptor

    n = 40
    rom = roms("pokelike")
    vec = VecEnvironment(n, rom, max_episode_steps=5)
    ad = PufferVecAdaptor(vec)
    cpu = _capi.Handle(oracle_lib, n, rom)
    cpu.tick(60, True)
    oc, rc, dc = np.zeros((n, _capi.OBS_BYTES), np.uint8), np.zeros(n), np.zeros(n, np.uint8)
    cpu.reset(oc, max_episode_steps=5)
    ad.async_reset()
    flat, rew, term, trunc, infos, ids, mask = ad.recv()
    assert flat.shape == (n, 23040) and len(ids) == n and mask.all() and np.array_equal(flat.cpu().numpy(), oc)
    # stagger the episodes: envs 0..9 get a shorter first episode
    short = np.zeros(n, np.uint8)
    short[:10] = 1
    vec.reset(mask=short, max_episode_steps=3)
    vec.max_episode_steps = 5
    cpu.reset(oc, mask=short, max_episode_steps=3)
    pending = np.zeros(n, np.uint8)
    rng = np.random.default_rng(9)
    n_infos = 0
    for t in range(14):
        act = rng.integers(0, 8, n).astype(np.uint8)
        ad.send(torch.from_numpy(act).cuda())
        flat, rew, term, trunc, infos, ids, mask = ad.recv()
        if pending.any():
            cpu.reset(oc, mask=pending, max_episode_steps=5)
        cpu.step_masked(act, pending, oc, rc, dc)
        assert np.array_equal(rew.cpu().numpy(), rc), t
        assert np.array_equal(term.cpu().numpy().astype(np.uint8), dc) and np.array_equal(trunc.cpu().numpy().astype(np.uint8), dc), t
        assert np.array_equal(flat.cpu().numpy(), oc), t
        assert (rc[pending.astype(bool)] == 0).all() and not dc[pending.astype(bool)].any()
        assert len(infos) == int(dc.sum())
        n_infos += len(infos)
        pending = dc.copy()
    assert n_infos >= n  # every env finished at least one episode
    for e in (0, 9, 10, n - 1):
        assert vec.save_state(e) == cpu.save_state(e)
    vec.close()


def test_mixed_initial_states(cuda_lib, oracle_lib, roms):
    """BASELINE.json config 5 shape: env i resets from state i mod K (here K synthetic states)."""
    import torch

    from pokegym_b200 import VecEnvironment, _capi

    rom = roms("pokelike")
    src = _capi.Handle(oracle_lib, 1, rom)
    blobs = []
    for k in range(3):
        src.tick(25 + 7 * k, True)
        src.run_action(np.array([k], np.uint8))
        blobs.append(src.save_state(0))
    vec = VecEnvironment(7, rom, state_paths=blobs)
    vec.reset()
    for e in range(7):
        assert vec.save_state(e) == _loaded(oracle_lib, rom, blobs[e % 3]), e
    vec.close()


def _loaded(oracle_lib, rom, blob):
    from pokegym_b200 import _capi

    h = _capi.Handle(oracle_lib, 1, rom)
    h.set_initial_template(h.add_state_template(blob))
    o = np.zeros((1, _capi.OBS_BYTES), np.uint8)
    h.reset(o)
    return h.save_state(0)


def test_error_paths_and_limits(cuda_lib, roms, monkeypatch):
    import torch

    from pokegym_b200 import _capi

    rom = roms("pokelike")
    with pytest.raises(_capi.GbEnvError):
        _capi.Handle(cuda_lib, 0, rom)  # no envs
    with pytest.raises(_capi.GbEnvError):
        _capi.Handle(cuda_lib, 4, rom[:1000])  # not a ROM image
    with pytest.raises(_capi.GbEnvError):
        _capi.Handle(cuda_lib, 4, rom, device_id=99)
    h = _capi.Handle(cuda_lib, 4, rom)
    with pytest.raises(_capi.GbEnvError, match="save-state"):
        h.add_state_template(b"\x09" + bytes(142_609 - 1))  # wrong length
    bad = bytearray(h.save_state(0))
    bad[9125] = 0x77  # a framebuffer word PyBoy never produces
    with pytest.raises(_capi.GbEnvError, match="framebuffer"):
        h.add_state_template(bytes(bad))
    with pytest.raises(_capi.GbEnvError):
        h.load_template(5)
    with pytest.raises(_capi.GbEnvError):
        h.read_mem(9, 0xC000, 1)
    with pytest.raises(_capi.GbEnvError):
        h.set_lanes_per_warp(33)
    obs = torch.zeros((4, _capi.OBS_BYTES), dtype=torch.uint8, device="cuda")
    with pytest.raises(_capi.GbEnvError):
        h.reset(obs, obs_stride=100)
    # results do not depend on the lanes-per-warp tuning knob
    states = []
    for lanes in (1, 3, 12, 32):
        g = _capi.Handle(cuda_lib, 40, rom)
        g.set_lanes_per_warp(lanes)
        g.tick(20, True)
        a = torch.arange(40, dtype=torch.uint8, device="cuda") % 8
        for _ in range(3):
            g.run_action(a)
        states.append([g.save_state(e) for e in (0, 13, 39)])
    assert states[0] == states[1] == states[2]
    # exploration storage: a visited-bitmap pool that runs dry is an error, not a silent loss
    monkeypatch.setenv("GBENV_VISITED_PAGES", "3")
    monkeypatch.setenv("GBENV_COUNTS_MAP", "0")
    s = _capi.Handle(cuda_lib, 2, rom)
    s.tick(30, True)
    rew = torch.zeros(2, dtype=torch.float64, device="cuda")
    done = torch.zeros(2, dtype=torch.uint8, device="cuda")
    obs2 = torch.zeros((2, _capi.OBS_BYTES), dtype=torch.uint8, device="cuda")
    s.reset(obs2)  # one page per env
    s.check()
    act = torch.zeros(2, dtype=torch.uint8, device="cuda")
    s.write_mem(1, 0xD35E, [1])  # teleport env 1 to another map: the third and last page
    s.step(act, obs2, rew, done)
    s.check()
    s.write_mem(1, 0xD35E, [2])
    s.step(act, obs2, rew, done)
    with pytest.raises(_capi.GbEnvError, match="visited-bitmap page pool"):
        s.check()
    with pytest.raises(_capi.GbEnvError, match="exhausted"):
        s.step(act, obs2, rew, done)
    assert s.counters().faults == 1
    with pytest.raises(_capi.GbEnvError, match="exploration map"):
        s.counts_map(0)


def test_env_groups_match_one_handle_and_the_oracle(cuda_lib, oracle_lib, roms):
    """pokegym_b200.EnvGroups: the batch held as two handles stepping on two CUDA streams.  Env i of the batch is env
    i - g * n of group g, so observations, rewards, dones and the reduced info vector must equal the oracle's for one
    batch of the same size, step by step, including a masked reset in the middle."""
    import torch

    from pokegym_b200 import EnvGroups, _capi

    E, G = 128, 2
    rom = roms("pokelike")
    groups = EnvGroups(cuda_lib, E, rom, n_groups=G)
    groups.for_each(lambda h, g, st: h.tick(60, True, stream=st))
    cpu = _capi.Handle(oracle_lib, E, rom)
    cpu.tick(60, True)
    dev = groups.device
    obs = torch.zeros((E, _capi.OBS_BYTES), dtype=torch.uint8, device=dev)
    rew = torch.zeros(E, dtype=torch.float64, device=dev)
    done = torch.zeros(E, dtype=torch.uint8, device=dev)
    oc, rc, dc = np.zeros((E, _capi.OBS_BYTES), np.uint8), np.zeros(E), np.zeros(E, np.uint8)
    groups.reset(obs, max_episode_steps=6)
    cpu.reset(oc, max_episode_steps=6)
    torch.cuda.synchronize()
    assert np.array_equal(obs.cpu().numpy(), oc)
    rng = np.random.default_rng(21)
    for t in range(9):
        act = rng.integers(0, 8, E).astype(np.uint8)
        groups.step(torch.from_numpy(act).to(dev), obs, rew, done)  # returns with both groups still running
        groups.join()
        cpu.step(act, oc, rc, dc)
        torch.cuda.synchronize()
        assert np.array_equal(rew.cpu().numpy(), rc), t
        assert np.array_equal(done.cpu().numpy(), dc), t
        assert np.array_equal(obs.cpu().numpy(), oc), t
        if dc.any():
            groups.reset(obs, mask=done.clone(), max_episode_steps=6)
            cpu.reset(oc, mask=dc.copy(), max_episode_steps=6)
            torch.cuda.synchronize()
            assert np.array_equal(obs.cpu().numpy(), oc), t
    s_g = torch.zeros(_capi.INFO_SCALARS, dtype=torch.float64, device=dev)
    groups.reduce_info(s_g)
    s_c = np.zeros(_capi.INFO_SCALARS)
    cpu.reduce_info(s_c)
    torch.cuda.synchronize()
    assert np.allclose(s_g.cpu().numpy(), s_c, rtol=1e-12, atol=0)  # partial sums per group: the association differs
    assert groups.counters().instructions == cpu.counters().instructions
    for e in (0, E // G - 1, E // G, E - 1):
        assert groups.handles[e // (E // G)].save_state(e % (E // G)) == cpu.save_state(e)
    groups.close()
    with pytest.raises(ValueError):
        EnvGroups(cuda_lib, 100, rom, n_groups=3)
