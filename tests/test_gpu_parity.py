"""CUDA path vs CPU oracle, bit-exact, through the C ABI (run on the B200 box: -m gpu).

The comparator is the PyBoy v9 save-state itself: after every few steps both implementations dump
every env they hold to a 142,610-byte blob (registers, VRAM, OAM, WRAM, HRAM, IO, LCD clocks, timer,
MBC, cart RAM, joypad, per-scanline parameters and the full framebuffer) and the blobs must be equal.
"""
import numpy as np
import pytest

from pokegym_b200 import _capi
from pokegym_b200.state_file import diff_states

pytestmark = pytest.mark.gpu


def _pair(cuda_lib, oracle_lib, rom, n):
    return _capi.Handle(cuda_lib, n, rom, 0), _capi.Handle(oracle_lib, n, rom)


def _assert_same_states(gpu, cpu, envs, where):
    for e in envs:
        a, b = gpu.save_state(e), cpu.save_state(e)
        if a != b:
            raise AssertionError(f"{where}: env {e} differs: {diff_states(b, a)[:8]}")
        xa, xb = gpu.core_extra(e), cpu.core_extra(e)
        assert (xa.stat_mode, xa.ly_window, xa.fault) == (xb.stat_mode, xb.ly_window, xb.fault), where


@pytest.mark.parametrize("rom_name,n,steps", [("pokelike", 72, 40), ("conformance", 40, 60), ("conformance_b", 40, 40), ("pokelike_timer", 33, 60), ("busy", 32, 30), ("halt_edge", 48, 120), ("lcd_probe", 40, 40), ("lcd_probe_b", 33, 40)])
def test_run_action_matches_oracle(cuda_lib, oracle_lib, roms, rom_name, n, steps):
    import torch

    gpu, cpu = _pair(cuda_lib, oracle_lib, roms(rom_name), n)
    _assert_same_states(gpu, cpu, [0, n - 1], "power-on")
    gpu.tick(3, True)
    cpu.tick(3, True)
    _assert_same_states(gpu, cpu, [0, n // 2, n - 1], "after 3 ticks")
    rng = np.random.default_rng(123)
    for s in range(steps):
        act = rng.integers(0, 8, n).astype(np.uint8)
        gpu.run_action(torch.from_numpy(act).cuda())
        cpu.run_action(act)
        if s % 4 == 3 or s == steps - 1:
            _assert_same_states(gpu, cpu, range(n) if s == steps - 1 else [0, 1, n - 1], f"{rom_name} step {s}")
    cg, cc = gpu.counters(), cpu.counters()
    assert cg.faults == 0 and cc.faults == 0
    assert cg.instructions == cc.instructions and cg.cycles == cc.cycles


def test_step_reward_obs_match_oracle(cuda_lib, oracle_lib, roms):
    import torch

    n, steps = 72, 60
    gpu, cpu = _pair(cuda_lib, oracle_lib, roms("pokelike"), n)
    gpu.tick(40, True)
    cpu.tick(40, True)
    og = torch.zeros((n, _capi.OBS_BYTES), dtype=torch.uint8, device="cuda")
    rg = torch.zeros(n, dtype=torch.float64, device="cuda")
    dg = torch.zeros(n, dtype=torch.uint8, device="cuda")
    oc = np.zeros((n, _capi.OBS_BYTES), dtype=np.uint8)
    rc = np.zeros(n)
    dc = np.zeros(n, dtype=np.uint8)
    gpu.reset(og, max_episode_steps=50)
    cpu.reset(oc, max_episode_steps=50)
    assert np.array_equal(og.cpu().numpy(), oc)
    rng = np.random.default_rng(5)
    total = 0.0
    for s in range(steps):
        act = rng.integers(0, 8, n).astype(np.uint8)
        gpu.step(torch.from_numpy(act).cuda(), og, rg, dg)
        cpu.step(act, oc, rc, dc)
        assert np.array_equal(rg.cpu().numpy(), rc), f"reward differs at step {s}"
        assert np.array_equal(dg.cpu().numpy(), dc), f"done differs at step {s}"
        assert np.array_equal(og.cpu().numpy(), oc), f"obs differs at step {s}"
        total += float(np.abs(rc).sum())
        if s == 49:
            assert dc.all()
            m = np.zeros(n, dtype=np.uint8)
            m[::2] = 1  # reset every other env: the others keep running past done, as the reference allows
            gpu.reset(og, mask=m, max_episode_steps=50)
            cpu.reset(oc, mask=m, max_episode_steps=50)
            assert np.array_equal(og.cpu().numpy(), oc)
    assert total > 0
    _assert_same_states(gpu, cpu, range(0, n, 7), "after steps")
    ig = torch.zeros((n, _capi.INFO_SCALARS), dtype=torch.float64, device="cuda")
    ic = np.zeros((n, _capi.INFO_SCALARS))
    gpu.get_info(ig)
    cpu.get_info(ic)
    gpu.sync()
    assert np.array_equal(ig.cpu().numpy(), ic)
    assert np.array_equal(gpu.counts_map(3), cpu.counts_map(3))


@pytest.mark.parametrize("lanes", [1, 8, 32])
def test_config5_mixed_reference_states_divergence_stress(cuda_lib, oracle_lib, roms, lanes):
    """BASELINE.json config 5: env i is reset from reference save-state i mod 40 (tests/golden/red_states_mixed.npz: overworld,
    wild and trainer battles, menus, text boxes, ROM banks 1-38, HALTed and mid-instruction PCs, picked by
    tools/select_mixed_states.py), so neighbouring lanes of a warp run unrelated code from step one.  200 random-action steps,
    every step: reward / done / observation equal to the oracle; full emulator state of every env every 50 steps.  Illegal
    opcodes are expected (the states' PCs land in the synthetic ROM's data) and must match too."""
    import torch
    from helpers import GOLDEN

    blobs = [b.tobytes() for b in np.load(GOLDEN / "red_states_mixed.npz")["states"]]
    n, steps = 160, 200
    gpu, cpu = _pair(cuda_lib, oracle_lib, roms("pokelike"), n)
    gpu.set_lanes_per_warp(lanes)
    for k, blob in enumerate(blobs):
        ids = np.arange(k, n, len(blobs), dtype=np.int32)
        for h in (gpu, cpu):
            h.set_initial_template(h.add_state_template(blob), ids)
    og = torch.zeros((n, _capi.OBS_BYTES), dtype=torch.uint8, device="cuda")
    rg = torch.zeros(n, dtype=torch.float64, device="cuda")
    dg = torch.zeros(n, dtype=torch.uint8, device="cuda")
    oc, rc, dc = np.zeros((n, _capi.OBS_BYTES), np.uint8), np.zeros(n), np.zeros(n, np.uint8)
    gpu.reset(og, max_episode_steps=120)
    cpu.reset(oc, max_episode_steps=120)
    assert np.array_equal(og.cpu().numpy(), oc)
    _assert_same_states(gpu, cpu, range(n), "after the first reset")
    rng = np.random.default_rng(55)
    for s in range(steps):
        act = rng.integers(0, 8, n).astype(np.uint8)
        gpu.step(torch.from_numpy(act).cuda(), og, rg, dg)
        cpu.step(act, oc, rc, dc)
        assert np.array_equal(rg.cpu().numpy(), rc), f"reward differs at step {s}"
        assert np.array_equal(dg.cpu().numpy(), dc), f"done differs at step {s}"
        assert np.array_equal(og.cpu().numpy(), oc), f"obs differs at step {s}"
        if s == 119:  # episode end: device-side auto-reset of the finished envs (no state reload after the first reset)
            assert dc.all()
            gpu.reset_dev(og, mask_dev=dg, max_episode_steps=120)
            cpu.reset(oc, mask=dc, max_episode_steps=120)
            assert np.array_equal(og.cpu().numpy(), oc)
        if s % 50 == 49 or s == steps - 1:
            _assert_same_states(gpu, cpu, range(n), f"mixed states step {s}")
    cg, cc = gpu.counters(), cpu.counters()
    assert cg.instructions == cc.instructions and cg.cycles == cc.cycles and cg.faults == cc.faults


def test_step_host_entry_point(cuda_lib, oracle_lib, roms):
    n = 40
    gpu, cpu = _pair(cuda_lib, oracle_lib, roms("pokelike"), n)
    obs = [np.zeros((n, _capi.OBS_BYTES), dtype=np.uint8) for _ in range(2)]
    rew = [np.zeros(n) for _ in range(2)]
    done = [np.zeros(n, dtype=np.uint8) for _ in range(2)]
    gpu.reset_host(obs[0])
    cpu.reset_host(obs[1])
    rng = np.random.default_rng(9)
    for s in range(5):
        act = rng.integers(0, 8, n).astype(np.uint8)
        gpu.step_host(act, obs[0], rew[0], done[0])
        cpu.step_host(act, obs[1], rew[1], done[1])
        assert np.array_equal(obs[0], obs[1]) and np.array_equal(rew[0], rew[1]) and np.array_equal(done[0], done[1])


def test_memory_api_and_inputs(cuda_lib, oracle_lib, roms):
    n = 5
    gpu, cpu = _pair(cuda_lib, oracle_lib, roms("pokelike"), n)
    for h in (gpu, cpu):
        h.tick(10, False)
        h.write_mem(2, 0xC123, [1, 2, 3, 4, 5])
        h.write_mem(2, 0xFF42, [17])  # SCY through the bus
        h.write_mem(2, 0xFF46, [0xC1])  # OAM DMA through the bus
        h.send_input(4, True)
        h.tick(2, True)
        h.send_input(4, False)
        h.tick(2, True)
    for addr, cnt in ((0xC120, 16), (0xFF40, 12), (0xFE00, 160), (0x0100, 80), (0x4000, 16), (0xFF80, 127), (0xA000, 8), (0xE123, 4)):
        assert np.array_equal(gpu.read_mem(2, addr, cnt), cpu.read_mem(2, addr, cnt)), hex(addr)
    assert np.array_equal(gpu.screen(2), cpu.screen(2))
    _assert_same_states(gpu, cpu, range(n), "memory api")


def test_readme_config_72_envs_long_run(cuda_lib, oracle_lib, roms):
    """BASELINE.json config 2 shape: 72 envs (the README training config), random actions, full Environment.step,
    bit-exact RAM / framebuffer / reward / obs versus 72 oracle instances.  2,000 steps by default
    (GBENV_LONG_STEPS=10000 for the full 10k-step run)."""
    import hashlib
    import os

    import torch

    n, steps = 72, int(os.environ.get("GBENV_LONG_STEPS", "2000"))
    gpu, cpu = _pair(cuda_lib, oracle_lib, roms("pokelike"), n)
    gpu.tick(60, True)
    cpu.tick(60, True)
    og = torch.zeros((n, _capi.OBS_BYTES), dtype=torch.uint8, device="cuda")
    rg = torch.zeros(n, dtype=torch.float64, device="cuda")
    dg = torch.zeros(n, dtype=torch.uint8, device="cuda")
    oc = np.zeros((n, _capi.OBS_BYTES), dtype=np.uint8)
    rc = np.zeros(n)
    dc = np.zeros(n, dtype=np.uint8)
    gpu.reset(og, max_episode_steps=700)
    cpu.reset(oc, max_episode_steps=700)
    gen = torch.Generator().manual_seed(0)
    actions = torch.randint(0, 8, (steps, n), generator=gen, dtype=torch.uint8)  # SURVEY.md 8d action generator
    rsum = np.zeros(n)
    for s in range(steps):
        a = actions[s].numpy()
        gpu.step(actions[s].cuda(), og, rg, dg)
        cpu.step(a, oc, rc, dc)
        r = rg.cpu().numpy()
        assert np.array_equal(r, rc), f"reward differs at step {s}"
        rsum += np.abs(r)
        if s % 50 == 49 or s == steps - 1:
            assert np.array_equal(og.cpu().numpy(), oc), f"obs differs at step {s}"
            assert np.array_equal(dg.cpu().numpy(), dc)
        if dc.any():  # the vectoriser resets an env on the call after done
            gpu.reset(og, mask=dc, max_episode_steps=700)
            cpu.reset(oc, mask=dc, max_episode_steps=700)
        if s % 250 == 249 or s == steps - 1:
            for e in (0, 17, 35, 71):
                assert hashlib.sha256(gpu.save_state(e)).digest() == hashlib.sha256(cpu.save_state(e)).digest(), (s, e)
    assert (rsum > 0).all()
    assert gpu.counters().faults == 0
    _assert_same_states(gpu, cpu, range(n), "end of long run")


@pytest.mark.parametrize("n,steps", [(32768, 3), (4096, 120)])
def test_full_size_batch_matches_replicated_small_batch(cuda_lib, oracle_lib, roms, n, steps):
    """BASELINE.json's batch sizes (32,768 envs: the per-GPU target; 4,096 envs: the benchmark configuration, 2 envs per
    warp, run for 120 steps across an episode boundary) through a size-independent property: env i is given the start
    state and the actions of env (i mod 64), so every output must equal the 64-env oracle run replicated.
    Exercises tile/lane indexing, the envs-per-warp heuristic, the visited-map slot budget and the sparse heat maps."""
    import torch

    base = 64
    rom = roms("pokelike")
    gpu = _capi.Handle(cuda_lib, n, rom, 0)
    cpu = _capi.Handle(oracle_lib, base, rom)
    gpu.tick(30, False)
    cpu.tick(30, False)
    og = torch.zeros((n, _capi.OBS_BYTES), dtype=torch.uint8, device="cuda")
    rg = torch.zeros(n, dtype=torch.float64, device="cuda")
    dg = torch.zeros(n, dtype=torch.uint8, device="cuda")
    oc = np.zeros((base, _capi.OBS_BYTES), dtype=np.uint8)
    rc = np.zeros(base)
    dc = np.zeros(base, dtype=np.uint8)
    gpu.reset(og, max_episode_steps=50)
    cpu.reset(oc, max_episode_steps=50)
    rng = np.random.default_rng(77)
    reps = n // base
    for s in range(steps):
        act = rng.integers(0, 8, base).astype(np.uint8)
        gpu.step(torch.from_numpy(np.tile(act, reps)).cuda(), og, rg, dg)
        cpu.step(act, oc, rc, dc)
        assert np.array_equal(rg.cpu().numpy().reshape(reps, base), np.broadcast_to(rc, (reps, base))), f"reward differs at step {s}"
        assert np.array_equal(dg.cpu().numpy().reshape(reps, base), np.broadcast_to(dc, (reps, base)))
        o = og.view(reps, base, _capi.OBS_BYTES)
        ref = torch.from_numpy(oc).cuda()
        assert bool((o == ref[None]).all()), f"obs differs at step {s}"
        if dc.all() and s % 50 == 49:  # episode end: reset everything, as a vectoriser would
            gpu.reset(og, max_episode_steps=50)
            cpu.reset(oc, max_episode_steps=50)
    for e in (0, 63, 64, n // 3, n // 2 + 5, n - 1):
        assert gpu.save_state(e) == cpu.save_state(e % base), e
    info = torch.zeros(_capi.INFO_SCALARS, dtype=torch.float64, device="cuda")
    gpu.reduce_info(info)
    ic = np.zeros(_capi.INFO_SCALARS)
    cpu.reduce_info(ic)
    ig = info.cpu().numpy()
    assert np.allclose(ig, ic * reps, rtol=1e-12, atol=0), np.nonzero(ig != ic * reps)
    # 32,768 envs keep their heat maps as per-env hashes of the touched cells (dense image rebuilt on request), 4,096 dense
    for e in (1, n // 2 + 1, n - 2):
        assert np.array_equal(gpu.counts_map(e), cpu.counts_map(e % base)), e
    assert gpu.counters().faults == 0


def test_paged_exploration_storage_many_maps(cuda_lib, oracle_lib, roms, monkeypatch):
    """Visited bitmaps and heat maps are paged out of shared pools (gb_wrap.cuh): envs that walk through 40+ maps, far more
    than an equal share of a small pool would allow, still match the oracle's observation (channel 3 = visited window,
    environment.py:256-274), reward (exploration term, :1344-1345, :1375), info rows and heat maps; pages return to the pool
    on reset.  A heat-map pool that runs dry is reported as an error."""
    import torch

    n, steps = 48, 56
    rom = roms("pokelike")
    monkeypatch.setenv("GBENV_VISITED_PAGES", "400")  # 48 envs x 992 entries would be 47,616: 400 is < 9 pages per env
    gpu = _capi.Handle(cuda_lib, n, rom, 0)
    monkeypatch.setenv("GBENV_COUNTS_BLOCKS", "50")
    tiny = _capi.Handle(cuda_lib, n, rom, 0)
    cpu = _capi.Handle(oracle_lib, n, rom)
    og = torch.zeros((n, _capi.OBS_BYTES), dtype=torch.uint8, device="cuda")
    ot = torch.zeros_like(og)
    rg = torch.zeros(n, dtype=torch.float64, device="cuda")
    dg = torch.zeros(n, dtype=torch.uint8, device="cuda")
    oc, rc, dc = np.zeros((n, _capi.OBS_BYTES), np.uint8), np.zeros(n), np.zeros(n, np.uint8)
    for h in (gpu, tiny):
        h.tick(40, True)
        h.reset(og)
    cpu.tick(40, True)
    cpu.reset(oc)
    rng = np.random.default_rng(3)
    walkers = (1, 17, 40)  # these envs are teleported through a new map (and rows in all four 64-row bands) every step
    tiny_failed = False
    for s in range(steps):
        if s == 30:  # episode boundary for half of the envs: their pages go back to the pool and are handed out again
            mask = (np.arange(n) % 2 == 1).astype(np.uint8)
            gpu.reset(og, mask=mask)
            cpu.reset(oc, mask=mask)
            assert np.array_equal(og.cpu().numpy(), oc), "obs after masked reset"
        for k, e in enumerate(walkers):
            m, r, c = (7 * s + 31 * k) % 248, (37 * s + 5 * k) % 250, (11 * s + 3 * k) % 250
            for h in (gpu, cpu) + (() if tiny_failed else (tiny,)):
                h.write_mem(e, 0xD35E, [m])
                h.write_mem(e, 0xD361, [r])
                h.write_mem(e, 0xD362, [c])
        act = rng.integers(0, 8, n).astype(np.uint8)
        a = torch.from_numpy(act).cuda()
        gpu.step(a, og, rg, dg)
        cpu.step(act, oc, rc, dc)
        if not tiny_failed:
            try:
                tiny.step(a, ot, rg.clone(), dg.clone())
            except _capi.GbEnvError as e:
                assert "heat-map block pool" in str(e)
                tiny_failed = True
        assert np.array_equal(og.cpu().numpy(), oc), f"obs step {s}"
        assert np.array_equal(rg.cpu().numpy(), rc), f"reward step {s}"
    gpu.check()
    ig = torch.zeros((n, _capi.INFO_SCALARS), dtype=torch.float64, device="cuda")
    ic = np.zeros((n, _capi.INFO_SCALARS))
    gpu.get_info(ig)
    cpu.get_info(ic)
    assert np.array_equal(ig.cpu().numpy(), ic)
    for e in (0, 1, 17, n - 1):
        assert np.array_equal(gpu.counts_map(e), cpu.counts_map(e)), e
    assert gpu.counters().faults == 0
    if not tiny_failed:
        with pytest.raises(_capi.GbEnvError, match="heat-map block pool"):
            tiny.check()


def test_deferred_ppu_matches_inline_and_oracle(cuda_lib, oracle_lib, roms, monkeypatch):
    """Deferred PPU (k_run_frames records the lines of the rendered frame, k_render_pending draws them; a VRAM / OAM write
    in between flushes first) against the same library with GBENV_DEFER=0 (lines drawn inside the emulation kernel) and
    against the oracle, on the LCD-observation ROM (VRAM / OAM / scroll / LCDC / LY writes in the middle of frames), through
    gbenv_step_masked: an env that is skipped keeps its framebuffer -- its stale pending range must not be drawn again."""
    import torch

    n, steps = 40, 24
    rom = roms("lcd_probe")
    monkeypatch.setenv("GBENV_DEFER", "1")
    deferred = _capi.Handle(cuda_lib, n, rom, 0)
    monkeypatch.setenv("GBENV_DEFER", "0")
    inline = _capi.Handle(cuda_lib, n, rom, 0)
    monkeypatch.delenv("GBENV_DEFER")
    cpu = _capi.Handle(oracle_lib, n, rom)
    hs = (deferred, inline, cpu)
    for h in hs:
        h.tick(5, True)  # every frame rendered: the lines of frame k are flushed at the start of frame k + 1
    dev = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    bufs = []
    for h in hs:
        gpu = h is not cpu
        o, r, d = np.zeros((n, _capi.OBS_BYTES), np.uint8), np.zeros(n), np.zeros(n, np.uint8)
        bufs.append((dev(o), dev(r), dev(d)) if gpu else (o, r, d))
        h.reset(bufs[-1][0])
    rng = np.random.default_rng(77)
    for s in range(steps):
        act = rng.integers(0, 8, n).astype(np.uint8)
        skip = (rng.random(n) < 0.2).astype(np.uint8) if s % 3 == 1 else np.zeros(n, np.uint8)
        for h, (o, r, d) in zip(hs, bufs):
            gpu = h is not cpu
            h.step_masked(dev(act) if gpu else act, dev(skip) if gpu else skip, o, r, d)
        torch.cuda.synchronize()
        for k in (0, 1):
            assert np.array_equal(bufs[k][0].cpu().numpy(), bufs[2][0]), (s, k)
            assert np.array_equal(bufs[k][1].cpu().numpy(), bufs[2][1]), (s, k)
        if s % 6 == 5:
            for e in range(n):
                a, b, c = deferred.save_state(e), inline.save_state(e), cpu.save_state(e)
                assert a == c, f"step {s} env {e}: deferred vs oracle {diff_states(c, a)[:6]}"
                assert b == c, f"step {s} env {e}: inline vs oracle {diff_states(c, b)[:6]}"
    for h in hs:
        h.close()
