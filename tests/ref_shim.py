"""Run the UNMODIFIED reference wrapper (/root/reference/pokegym/environment.py) on a PyBoy-shaped shim.

PyBoy, gymnasium, skimage, matplotlib and mediapy are absent from this image, so the six imports the
reference makes are stubbed and `pyboy.PyBoy` is replaced by an adaptor over the CPU oracle's emulator
core (oracle/gb_core.c through liboracle.so).  Everything above PyBoy -- Environment.reset/step, ram_map,
ram_map_leanke, game_map, the ram_reader Game API -- is the reference's own code, loaded from
/root/reference at test time (never copied).  This pins the WRAPPER semantics (reward, observation,
RAM side effects, done) of oracle/pokegym_wrapper.c and of the CUDA kernels to the reference.

Only usable in the build container (needs /root/reference); GPU-box tests use the golden vectors this
harness produced (tests/golden/).
"""
from __future__ import annotations

import contextlib
import io
import copy
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

REFERENCE_ROOT = Path("/root/reference")

# PyBoy 1.6 WindowEvent numbering (pyboy/utils.py)
_EVENTS = ["QUIT", "PRESS_ARROW_UP", "PRESS_ARROW_DOWN", "PRESS_ARROW_RIGHT", "PRESS_ARROW_LEFT", "PRESS_BUTTON_A", "PRESS_BUTTON_B",
           "PRESS_BUTTON_SELECT", "PRESS_BUTTON_START", "RELEASE_ARROW_UP", "RELEASE_ARROW_DOWN", "RELEASE_ARROW_RIGHT", "RELEASE_ARROW_LEFT",
           "RELEASE_BUTTON_A", "RELEASE_BUTTON_B", "RELEASE_BUTTON_SELECT", "RELEASE_BUTTON_START"]
# event -> (joypad button id of include/gbenv.h: Right Left Up Down A B Select Start, pressed)
_EVENT_TO_BUTTON = {"ARROW_RIGHT": 0, "ARROW_LEFT": 1, "ARROW_UP": 2, "ARROW_DOWN": 3, "BUTTON_A": 4, "BUTTON_B": 5, "BUTTON_SELECT": 6, "BUTTON_START": 7}


class WindowEvent:
    pass


for _i, _n in enumerate(_EVENTS):
    setattr(WindowEvent, _n, _i)


class ShimConfig:
    """Set before constructing the reference Environment: which ROM / oracle library the fake PyBoy uses."""

    rom: bytes = b""
    oracle_lib = None
    last_pyboy = None


class _Screen:
    def __init__(self, pyboy):
        self._p = pyboy

    def raw_screen_buffer_dims(self):
        return (144, 160)

    def screen_ndarray(self):
        return self._p.handle.screen(0)


class _BotSupport:
    def __init__(self, pyboy):
        self._p = pyboy

    def screen(self):
        return _Screen(self._p)


class FakePyBoy:
    """The 11-method PyBoy surface pokegym uses, over one oracle env."""

    def __init__(self, gamerom_file, **kwargs):
        from pokegym_b200 import _capi

        self.handle = _capi.Handle(ShimConfig.oracle_lib, 1, ShimConfig.rom)
        self.events = []
        self.rendering = True
        self.n_ticks = 0
        self.mem_writes = []  # (addr, value) issued by the wrapper through set_memory_value
        ShimConfig.last_pyboy = self

    def botsupport_manager(self):
        return _BotSupport(self)

    def set_emulation_speed(self, v):
        pass

    def load_state(self, f):
        blob = f.read()
        tid = self.handle.add_state_template(blob)
        self.handle.load_template(tid)

    def save_state(self, f):
        f.write(self.handle.save_state(0))

    def get_memory_value(self, addr):
        return int(self.handle.read_mem(0, addr, 1)[0])

    def set_memory_value(self, addr, value):
        self.mem_writes.append((addr, value))
        self.handle.write_mem(0, addr, [value & 0xFF])

    def send_input(self, event):
        self.events.append(event)

    def _rendering(self, value):
        self.rendering = bool(value)

    def tick(self):
        for ev in self.events:  # PyBoy.tick applies queued inputs first
            name = _EVENTS[ev]
            pressed = name.startswith("PRESS_")
            key = name.split("_", 1)[1]
            self.handle.send_input(_EVENT_TO_BUTTON[key], pressed)
        self.events = []
        self.handle.tick(1, self.rendering)
        self.n_ticks += 1
        return False

    def stop(self, save=True):
        self.handle.close()


def _install_stubs():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    gym = mod("gymnasium")

    class Env:
        pass

    class Box:
        def __init__(self, low, high, dtype=None, shape=None):
            self.low, self.high, self.dtype, self.shape = low, high, dtype, shape

    class Discrete:
        def __init__(self, n):
            self.n = n

    spaces = mod("gymnasium.spaces")
    spaces.Box, spaces.Discrete = Box, Discrete
    gym.Env, gym.spaces = Env, spaces
    sk = mod("skimage")
    skt = mod("skimage.transform")
    skt.resize = lambda *a, **k: None
    sk.transform = skt
    mpl = mod("matplotlib")
    plt = mod("matplotlib.pyplot")
    plt.imsave = lambda *a, **k: None
    mpl.pyplot = plt
    mod("mediapy")
    pb = mod("pyboy")
    pb.PyBoy = FakePyBoy
    pb.WindowEvent = WindowEvent
    pbu = mod("pyboy.utils")
    pbu.WindowEvent = WindowEvent
    pb.utils = pbu


_ENV_CLS = None
_WORKDIR = None


def reference_environment_class():
    """Import pokegym.Environment from /root/reference with the stubs in place (once per process)."""
    global _ENV_CLS, _WORKDIR
    if _ENV_CLS is not None:
        return _ENV_CLS
    if not REFERENCE_ROOT.exists():
        raise RuntimeError("reference tree not available")
    _install_stubs()
    _WORKDIR = tempfile.mkdtemp(prefix="refshim_")  # the reference creates experiments/, videos/, csv/ in cwd
    os.chdir(_WORKDIR)
    sys.path.insert(0, str(REFERENCE_ROOT))
    with contextlib.redirect_stdout(io.StringIO()):
        import pokegym  # noqa: F401  (spawns the reference's multiprocessing.Manager)
        from pokegym.environment import Environment
    _ENV_CLS = Environment
    return Environment


def make_reference_env(rom: bytes, oracle_lib, state_path: str):
    ShimConfig.rom = rom
    ShimConfig.oracle_lib = oracle_lib
    Environment = reference_environment_class()
    with contextlib.redirect_stdout(io.StringIO()):
        env = Environment(rom_path="unused.gb", state_path=state_path, headless=True, quiet=True)
    return env, ShimConfig.last_pyboy


def run_reference_episode(rom: bytes, oracle_lib, state_blob: bytes, actions, max_episode_steps=20480, reward_scale=4.0, n_resets=1,
                          reset_at=()):
    """Returns dict(rewards, dones, obs list, states list, mem_writes per step) from the reference wrapper."""
    with tempfile.NamedTemporaryFile(suffix=".state", delete=False) as f:
        f.write(state_blob)
        path = f.name
    env, pyboy = make_reference_env(rom, oracle_lib, path)
    out = dict(rewards=[], dones=[], obs=[], states=[], writes=[])
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        obs, _ = env.reset(max_episode_steps=max_episode_steps, reward_scale=reward_scale)
    out["reset_obs"] = [np.array(obs, copy=True)]
    out["reset_state"] = [pyboy.handle.save_state(0)]
    for i, a in enumerate(actions):
        if i in reset_at:
            with contextlib.redirect_stdout(sink):
                obs, _ = env.reset(max_episode_steps=max_episode_steps, reward_scale=reward_scale)
            out["reset_obs"].append(np.array(obs, copy=True))
            out["reset_state"].append(pyboy.handle.save_state(0))
        pyboy.mem_writes = []
        with contextlib.redirect_stdout(sink):
            obs, rew, done, trunc, info = env.step(int(a))
        assert done == trunc
        out["rewards"].append(float(rew))
        out["dones"].append(bool(done))
        out["obs"].append(np.array(obs, copy=True))
        out["states"].append(pyboy.handle.save_state(0))
        out["writes"].append(list(pyboy.mem_writes))
        out.setdefault("infos", []).append(copy.deepcopy(info) if info else info)  # counts_map is mutated by later steps
    out["env"] = env
    os.unlink(path)
    return out
