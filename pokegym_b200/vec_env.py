"""Drop-in Python boundary: `Environment` (pokegym's single-env Gymnasium API) and `VecEnvironment`
(N envs on one GPU, tensors in / tensors out), both over the CUDA library through the C ABI.

Reference interface mirrored here (/root/reference/pokegym/environment.py):
  Environment(rom_path, state_path, headless, save_video, quiet, verbose, **kw)            :437-446
  reset(seed=None, options=None, max_episode_steps=20480, reward_scale=4.0) -> (obs, {})   :1233,1334
  step(action, fast_video=True) -> (obs, reward, done, done, info)                         :1336,1812
  render() -> obs, close(), observation_space Box(0,255,(72,80,4),uint8), action_space Discrete(8)  :154-167,256,412

There is no CPU fallback: constructing either class needs libgbenv.so and a CUDA device.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Iterable, Optional, Sequence, Union

import numpy as np

from . import _capi
from .info import INFO_NAMES, build_info, info_row_to_dict  # noqa: F401

try:  # gymnasium is not installed in this image; use it when it is
    from gymnasium.spaces import Box, Discrete  # type: ignore
except Exception:  # pragma: no cover - exercised implicitly

    class Box:  # minimal stand-in with the attributes vectorisers read
        def __init__(self, low, high, shape=None, dtype=np.uint8):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

        def sample(self):
            return np.random.randint(self.low, self.high + 1, size=self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and x.dtype == self.dtype

    class Discrete:
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.dtype(np.int64)

        def sample(self):
            return int(np.random.randint(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n


OBS_SHAPE = (_capi.OBS_H, _capi.OBS_W, _capi.OBS_C)
_LIB = None


def _lib() -> _capi.GbEnvLib:
    global _LIB
    if _LIB is None:
        _LIB = _capi.GbEnvLib(_capi.DEFAULT_LIB)  # raises GbEnvError when the CUDA library is missing
    return _LIB


def _default_state() -> Optional[Path]:
    """The save-state pokegym.Environment falls back to when state_path is None (environment.py:119-120 loads the packaged
    current_state/Bulbasaur.state): $POKEGYM_STATE, else that file inside an installed / checked-out pokegym package."""
    env = os.environ.get("POKEGYM_STATE")
    if env and Path(env).exists():
        return Path(env)
    try:
        import importlib.util

        spec = importlib.util.find_spec("pokegym")
        if spec and spec.origin:
            p = Path(spec.origin).parent / "current_state" / "Bulbasaur.state"
            if p.exists():
                return p
    except Exception:
        pass
    p = Path("/root/reference/pokegym/current_state/Bulbasaur.state")
    return p if p.exists() else None


def _read_rom(rom_path: Union[str, os.PathLike, bytes]) -> bytes:
    if isinstance(rom_path, (bytes, bytearray)):
        return bytes(rom_path)
    p = Path(rom_path)
    if p.exists():
        return p.read_bytes()
    env = os.environ.get("POKEGYM_ROM")
    if env and Path(env).exists():
        return Path(env).read_bytes()
    raise FileNotFoundError(
        f"ROM {rom_path!r} not found (pokegym expects a user-supplied pokemon_red.gb; set POKEGYM_ROM, or pass "
        f"pokegym_b200.tools.synth_rom.build_pokelike_rom() for the synthetic test ROM)"
    )


class VecEnvironment:
    """N Game Boy envs on one GPU.  All tensors live on `device`; nothing returns to the host per step.

    obs      uint8 [N, 72, 80, 4]  (a view into `rollout[t]` when a rollout tensor is supplied)
    reward   float64 [N]           (the reference returns Python floats)
    done     bool [N]              (terminated == truncated in the reference)
    """

    def __init__(self, num_envs: int, rom_path, state_paths: Union[None, str, bytes, Sequence] = None, device="cuda:0", rollout=None,
                 auto_reset: bool = False, max_episode_steps: int = 20480, reward_scale: float = 4.0, boot_frames: int = 60,
                 validate_actions: bool = False):
        import torch

        self.torch = torch
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _capi.GbEnvError("VecEnvironment runs on CUDA devices only (there is no CPU fallback)")
        self.num_envs = int(num_envs)
        self.handle = _capi.Handle(_lib(), self.num_envs, _read_rom(rom_path), device_id=self.device.index or 0)
        self.observation_space = Box(low=0, high=255, shape=OBS_SHAPE, dtype=np.uint8)
        self.action_space = Discrete(_capi.NUM_ACTIONS)
        self.max_episode_steps, self.reward_scale, self.auto_reset = int(max_episode_steps), float(reward_scale), bool(auto_reset)
        self.rollout = rollout
        self.validate_actions = bool(validate_actions)  # off by default: the check is a device->host sync; the kernel masks actions with & 7
        self._t = 0
        with torch.cuda.device(self.device):
            self._obs = torch.zeros((self.num_envs, *OBS_SHAPE), dtype=torch.uint8, device=self.device)
            self._reward = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
            self._done = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
            self._done_prev = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
            self._info = torch.zeros((self.num_envs, _capi.INFO_SCALARS), dtype=torch.float64, device=self.device)
        if state_paths is None:
            # no save-state: boot the ROM for a few frames so every env sits in its main loop
            self.handle.tick(boot_frames, True, stream=self._stream())
        else:
            blobs = [state_paths] if isinstance(state_paths, (str, bytes, os.PathLike)) else list(state_paths)
            tids = [self.handle.add_state_template(b if isinstance(b, bytes) else Path(b).read_bytes()) for b in blobs]
            if len(tids) == 1:
                self.handle.set_initial_template(tids[0])
            else:  # env i starts from state i mod len(states) (BASELINE.json config 5)
                for k, tid in enumerate(tids):
                    ids = np.arange(k, self.num_envs, len(tids), dtype=np.int32)
                    if ids.size:
                        self.handle.set_initial_template(tid, ids)

    # -- helpers -------------------------------------------------------------
    def _stream(self) -> int:
        return int(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _obs_target(self):
        if self.rollout is None:
            return self._obs
        r = self.rollout
        ok = (self.torch.is_tensor(r) and r.dtype == self.torch.uint8 and r.device == self.device and r.is_contiguous() and r.dim() >= 3
              and r.shape[1] == self.num_envs and r[0, 0].numel() == _capi.OBS_BYTES)
        if not ok:  # raw pointers cross the C ABI: a wrong layout would be out-of-bounds device writes, not an exception
            raise ValueError(f"rollout must be a contiguous uint8 tensor [T, {self.num_envs}, 72, 80, 4] (or [T, N, 23040]) on {self.device}")
        return r[self._t % r.shape[0]]

    # -- API -----------------------------------------------------------------
    def reset(self, mask=None, max_episode_steps: Optional[int] = None, reward_scale: Optional[float] = None):
        """Environment.reset for every env (or those with mask[e] != 0).  Returns (obs, {})."""
        if max_episode_steps is not None:
            self.max_episode_steps = int(max_episode_steps)
        if reward_scale is not None:
            self.reward_scale = float(reward_scale)
        obs = self._obs_target()
        torch = self.torch
        if mask is not None and not torch.is_tensor(mask):
            mask = torch.as_tensor(np.ascontiguousarray(mask, dtype=np.uint8), device=self.device)
        if mask is not None:  # a device mask (e.g. the done vector): the reset stays on the device, no host round trip
            if mask.dtype == torch.bool:
                mask = mask.to(torch.uint8)
            if mask.dtype != torch.uint8 or mask.device != self.device or not mask.is_contiguous() or mask.numel() != self.num_envs:
                mask = mask.to(device=self.device, dtype=torch.uint8).contiguous().view(-1)
                if mask.numel() != self.num_envs:
                    raise ValueError(f"reset mask must have {self.num_envs} entries")
        self.handle.reset_dev(obs, mask_dev=mask, max_episode_steps=self.max_episode_steps, reward_scale=self.reward_scale,
                              obs_stride=_capi.OBS_BYTES, stream=self._stream())
        return obs.view(self.num_envs, *OBS_SHAPE), {}

    def step(self, actions, skip=None):
        """Environment.step for every env.  `actions`: uint8/int tensor [N] on the device (or array-like).
        `skip` (uint8 device tensor [N], optional): envs with skip[e] != 0 sit this step out -- reward 0, done False, their
        observation row keeps what it holds (e.g. the reset observation just written by reset(mask=...))."""
        torch = self.torch
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions), device=self.device)
        if actions.dtype != torch.uint8 or actions.device != self.device or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.uint8).contiguous()
        if actions.numel() != self.num_envs:
            raise ValueError(f"expected {self.num_envs} actions, got {actions.numel()}")
        if self.validate_actions and bool((actions >= _capi.NUM_ACTIONS).any()):  # the reference raises IndexError (ACTIONS[action])
            raise IndexError("action out of range 0..7")
        self._t += 1
        obs = self._obs_target()
        if skip is None:
            self.handle.step(actions, obs, self._reward, self._done, obs_stride=_capi.OBS_BYTES, stream=self._stream())
        else:
            if skip.dtype != torch.uint8 or skip.device != self.device or not skip.is_contiguous() or skip.numel() != self.num_envs:
                raise ValueError(f"skip must be a contiguous uint8 tensor with {self.num_envs} entries on {self.device}")
            if self.rollout is not None:  # a new rollout slot: the rows of the envs that sit out must carry over
                prev = self.rollout[(self._t - 1) % self.rollout.shape[0]]
                obs.view(self.num_envs, -1)[skip.bool()] = prev.view(self.num_envs, -1)[skip.bool()]
            self.handle.step_masked(actions, skip, obs, self._reward, self._done, obs_stride=_capi.OBS_BYTES, stream=self._stream())
        done = self._done.bool()
        if self.auto_reset:
            # envs that just finished are reset on the device, masked by the done vector the step wrote: their row of `obs`
            # becomes the reset observation, as a vectoriser that resets after `done` would deliver it (no host round trip)
            self._done_prev.copy_(self._done)
            self.handle.reset_dev(obs, mask_dev=self._done_prev, max_episode_steps=self.max_episode_steps, reward_scale=self.reward_scale,
                                  obs_stride=_capi.OBS_BYTES, stream=self._stream())
        return obs.view(self.num_envs, *OBS_SHAPE), self._reward, done, done, {}

    @staticmethod
    def compact_view(obs):
        """Optional 1+1-channel view for new policies: the reference's channels 0-2 are the same grey level,
        so `obs[..., 2:]` = (grey, visited) carries everything.  A strided view of the same memory, no copy."""
        return obs[..., 2:]

    def info(self):
        """Per-env info rows: float64 [N, 72] (GBENV_INFO_SCALARS); column names in pokegym_b200.info.INFO_NAMES."""
        self.handle.get_info(self._info, stream=self._stream())
        return self._info

    def info_sum(self, out=None):
        """Sum of the info rows over this GPU's envs (the vector multi-GPU runs all-reduce)."""
        torch = self.torch
        if out is None:
            out = torch.zeros(_capi.INFO_SCALARS, dtype=torch.float64, device=self.device)
        self.handle.reduce_info(out, stream=self._stream())
        return out

    def full_info(self, env: int = 0) -> dict:
        """The reference's complete info dict for one env (environment.py:1621-1810): scalar row + the 130 named
        event flags (read from the env's WRAM) + its counts_map.  Rare path (episode end / every 10,000 steps)."""
        row = self.info()[env].cpu().numpy()
        wram = self.handle.read_mem(env, 0xD700, 0x200)  # every event flag lives in 0xD7B1..0xD838
        try:
            cm = self.handle.counts_map(env).astype(np.float64)
        except _capi.GbEnvError:  # heat maps switched off (GBENV_COUNTS_MAP=0)
            cm = None
        return build_info(row, lambda a: int(wram[a - 0xD700]), counts_map=cm, reward_scale=self.reward_scale)

    def save_state(self, env: int = 0) -> bytes:
        return self.handle.save_state(env)

    def load_state(self, blob: bytes, env_ids=None):
        """PyBoy.load_state for the listed envs (None: all): pyboy_binding.load_pyboy_state (:65-69)."""
        self.handle.load_template(self.handle.add_state_template(blob), env_ids)

    def close(self):
        self.handle.close()


class Environment:
    """pokegym.Environment on the GPU path: a batch of one (the CUDA library is used even for N = 1)."""

    def __init__(self, rom_path="pokemon_red.gb", state_path=None, headless=True, save_video=False, quiet=False, verbose=False, device="cuda:0", **kwargs):
        if save_video or not headless:
            raise NotImplementedError("video / SDL2 window output is out of scope (SURVEY.md section 8b non-goals)")
        if state_path is None and not isinstance(rom_path, (bytes, bytearray)):
            # environment.py:119-120: the reference falls back to its packaged current_state/Bulbasaur.state.  With a ROM
            # given as bytes (the synthetic test ROMs) there is no such convention and the ROM is booted instead.
            state_path = _default_state()
            if state_path is None:
                raise FileNotFoundError("no state_path given and pokegym's default current_state/Bulbasaur.state was not found "
                                        "(set POKEGYM_STATE or pass state_path)")
        self.vec = VecEnvironment(1, rom_path, state_paths=state_path, device=device)
        self.observation_space = self.vec.observation_space
        self.action_space = self.vec.action_space
        self.verbose = verbose
        self.time = 0
        self.max_episode_steps = 20480
        self._actions = self.vec.torch.zeros(1, dtype=self.vec.torch.uint8, device=self.vec.device)

    def reset(self, seed=None, options=None, max_episode_steps=20480, reward_scale=4.0):
        """Resets the game. Seeding is NOT supported (as in the reference)."""
        self.time = 0
        self.max_episode_steps = max_episode_steps
        obs, _ = self.vec.reset(max_episode_steps=max_episode_steps, reward_scale=reward_scale)
        self._last_obs = obs[0].cpu().numpy()
        return self._last_obs, {}

    def step(self, action, fast_video=True):
        self._actions.fill_(int(action))
        obs, reward, done, _, _ = self.vec.step(self._actions)
        self.time += 1
        d = bool(done[0].item())
        info = {}
        if d or self.time % 10000 == 0:  # environment.py:1621
            info = self.vec.full_info(0)
        self._last_obs = obs[0].cpu().numpy()
        return self._last_obs, float(reward[0].item()), d, d, info

    def render(self):
        """The observation the last reset()/step() returned (environment.py:256-274 rebuilds the same array)."""
        return self._last_obs

    # environment.py:208-227: save_state / load_first_state / load_last_state / load_random_state over self.initial_states
    def save_state(self):
        if not hasattr(self, "initial_states"):
            self.initial_states = []
        self.initial_states.append(self.vec.save_state(0))

    def load_first_state(self):
        return self.initial_states[0]

    def load_last_state(self):
        return self.initial_states[-1]

    def load_random_state(self):
        import random

        return self.initial_states[random.randrange(len(self.initial_states))]

    def close(self):
        self.vec.close()
