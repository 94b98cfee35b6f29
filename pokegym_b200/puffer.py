"""PufferLib-shaped vector adaptor: async_reset / recv / send over VecEnvironment.

The reference is driven by a PufferLib fork that is neither in the reference tree nor installed here, so
this contract is UNPINNED: observations are flattened to uint8[N, 23040], `recv` returns
(obs, rewards, terminals, truncations, infos, env_ids, mask) and an env is reset on the call after `done`.
"""
from __future__ import annotations

import numpy as np

from . import _capi
from .info import info_row_to_dict
from .vec_env import VecEnvironment


class PufferVecAdaptor:
    def __init__(self, vec: VecEnvironment):
        self.vec = vec
        self.num_envs = vec.num_envs
        self.env_ids = np.arange(self.num_envs)
        self._pending_reset = np.zeros(self.num_envs, dtype=np.uint8)
        self._last = None

    def async_reset(self, seed=None):
        obs, _ = self.vec.reset()
        t = self.vec.torch
        self._last = (obs, t.zeros(self.num_envs, dtype=t.float64, device=self.vec.device), t.zeros(self.num_envs, dtype=t.bool, device=self.vec.device), [])

    def send(self, actions):
        if self._pending_reset.any():  # envs that reported done on the previous recv restart now
            self.vec.reset(mask=self._pending_reset)
        obs, rew, done, _, _ = self.vec.step(actions)
        d = done.cpu().numpy()
        infos = []
        if d.any():
            rows = self.vec.info().cpu().numpy()
            infos = [info_row_to_dict(rows[e]) for e in np.nonzero(d)[0]]
        self._pending_reset = d.astype(np.uint8)
        self._last = (obs, rew, done, infos)

    def recv(self):
        obs, rew, done, infos = self._last
        flat = obs.reshape(self.num_envs, _capi.OBS_BYTES)
        mask = np.ones(self.num_envs, dtype=bool)
        return flat, rew, done, done, infos, self.env_ids, mask
