"""LogWrapper mirror (reference: jax_marl/wrappers/baselines.py:28-107), with the bookkeeping fused into the
step kernel through `oc_step_extras` instead of a chain of elementwise ops."""
import dataclasses
from typing import Any

import torch

from .env import State


@dataclasses.dataclass
class LogEnvState:
    """wrappers/baselines.py:45-51; the four trackers are float32[..., num_agents] (jnp.zeros default dtype)."""
    env_state: State
    episode_returns: Any
    episode_lengths: Any
    returned_episode_returns: Any
    returned_episode_lengths: Any

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)


class JaxMARLWrapper:
    def __init__(self, env):
        self._env = env

    def __getattr__(self, name):
        return getattr(self._env, name)


class LogWrapper(JaxMARLWrapper):
    """Log the episode returns and lengths (agents terminate together)."""

    def __init__(self, env, replace_info=False):
        super().__init__(env)
        self.replace_info = replace_info

    def reset(self, key):
        obs, env_state = self._env.reset(key)
        batch = tuple(key.shape[:-1])
        z = lambda: torch.zeros(batch + (self._env.num_agents,), dtype=torch.float32, device=self._env.device)
        return obs, LogEnvState(env_state, z(), z(), z(), z())

    def step(self, key, state, action, *, donate=False, out=None, stats=None):
        batch = tuple(key.shape[:-1])      # works for raw keys [..., 2] and for random.LazySplit
        if donate:
            trackers = {f: getattr(state, f) for f in ("episode_returns", "episode_lengths",
                                                        "returned_episode_returns", "returned_episode_lengths")}
        else:
            trackers = {f: getattr(state, f).clone() for f in ("episode_returns", "episode_lengths",
                                                                "returned_episode_returns", "returned_episode_lengths")}
        returned_episode = torch.empty(batch + (self._env.num_agents,), dtype=torch.bool, device=self._env.device)
        log = dict(trackers, returned_episode=returned_episode)
        obs, env_state, reward, done, info = self._env.step(key, state.env_state, action, donate=donate, out=out,
                                                            log=log, stats=stats)
        new_state = LogEnvState(env_state, **trackers)
        if self.replace_info:
            info = {}
        info["returned_episode_returns"] = new_state.returned_episode_returns
        info["returned_episode_lengths"] = new_state.returned_episode_lengths
        info["returned_episode"] = returned_episode
        return obs, new_state, reward, done, info
