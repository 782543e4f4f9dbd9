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


    def rollout(self, keys, state, actions, *, out=None, stats=None):
        """T LogWrapper steps in ONE launch (`env.rollout`, oc_rollout): the scan of `_env_step` over the wrapped env
        (baselines/IPPO_CL.py:520-560) with the T steps' actions given.  `state` (LogEnvState) is updated IN PLACE; `info`
        carries `returned_episode` per step (bool[T, N, 2]) and the trackers after the last step."""
        n = int(state.env_state.time.shape[0])
        T = int(keys.shape[0]) if not isinstance(keys, tuple) else int(keys[0].shape[0])
        trackers = {f: getattr(state, f) for f in ("episode_returns", "episode_lengths", "returned_episode_returns",
                                                    "returned_episode_lengths")}
        returned_episode = torch.empty((T, n, self._env.num_agents), dtype=torch.bool, device=self._env.device)
        log = dict(trackers, returned_episode=returned_episode)
        obs, _, reward, done, info = self._env.rollout(keys, state.env_state, actions, out=out, log=log, stats=stats)
        if self.replace_info:
            info = {}
        info["returned_episode_returns"] = state.returned_episode_returns
        info["returned_episode_lengths"] = state.returned_episode_lengths
        info["returned_episode"] = returned_episode
        return obs, state, reward, done, info


class CTRolloutManager(JaxMARLWrapper):
    """Rollout manager of the Q-learning baselines (reference: jax_marl/wrappers/baselines.py:153-249) for the way
    the VDN scripts use it -- `CTRolloutManager(LogWrapper(make(...)), batch_size=num_envs, preprocess_obs=False)`
    (baselines/vdn_mlp.py:339-346): `batch_reset(key)` / `batch_step(key, states, actions)` split `key` into
    `batch_size` per-env keys (fused into the kernel on the step side), and add the global state
    `obs["__all__"]` = the two agents' flattened observations side by side (:196) and `reward["__all__"]` = agent_0's
    reward (:197).  `preprocess_obs=True` (pad + one-hot agent id, :237-249) is not on the Overcooked VDN path and is
    not provided."""

    def __init__(self, env, batch_size, preprocess_obs=False):
        super().__init__(env)
        if preprocess_obs:
            raise NotImplementedError("preprocess_obs=True is not used by the Overcooked VDN baselines")
        self.batch_size = int(batch_size)
        self.training_agents = self.agents
        self.preprocess_obs = False
        inner = env
        while hasattr(inner, "_env"):
            inner = inner._env
        self._inner = inner
        self._impl = "threefry_partitionable" if inner.prng_mode else "threefry_legacy"

    def _obs_from_all(self, obs_all):
        """obs dict over ONE uint8[N, 2, h, w, 26] array written by the kernel: the per-agent observations are its [:, a]
        slices and the global state (wrappers/baselines.py:196) is the same memory flattened per env -- no concatenation."""
        return {"agent_0": obs_all[:, 0], "agent_1": obs_all[:, 1], "__all__": obs_all.view(self.batch_size, -1)}

    def batch_reset(self, key):
        from . import random as jr
        _, state = self._env.reset(jr.split(key, self.batch_size, impl=self._impl))
        inner_state = state.env_state if hasattr(state, "env_state") else state
        return self._obs_from_all(self._inner.get_obs_all(inner_state)), state

    def batch_step(self, key, states, actions, **kw):
        from . import random as jr
        inner = self._inner
        out = dict(kw.pop("out", None) or {})
        n, dev = self.batch_size, inner.device
        if out.get("obs_all") is None:
            out["obs_all"] = torch.empty((n, 2, inner.height, inner.width, 26), dtype=torch.uint8, device=dev)
        for k, dt in (("reward", torch.float32), ("shaped0", torch.float32), ("shaped1", torch.float32), ("done", torch.bool)):
            if out.get(k) is None:
                out[k] = torch.empty((n,), dtype=dt, device=dev)
        obs, states, reward, done, infos = self._env.step(jr.LazySplit(key, self.batch_size), states, actions, out=out, **kw)
        reward = dict(reward)
        reward["__all__"] = reward[self.training_agents[0]]                    # wrappers/baselines.py:197,229
        return obs, states, reward, done, infos
