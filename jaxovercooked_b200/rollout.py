"""Rollout buffer in the learners' layout.

The reference's trainers `batchify` the per-agent observations into an agent-major `(2N, h*w*26)` matrix every
step and stack `num_steps` of them into `Transition.obs` (baselines/utils.py:48-61, baselines/IPPO_CL.py:510,
556-563) -- one extra full copy of the observations per step.  Here the step kernel writes straight into the
trajectory tensor: `obs[t, agent]` is handed to `oc_step` as that step's `obs0` / `obs1`, so `obs[t].view(2N, -1)`
IS `batchify(last_obs)` (actor index = agent * N + env, row = C-order (y, x, c)) with no copy at all."""
import torch


class RolloutBuffer:
    def __init__(self, env, num_steps, num_envs):
        dev, h, w = env.device, env.height, env.width
        self.T, self.N = int(num_steps), int(num_envs)
        self.obs = torch.empty((self.T, 2, self.N, h, w, 26), dtype=torch.uint8, device=dev)
        self.reward = torch.empty((self.T, self.N), dtype=torch.float32, device=dev)
        self.shaped = torch.empty((self.T, 2, self.N), dtype=torch.float32, device=dev)
        self.done = torch.empty((self.T, self.N), dtype=torch.bool, device=dev)

    def slot(self, t):
        """`out=` argument of `env.step` for rollout step t."""
        return {"obs0": self.obs[t, 0], "obs1": self.obs[t, 1], "reward": self.reward[t],
                "shaped0": self.shaped[t, 0], "shaped1": self.shaped[t, 1], "done": self.done[t]}

    def batchified_obs(self, t=None):
        """(T, 2N, h*w*26) -- or (2N, h*w*26) for one step -- agent-major, exactly `batchify`'s layout."""
        if t is None:
            return self.obs.view(self.T, 2 * self.N, -1)
        return self.obs[t].view(2 * self.N, -1)


class PolicyRollout:
    """The rollout phase of the reference's trainers -- the `_env_step` body scanned `num_steps` times
    (baselines/IPPO_MLP.py:423-493: split rng, `pi, value = network.apply(params, obs_batch)`, `action = pi.sample`,
    `log_prob = pi.log_prob(action)`, split rng, `split(_rng, num_envs)`, `vmap(env.step)`, collect the Transition) --
    as ONE replayable CUDA graph of 2 * num_steps + 1 kernels: per step `oc_policy_forward_cells` (actions for both
    agents from the env kernel's compact observations) and `oc_step` (writes the next observations straight into the
    trajectory tensor and refreshes the compact observations in place), plus one `oc_prng_chain` for the 2 * num_steps
    subkeys of the `rng, _rng = split(rng)` chain.

    Trajectory tensors (all on the device, reused by every `run`; T = num_steps, N = num_envs, actors are agent-major as
    `batchify` orders them: row a * N + e):
        obs      uint8 [T + 1, 2, N, h, w, 26]   obs[t] is what step t's actions were computed from (`Transition.obs[t]` =
                                                 obs[t].view(2N, -1)); obs[T] is `last_obs` for the bootstrap value
        action   int32 [T, 2N]    log_prob / value  float32 [T, 2N]
        reward   float32 [T, N]   (both agents get it, overcooked.py:124)      shaped  float32 [T, 2, N]
        done     bool [T, N]      (== dones["__all__"] == each agent's done)
    """

    def __init__(self, env, network, params, state, rng, num_steps, *, log=None, stats=None, graph=True, shard=None):
        """`shard=(first, n_global)`: `state` holds envs [first, first + N) of a global batch of n_global environments (one GPU's
        slice of the env axis); per-env keys and the categorical draws are then exactly those of the single-device rollout over
        the whole batch (`split(key, n_global)[first + i]`, `oc_policy_set_shard`)."""
        from . import random as jr
        self.env, self.net, self.params, self.state = env, network, params, state
        self.T, self.N = int(num_steps), int(state.time.shape[0])
        T, N, dev, h, w = self.T, self.N, env.device, env.height, env.width
        self.rng = rng.clone()                                            # advanced in place by every run()
        self._impl = "threefry_legacy" if env.prng_mode == 0 else "threefry_partitionable"
        self.obs = torch.empty((T + 1, 2, N, h, w, 26), dtype=torch.uint8, device=dev)
        self.cells = env.empty_cells(N)
        self.action = torch.empty((T, 2 * N), dtype=torch.int32, device=dev)
        self.log_prob = torch.empty((T, 2 * N), dtype=torch.float32, device=dev)
        self.value = torch.empty((T, 2 * N), dtype=torch.float32, device=dev)
        self.reward = torch.empty((T, N), dtype=torch.float32, device=dev)
        self.shaped = torch.empty((T, 2, N), dtype=torch.float32, device=dev)
        self.done = torch.empty((T, N), dtype=torch.bool, device=dev)
        self.subkeys = torch.empty((2 * T, 2), dtype=torch.uint32, device=dev)
        self._log, self._stats = log, stats
        self._jr = jr
        self._first, self._n_global = (int(shard[0]), int(shard[1])) if shard is not None else (0, self.N)
        network.bind(params, env=env, shard=shard)
        first = env.get_obs(state, cells=self.cells)
        self.obs[0, 0].copy_(first["agent_0"])
        self.obs[0, 1].copy_(first["agent_1"])
        self._fresh = True
        self._graph = None
        if graph:
            # capture on a side stream; nothing is executed by the capture itself
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    self._body()
            torch.cuda.current_stream(dev).wait_stream(side)
            self._graph = g

    def _body(self):
        env, net, jr, T, N = self.env, self.net, self._jr, self.T, self.N
        jr.split_chain(self.rng, 2 * T, impl=self._impl, out=self.subkeys)      # rng, _rng = split(rng), twice per step
        for t in range(T):
            out = {"action": self.action[t], "log_prob": self.log_prob[t], "value": self.value[t]}
            net.act_cells(self.params, self.cells, self.subkeys[2 * t], out=out)
            step_out = {"obs0": self.obs[t + 1, 0], "obs1": self.obs[t + 1, 1], "reward": self.reward[t],
                        "shaped0": self.shaped[t, 0], "shaped1": self.shaped[t, 1], "done": self.done[t], "cells": self.cells}
            key = jr.LazySplit(self.subkeys[2 * t + 1], self._n_global, self._first, N)   # split(_rng, num_envs), derived in the kernel
            acts = {"agent_0": self.action[t, :N], "agent_1": self.action[t, N:]}
            env.step(key, self.state, acts, donate=True, out=step_out, log=self._log, stats=self._stats)

    def run(self):
        """One rollout of `num_steps` steps from where the previous one stopped.  Returns self (the tensors above)."""
        if not self._fresh:                # continue: the last observations of the previous rollout are the first of this one
            self.obs[0].copy_(self.obs[self.T])
        self._fresh = False
        self.net._handle(self.params)      # repacks the parameters first if an optimiser step changed them
        if self._graph is not None:
            self._graph.replay()
        else:
            self._body()
        return self

    # views in the shapes the reference's Transition uses ([T, num_actors, ...], num_actors = 2N agent-major)
    def transition_obs(self):
        return self.obs[: self.T].view(self.T, 2 * self.N, -1)

    def last_obs(self):
        return self.obs[self.T].view(2 * self.N, -1)

    def transition_reward(self, shaping_weight=0.0):
        """batchify(reward + shaped_reward * anneal) (baselines/IPPO_MLP.py:455-466): float32 [T, 2N]."""
        r = self.reward[:, None, :] + float(shaping_weight) * self.shaped
        return r.reshape(self.T, 2 * self.N)

    def transition_done(self):
        return self.done[:, None, :].expand(self.T, 2, self.N).reshape(self.T, 2 * self.N)
