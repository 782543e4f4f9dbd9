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
