"""Host mirror of the reference's actor-critic MLP (architectures/mlp.py:11-59) on top of `oc_policy_*`.

    network = ActorCritic(env.action_space().n, activation="tanh")
    params  = network.init(rng, obs_dim)                      # same pytree as flax: {'params': {'Dense_0': {'kernel', 'bias'}, ...}}
    pi, value = network.apply(params, obs_batch)              # baselines/IPPO_MLP.py:435
    action    = pi.sample(seed=rng); log_prob = pi.log_prob(action)          # :438-439
    action, log_prob, value = network.act(params, obs_batch, rng)           # the same three calls as ONE kernel

`obs_batch` is the uint8 `[num_actors, h*w*26]` tensor `batchify` builds (baselines/utils.py:48-61); rows of a
`RolloutBuffer` slot are already in that layout.  Parameters are float32 torch CUDA tensors in flax's `[in, out]`
kernel layout, so a reference checkpoint loads without transposition.  The kernels repack them
(`oc_policy_update`) whenever a parameter tensor's version counter changes.  Floating point: results are within
a tolerance of the float32 reference (tests/test_gpu_policy.py), not bit-exact.  No CPU fallback.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib

LAYERS = ("Dense_0", "Dense_1", "Dense_2", "Dense_3", "Dense_4", "Dense_5")
HIDDEN = 128


class Categorical:
    """The part of `distrax.Categorical` the baselines use: `.logits`, `.sample(seed=)`, `.log_prob`, `.entropy`."""

    def __init__(self, logits, prng_mode=_lib.PRNG_LEGACY):
        self.logits = logits
        self.prng_mode = prng_mode

    def sample(self, seed):
        """argmax(logits + gumbel(seed, (1, M, A))) -- the draw of `jax.random.categorical`, threefry on the device."""
        logits = self.logits.contiguous()
        M, A = logits.shape
        key = seed if seed.dtype == torch.uint32 else seed.to(torch.uint32)
        action = torch.empty((M,), dtype=torch.int32, device=logits.device)
        with torch.cuda.device(logits.device):
            _lib.check(_lib.load().oc_policy_sample(self.prng_mode, M, A, logits.data_ptr(), key.data_ptr(), action.data_ptr(),
                                                    None, torch.cuda.current_stream().cuda_stream), "oc_policy_sample")
        return action

    def log_prob(self, action):
        return torch.log_softmax(self.logits, dim=-1).gather(-1, action.long().unsqueeze(-1)).squeeze(-1)

    def entropy(self):
        lp = torch.log_softmax(self.logits, dim=-1)
        return -(lp.exp() * lp).sum(-1)


class ActorCritic:
    def __init__(self, action_dim, activation="tanh", prng_impl="threefry_legacy"):
        if activation not in ("tanh", "relu"):
            activation = "tanh"                        # architectures/mlp.py:20-23: anything but "relu" means tanh
        self.action_dim = int(action_dim)
        self.activation = activation
        self.prng_impl = prng_impl
        self.prng_mode = _lib.PRNG_LEGACY if prng_impl == "threefry_legacy" else _lib.PRNG_PARTITIONABLE
        self._lib = _lib.load()
        self._bound = {}                               # id(params['params']) -> (handle, versions, keepalive)

    # -- parameters -------------------------------------------------------------------------------------------
    def init(self, rng, x, device="cuda"):
        """Parameters with the reference's initialisers (orthogonal(sqrt 2 | 0.01 | 1), zero bias; mlp.py:26-53).
        `x` is an example input or the flattened observation size.  The draws come from torch seeded with `rng`,
        not from jax.random: load a reference checkpoint for identical weights."""
        obs_dim = int(x) if isinstance(x, (int, np.integer)) else int(np.prod(x.shape[-1:]))
        seed = int(np.asarray(rng.cpu() if isinstance(rng, torch.Tensor) else rng).astype(np.uint64).sum()) if not isinstance(rng, int) else rng
        gen = torch.Generator().manual_seed(seed)
        shapes = [(obs_dim, HIDDEN, math.sqrt(2)), (HIDDEN, HIDDEN, math.sqrt(2)), (HIDDEN, self.action_dim, 0.01),
                  (obs_dim, HIDDEN, math.sqrt(2)), (HIDDEN, HIDDEN, math.sqrt(2)), (HIDDEN, 1, 1.0)]
        out = {}
        for name, (n_in, n_out, scale) in zip(LAYERS, shapes):
            a = torch.randn(max(n_in, n_out), min(n_in, n_out), generator=gen, dtype=torch.float64)
            q, r = torch.linalg.qr(a)
            q = q * torch.sign(torch.diagonal(r))
            if n_in < n_out:
                q = q.T
            out[name] = {"kernel": (scale * q[:n_in, :n_out]).to(torch.float32).contiguous().to(device),
                         "bias": torch.zeros(n_out, dtype=torch.float32, device=device)}
        return {"params": out}

    @staticmethod
    def params_from_numpy(params, device="cuda"):
        return {"params": {k: {n: torch.as_tensor(np.asarray(v[n], np.float32)).contiguous().to(device) for n in ("kernel", "bias")}
                           for k, v in params["params"].items()}}

    def _desc(self, p, obs_dim, template, layout=None):
        d = _lib.PolicyDesc()
        d.obs_dim, d.action_dim = obs_dim, self.action_dim
        d.activation = 1 if self.activation == "relu" else 0
        d.prng_mode = self.prng_mode
        for i, name in enumerate(LAYERS):
            k, b = p[name]["kernel"], p[name]["bias"]
            want = (obs_dim if i in (0, 3) else HIDDEN, (HIDDEN, HIDDEN, self.action_dim, HIDDEN, HIDDEN, 1)[i])
            if tuple(k.shape) != want or tuple(b.shape) != (want[1],):
                raise ValueError(f"{name}: kernel {tuple(k.shape)} / bias {tuple(b.shape)}, expected {want} / ({want[1]},)")
            if not (k.is_cuda and b.is_cuda and k.dtype == torch.float32 and b.dtype == torch.float32
                    and k.is_contiguous() and b.is_contiguous()):
                raise ValueError(f"{name}: parameters must be contiguous float32 CUDA tensors")
            d.kernel[i], d.bias[i] = k.data_ptr(), b.data_ptr()
        d.obs_template = template.ctypes.data if template is not None else None
        d.layout = layout
        return d

    def bind(self, params, obs_template=None, env=None, shard=None):
        """Pack `params` for the kernels (done implicitly by apply/act).  `obs_template` (uint8, obs_dim bytes) lets
        layer 1 be evaluated relative to an observation most rows are close to; `env` (an `Overcooked`) uses the env's own
        static image for that and additionally enables `act_cells` / `apply_cells` on the env's compact observations.
        `shard=(first, global_count)`: this process holds envs (act_cells) / rows (act) [first, first + n) of a global batch and
        draws exactly the actions the single-device program would draw for them (`oc_policy_set_shard`)."""
        p = params["params"]
        obs_dim = int(p["Dense_0"]["kernel"].shape[0])
        tmpl = None
        if env is not None:
            if env.height * env.width * 26 != obs_dim:
                raise ValueError(f"the env's observations have {env.height * env.width * 26} bytes, the network reads {obs_dim}")
        elif obs_template is not None:
            tmpl = np.ascontiguousarray(np.asarray(obs_template, np.uint8).reshape(-1))
            if tmpl.size != obs_dim:
                raise ValueError(f"obs_template has {tmpl.size} bytes, the network reads {obs_dim}")
        old = self._bound.pop(id(p), None)
        if old is not None:
            self._lib.oc_policy_destroy(old[0])
        d = self._desc(p, obs_dim, tmpl, env._handle if env is not None else None)
        h = C.c_void_p()
        with torch.cuda.device(p["Dense_0"]["kernel"].device):
            torch.cuda.current_stream().synchronize()          # the parameters may still be being written
            _lib.check(self._lib.oc_policy_create(C.byref(d), C.byref(h)), "oc_policy_create")
        if shard is not None:
            _lib.check(self._lib.oc_policy_set_shard(h, int(shard[0]), int(shard[1])), "oc_policy_set_shard")
        self._bound[id(p)] = [h, self._versions(p), (p, tmpl, env), obs_dim]
        return self

    @staticmethod
    def _versions(p):
        return tuple((p[n][f]._version, p[n][f].data_ptr()) for n in LAYERS for f in ("kernel", "bias"))

    def _handle(self, params):
        p = params["params"]
        ent = self._bound.get(id(p))
        if ent is None:
            self.bind(params)
            ent = self._bound[id(p)]
        v = self._versions(p)
        if v != ent[1]:                                          # an optimiser step happened: repack on this stream
            d = self._desc(p, ent[3], None, ent[2][2]._handle if ent[2][2] is not None else None)
            _lib.check(self._lib.oc_policy_update(ent[0], C.byref(d), torch.cuda.current_stream().cuda_stream), "oc_policy_update")
            ent[1] = v
        return ent[0]

    def __del__(self):
        for ent in getattr(self, "_bound", {}).values():
            try:
                self._lib.oc_policy_destroy(ent[0])
            except Exception:
                pass

    # -- forward ----------------------------------------------------------------------------------------------
    def _rows(self, x, obs_dim):
        if x.dtype != torch.uint8 or not x.is_cuda:
            raise ValueError("observations must be a uint8 CUDA tensor (the env's own output)")
        x = x.reshape(-1, obs_dim)
        return x if x.is_contiguous() else x.contiguous()

    def _forward(self, params, x, key, want_logits=True, out=None):
        h = self._handle(params)
        obs_dim = self._bound[id(params["params"])][3]
        x = self._rows(x, obs_dim)
        M, dev = x.shape[0], x.device
        out = out or {}
        logits = out.get("logits") if "logits" in out else (torch.empty((M, self.action_dim), dtype=torch.float32, device=dev) if want_logits else None)
        value = out.get("value") if "value" in out else torch.empty((M,), dtype=torch.float32, device=dev)
        action = log_prob = None
        kp = None
        if key is not None:
            key = key if key.dtype == torch.uint32 else key.to(torch.uint32)
            action = out.get("action") if "action" in out else torch.empty((M,), dtype=torch.int32, device=dev)
            log_prob = out.get("log_prob") if "log_prob" in out else torch.empty((M,), dtype=torch.float32, device=dev)
            kp = key.data_ptr()
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(dev):
            _lib.check(self._lib.oc_policy_forward(h, M, x.data_ptr(), ptr(logits), ptr(value), kp, ptr(action), ptr(log_prob),
                                                   torch.cuda.current_stream().cuda_stream), "oc_policy_forward")
        return logits, value, action, log_prob

    def _forward_cells(self, params, cells, key, want_logits=True, out=None):
        h = self._handle(params)
        if cells.dtype != torch.uint8 or not cells.is_cuda or cells.dim() != 2 or not cells.is_contiguous():
            raise ValueError("cells must be a contiguous uint8 CUDA tensor [N, cell_stride] (env.empty_cells / out['cells'])")
        N, dev = cells.shape[0], cells.device
        M = 2 * N
        out = out or {}
        logits = out.get("logits") if "logits" in out else (torch.empty((M, self.action_dim), dtype=torch.float32, device=dev) if want_logits else None)
        value = out.get("value") if "value" in out else torch.empty((M,), dtype=torch.float32, device=dev)
        action = log_prob = kp = None
        if key is not None:
            key = key if key.dtype == torch.uint32 else key.to(torch.uint32)
            action = out.get("action") if "action" in out else torch.empty((M,), dtype=torch.int32, device=dev)
            log_prob = out.get("log_prob") if "log_prob" in out else torch.empty((M,), dtype=torch.float32, device=dev)
            kp = key.data_ptr()
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(dev):
            _lib.check(self._lib.oc_policy_forward_cells(h, N, cells.data_ptr(), ptr(logits), ptr(value), kp, ptr(action),
                                                         ptr(log_prob), torch.cuda.current_stream().cuda_stream), "oc_policy_forward_cells")
        return logits, value, action, log_prob

    def apply_cells(self, params, cells):
        """`apply` on the env's compact observations: rows [0, N) agent_0, [N, 2N) agent_1 (the batchify order)."""
        logits, value, _, _ = self._forward_cells(params, cells, None)
        return Categorical(logits, self.prng_mode), value

    def act_cells(self, params, cells, rng, *, out=None, want_logits=False):
        """`act` on the env's compact observations (needs `bind(params, env=env)`); the 26-channel images are not read."""
        logits, value, action, log_prob = self._forward_cells(params, cells, rng, want_logits=want_logits, out=out)
        return (action, log_prob, value, logits) if want_logits else (action, log_prob, value)

    def apply(self, params, x):
        """`pi, value = network.apply(params, obs_batch)`."""
        logits, value, _, _ = self._forward(params, x, None)
        return Categorical(logits, self.prng_mode), value

    def act(self, params, x, rng, *, out=None, want_logits=False):
        """apply + pi.sample(seed=rng) + pi.log_prob(action) in one kernel -> (action, log_prob, value[, logits])."""
        logits, value, action, log_prob = self._forward(params, x, rng, want_logits=want_logits, out=out)
        return (action, log_prob, value, logits) if want_logits else (action, log_prob, value)
