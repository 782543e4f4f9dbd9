"""TEST INFRASTRUCTURE ONLY -- NumPy float32 restatement of the reference's actor-critic MLP forward and of the
action sampling around it.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

Follows (reference file:line)
    architectures/mlp.py:11-59      ActorCritic.__call__: actor Dense(128)-act-Dense(128)-act-Dense(action_dim),
                                    critic Dense(128)-act-Dense(128)-act-Dense(1), value = squeeze(critic, -1);
                                    flax auto-names the six layers Dense_0..Dense_5 in that order; kernels are
                                    [in, out], y = x @ kernel + bias; uint8 observations are promoted to float32
    baselines/utils.py:48-61        batchify: rows are agent-major, each row the (y, x, channel)-flattened obs
    baselines/IPPO_MLP.py:435-439   pi, value = network.apply(params, obs_batch); action = pi.sample(seed=rng);
                                    log_prob = pi.log_prob(action)
`pi` is `distrax.Categorical(logits=...)` (third-party, not under /root/reference): sample(seed) draws
`jax.random.categorical(key, logits, axis=-1, shape=(1, M))[0]` = argmax(logits + gumbel(key, (1, M, A))), and
log_prob(a) = log_softmax(logits)[a].

PARITY STATUS: the wiring of ActorCritic is pinned by executing the reference's own mlp.py on the NumPy stand-in
for flax/distrax (tests/golden/make_policy_golden.py -> tests/golden/P_*.npz).  The arithmetic is floating point:
the CUDA path is compared within a stated tolerance, never bit-exact; the categorical draw is "parity unpinned"
against a real JAX (see oracle/prng.py), and only defined up to near-ties of logits + gumbel.
"""
import numpy as np

from . import prng

LAYERS = ("Dense_0", "Dense_1", "Dense_2", "Dense_3", "Dense_4", "Dense_5")


def orthogonal(rng, shape, scale):
    """A column-orthonormal [in, out] matrix times `scale` (the shape of flax's orthogonal init; NOT its draws)."""
    n_in, n_out = shape
    a = rng.standard_normal((max(n_in, n_out), min(n_in, n_out)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if n_in < n_out:
        q = q.T
    return (scale * q[:n_in, :n_out]).astype(np.float32)


def make_params(seed, obs_dim, action_dim=6, hidden=128, bias_scale=0.1):
    """Deterministic parameters in the reference's pytree layout ({'params': {'Dense_i': {'kernel','bias'}}}).
    Same scales as mlp.py's initialisers, but with non-zero biases so that a mis-wired bias is caught."""
    rng = np.random.default_rng(seed)
    s2 = float(np.sqrt(2.0))
    shapes = [((obs_dim, hidden), s2), ((hidden, hidden), s2), ((hidden, action_dim), 0.01),
              ((obs_dim, hidden), s2), ((hidden, hidden), s2), ((hidden, 1), 1.0)]
    out = {}
    for name, (shape, scale) in zip(LAYERS, shapes):
        out[name] = {"kernel": orthogonal(rng, shape, scale),
                     "bias": (bias_scale * rng.standard_normal(shape[1])).astype(np.float32)}
    return {"params": out}


def _act(x, activation):
    return np.maximum(x, np.float32(0)) if activation == "relu" else np.tanh(x)


def forward(params, obs, activation="tanh"):
    """obs: uint8 or float [M, obs_dim] -> (logits float32[M, A], value float32[M])."""
    p = params["params"]
    x = np.asarray(obs, dtype=np.float32).reshape(len(obs), -1)

    def dense(h, name):
        return (h @ np.asarray(p[name]["kernel"], np.float32) + np.asarray(p[name]["bias"], np.float32)).astype(np.float32)

    a = dense(_act(dense(_act(dense(x, "Dense_0"), activation), "Dense_1"), activation), "Dense_2")
    c = dense(_act(dense(_act(dense(x, "Dense_3"), activation), "Dense_4"), activation), "Dense_5")
    return a, c[:, 0]


def gumbel_noise(key, m, a, partitionable=False):
    """gumbel(key, (1, m, a)) as float32[m, a]."""
    return prng.gumbel_f32(np.asarray(key, np.uint32), m * a, partitionable).reshape(m, a)


def sample(key, logits, partitionable=False):
    """Categorical(logits).sample(seed=key) -> (action int32[M], perturbed float32[M, A])."""
    z = (np.asarray(logits, np.float32) + gumbel_noise(key, logits.shape[0], logits.shape[1], partitionable)).astype(np.float32)
    return np.argmax(z, axis=-1).astype(np.int32), z


def log_prob(logits, action):
    l = np.asarray(logits, np.float32)
    m = l.max(axis=-1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(l - m).sum(axis=-1, dtype=np.float32)).astype(np.float32)
    return (l[np.arange(len(l)), action] - lse).astype(np.float32)
