"""CPU: the actor-critic oracle against the fixtures produced by executing the reference's architectures/mlp.py."""
import json
import os

import numpy as np
import pytest

from oracle import policy_oracle as po
from oracle import prng
from tests.helpers import GOLDEN_DIR, policy_cases


def load_policy_case(case):
    g = np.load(os.path.join(GOLDEN_DIR, case + ".npz"))
    meta = json.loads(str(g["meta"]))
    params = po.make_params(meta["seed"], meta["obs_dim"], meta["action_dim"])
    chk = float(sum(np.asarray(v[k], np.float64).sum() for v in params["params"].values() for k in ("kernel", "bias")))
    assert abs(chk - meta["checksum"]) < 1e-9, "make_params no longer reproduces the fixture's parameters"
    return g, meta, params


@pytest.mark.parametrize("case", policy_cases())
@pytest.mark.parametrize("act", ["tanh", "relu"])
def test_oracle_matches_reference_mlp(case, act):
    g, meta, params = load_policy_case(case)
    logits, value = po.forward(params, g["obs"], act)
    assert logits.dtype == np.float32 and value.dtype == np.float32
    np.testing.assert_array_equal(logits, g["logits_" + act])
    np.testing.assert_array_equal(value, g["value_" + act])


def test_there_are_policy_fixtures():
    assert len(policy_cases()) >= 4


def test_orthogonal_init_shape_and_scale():
    rng = np.random.default_rng(0)
    for shape, scale in (((520, 128), 2 ** 0.5), ((128, 6), 0.01), ((128, 1), 1.0), ((64, 128), 1.0)):
        w = po.orthogonal(rng, shape, scale)
        assert w.shape == shape and w.dtype == np.float32
        gram = (w.T @ w) if shape[0] >= shape[1] else (w @ w.T)
        np.testing.assert_allclose(gram, scale * scale * np.eye(min(shape)), atol=1e-5)


@pytest.mark.parametrize("part", [False, True])
def test_sample_and_log_prob(part):
    rng = np.random.default_rng(3)
    logits = rng.standard_normal((257, 6)).astype(np.float32)
    key = prng.PRNGKey(11)
    a, z = po.sample(key, logits, part)
    assert a.dtype == np.int32 and a.min() >= 0 and a.max() <= 5
    # the perturbation is gumbel(key, (1, M, A)) in flat row-major order
    g = prng.gumbel_f32(key, 257 * 6, part).reshape(257, 6)
    np.testing.assert_array_equal(z, logits + g)
    # empirical frequencies follow softmax(logits) for a constant-logit batch
    l1 = np.tile(np.array([[0.0, 1.0, -1.0, 0.5, 0.0, -2.0]], np.float32), (20000, 1))
    a1, _ = po.sample(prng.PRNGKey(5), l1, part)
    freq = np.bincount(a1, minlength=6) / len(a1)
    p = np.exp(l1[0]) / np.exp(l1[0]).sum()
    assert np.abs(freq - p).max() < 0.015
    lp = po.log_prob(logits, a)
    ref = logits[np.arange(257), a] - np.log(np.exp(logits.astype(np.float64)).sum(-1))
    np.testing.assert_allclose(lp, ref, atol=2e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree is not mounted")
def test_live_reference_mlp_on_fresh_inputs():
    from tests.golden import jaxshim
    ns = jaxshim.load_reference()
    rng = np.random.default_rng(99)
    for K, act in ((520, "tanh"), (1170, "relu"), (260, "anything-else-means-tanh")):
        params = po.make_params(int(rng.integers(1 << 30)), K)
        obs = (rng.integers(0, 24, (33, K)) * (rng.random((33, K)) < 0.05)).astype(np.uint8)
        pi, value = ns.ActorCritic(6, activation=act).apply(params, obs)
        lo, vo = po.forward(params, obs, act)
        np.testing.assert_array_equal(lo, np.asarray(pi.logits))
        np.testing.assert_array_equal(vo, np.asarray(value))
