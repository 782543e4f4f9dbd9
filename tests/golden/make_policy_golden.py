"""Golden outputs of the reference's ActorCritic (architectures/mlp.py), produced by EXECUTING that file unmodified
on the NumPy stand-in for flax.linen / distrax (tests/golden/jaxshim.py).  Run from the repo root, in the container
that has /root/reference:   python -m tests.golden.make_policy_golden

Each P_<name>.npz holds: the observation rows (uint8 [M, K], agent-major as baselines/utils.py:48-61 batchifies them),
the parameter seed for oracle.policy_oracle.make_params (the parameters themselves are regenerated, not stored),
a float64 checksum of the parameters, and the reference's `logits` [M, A] and `value` [M] for tanh and relu."""
import json
import os

import numpy as np

from oracle import policy_oracle as po
from tests.golden import jaxshim
from tests.helpers import GOLDEN_DIR, load_golden


def obs_rows(case, steps):
    g, _, _ = load_golden(case)
    o0, o1 = g["s_obs0"][steps], g["s_obs1"][steps]                 # [T, N, h, w, 26]
    rows = np.concatenate([o0.reshape(-1, o0[0, 0].size), o1.reshape(-1, o1[0, 0].size)], axis=0)   # agent-major
    return np.ascontiguousarray(rows)


def checksum(params):
    return float(sum(np.asarray(v[k], np.float64).sum() for v in params["params"].values() for k in ("kernel", "bias")))


def main():
    ns = jaxshim.load_reference()
    rng = np.random.default_rng(7)
    cases = {
        "cramped_room": obs_rows("C_cramped_room_400", np.r_[0:40, 355:405]),          # incl. urgency + auto-reset frames
        "padded_5x9": obs_rows("E_padded_coord_ring", np.r_[0:48]),
        "dense_random": rng.integers(0, 4, (96, 520), dtype=np.uint8) * (rng.random((96, 520)) < 0.3).astype(np.uint8),
        "byte_range": rng.integers(0, 256, (32, 208), dtype=np.uint8),
    }
    for name, rows in cases.items():
        K = rows.shape[1]
        seed = 1000 + K
        params = po.make_params(seed, K)
        out = {"obs": rows, "meta": json.dumps({"seed": seed, "obs_dim": K, "action_dim": 6, "checksum": checksum(params)})}
        for act in ("tanh", "relu"):
            pi, value = ns.ActorCritic(6, activation=act).apply(params, rows)
            out["logits_" + act] = np.asarray(pi.logits, np.float32)
            out["value_" + act] = np.asarray(value, np.float32)
            lo, vo = po.forward(params, rows, act)
            assert np.array_equal(lo, out["logits_" + act]) and np.array_equal(vo, out["value_" + act]), name
        np.savez_compressed(os.path.join(GOLDEN_DIR, "P_" + name + ".npz"), **out)
        print(name, rows.shape, "nonzero/row %.1f" % (rows != 0).sum(1).mean())


if __name__ == "__main__":
    main()
