"""The C ABI called from plain C (tests/c_abi/harness.c): no Python, no ctypes mirror between the caller and liboc_b200.so.
The harness replays a golden fixture (recorded from the reference's own source) through oc_layout_create -> oc_reset ->
T x oc_step, then through one oc_rollout, and dumps everything; this test byte-compares the dump with the fixture."""
import os
import subprocess

import numpy as np
import pytest

from tests.helpers import DYN_FIELDS, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "c_abi", "harness")
LAYOUT_LISTS = ("wall_idx", "agent_idx", "goal_idx", "pot_idx", "onion_pile_idx", "plate_pile_idx")

CASES = ["A_cramped_room", "A_asymm_advantages", "A_counter_circuit", "A_easy_layout_padded", "A_bottleneck_large"]


def _fixture_bytes(g, meta, lay):
    E, T = g["reset_keys"].shape[0], g["actions"].shape[0]
    n = {k: int(np.asarray(lay[k]).size) for k in LAYOUT_LISTS}
    hdr = np.array([0x0C0C0C0C, lay["height"], lay["width"], n["wall_idx"], n["goal_idx"], n["pot_idx"], n["onion_pile_idx"],
                    n["plate_pile_idx"], int(meta["random_reset"]), meta["max_steps"], meta["task_id"],
                    int(meta["partitionable"]), E, T], dtype=np.uint32).view(np.int32)
    parts = [hdr] + [np.asarray(lay[k], np.int32).reshape(-1) for k in LAYOUT_LISTS]
    parts += [g["reset_keys"].astype(np.uint32), g["step_keys"].astype(np.uint32), g["actions"].astype(np.int32)]
    return b"".join(np.ascontiguousarray(p).tobytes() for p in parts)


def _expected(g):
    T = g["actions"].shape[0]
    st = lambda pre, t=None: [np.ascontiguousarray(g[pre + f] if t is None else g[pre + f][t]).tobytes() for f in DYN_FIELDS]
    a = [g["r_obs0"].tobytes(), g["r_obs1"].tobytes()] + st("r_")
    per_step = lambda t: [g["s_obs0"][t].tobytes(), g["s_obs1"][t].tobytes(), g["s_reward"][t].astype(np.float32).tobytes(),
                          g["s_shaped0"][t].astype(np.float32).tobytes(), g["s_shaped1"][t].astype(np.float32).tobytes(),
                          g["s_done"][t].astype(np.uint8).tobytes()]
    for t in range(T):
        a += per_step(t) + st("s_", t)
    b = []
    for t in range(T):
        b += per_step(t)
    b += st("s_", T - 1) + [g["s_obs0"][T - 1].tobytes(), g["s_obs1"][T - 1].tobytes()]
    return b"".join(a), b"".join(b)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_plain_c_caller_reproduces_the_golden_trajectory(case, tmp_path):
    if not os.path.exists(HARNESS):          # normally built by __graft_entry__.build(); gcc is on every box
        import __graft_entry__ as ge
        ge._build_c_harness()
    assert os.path.exists(HARNESS), "tests/c_abi/harness is built by __graft_entry__.build()"
    g, meta, lay = load_golden(case)
    assert not meta["log_wrapper"] and not meta.get("crafted")
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    fin.write_bytes(_fixture_bytes(g, meta, lay))
    r = subprocess.run([HARNESS, str(fin), str(fout)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr + r.stdout
    got = fout.read_bytes()
    exp_a, exp_b = _expected(g)
    assert len(got) == len(exp_a) + len(exp_b), (len(got), len(exp_a), len(exp_b))
    assert got[:len(exp_a)] == exp_a, "oc_reset / oc_step from C differ from the reference's trajectory"
    assert got[len(exp_a):] == exp_b, "oc_rollout from C differs from the reference's trajectory"


def test_harness_source_is_plain_c99():
    """The harness (and therefore include/oc_b200.h) compiles as C99 with no CUDA compiler involved."""
    src = os.path.join(ROOT, "tests", "c_abi", "harness.c")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        "-I", "/usr/local/cuda/include", src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
