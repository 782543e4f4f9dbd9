"""CPU-side checks: the C-ABI library loads and exports every symbol include/oc_b200.h declares; host-side layout
logic (registry, grid parser, padding) matches the reference-derived expectations.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    hdr = open(os.path.join(ROOT, "include", "oc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(oc_[a-z_0-9]+)\s*\(", hdr))
    assert {"oc_reset", "oc_step", "oc_get_obs", "oc_layout_create", "oc_prng_split"} <= declared
    lib = ctypes.CDLL(os.path.join(ROOT, "jaxovercooked_b200", "liboc_b200.so"))
    for name in declared:
        assert hasattr(lib, name), f"liboc_b200.so does not export {name}"
    lib.oc_version.restype = ctypes.c_int
    assert lib.oc_version() == 120
    lib.oc_status_string.restype = ctypes.c_char_p
    assert lib.oc_status_string(-2) == b"unsupported layout"


def test_binding_covers_header():
    from jaxovercooked_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "oc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(oc_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    lib = _lib.load()
    assert lib.oc_version() == 120


def test_no_cpu_fallback_without_cuda():
    import torch
    import jaxovercooked_b200 as ocb
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(Exception):
        ocb.make("overcooked", layout=ocb.overcooked_layouts["cramped_room"])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jaxovercooked_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"


def test_registry_and_padding():
    from jaxovercooked_b200.layouts import overcooked_layouts, cl_sequence_padded, layout_grid_to_dict
    assert len(overcooked_layouts) == 38
    cr = overcooked_layouts["cramped_room"]
    assert (cr["height"], cr["width"]) == (4, 5) and cr["agent_idx"].tolist() == [6, 8] and cr["pot_idx"].tolist() == [2]
    assert overcooked_layouts["forced_coord"]["agent_idx"].tolist() == [11, 8]      # order matters: agent_0 first
    # SURVEY.md App. C.2 (derived from baselines/IPPO_CL.py:182-278)
    exp = {"cramped_room": ([12, 14], [4], [32], [11, 15], [30]),
           "asymm_advantages": ([29, 32], [22, 31], [12, 17], [9, 14], [39, 41]),
           "coord_ring": ([13, 21], [5, 15], [40], [29, 39], [20]),
           "forced_coord": ([21, 14], [5, 15], [41], [11, 20], [29]),
           "counter_circuit": ([11, 33], [3, 4], [25], [39, 40], [18])}
    for name, lay in cl_sequence_padded().items():
        assert (lay["height"], lay["width"]) == (5, 9)
        got = tuple(lay[k].tolist() for k in ("agent_idx", "pot_idx", "goal_idx", "onion_pile_idx", "plate_pile_idx"))
        assert got == exp[name], name
        walls = set(lay["wall_idx"].tolist())
        border = [i for i in range(45) if i // 9 in (0, 4) or i % 9 in (0, 8)]
        assert all(b in walls for b in border)
    g = layout_grid_to_dict("\nWWPWW\nOA AO\nW   W\nWBWXW\n")
    for k in ("agent_idx", "goal_idx", "plate_pile_idx", "onion_pile_idx", "pot_idx"):
        assert g[k].tolist() == cr[k].tolist()
    assert sorted(set(g["wall_idx"].tolist())) == sorted(cr["wall_idx"].tolist())


def test_all_registered_layouts_satisfy_the_abi_contract():
    """oc_layout_create's preconditions (enclosed grid, specials on walls, two free agent cells) hold for the registry."""
    from jaxovercooked_b200.layouts import overcooked_layouts
    for name, l in overcooked_layouts.items():
        h, w = l["height"], l["width"]
        walls = set(l["wall_idx"].tolist())
        assert h <= 16 and w <= 16 and len(l["pot_idx"]) >= 1
        for i in range(h * w):
            if i // w in (0, h - 1) or i % w in (0, w - 1):
                assert i in walls, (name, i)
        for k in ("goal_idx", "pot_idx", "onion_pile_idx", "plate_pile_idx"):
            assert all(int(i) in walls for i in l[k]), (name, k)
        assert len(l["agent_idx"]) == 2 and all(int(i) not in walls for i in l["agent_idx"])


def test_actor_critic_host_mirror_without_a_gpu():
    """Parameter pytree, initialiser scales and argument checks of the policy mirror (architectures/mlp.py:26-53); the
    forward itself needs the GPU and must refuse anything else instead of falling back."""
    import numpy as np
    import torch
    from jaxovercooked_b200.policy import ActorCritic, LAYERS

    net = ActorCritic(6, activation="something-else")          # mlp.py:20-23: anything but "relu" means tanh
    assert net.activation == "tanh"
    params = net.init(3, 520, device="cpu")
    assert tuple(params["params"]) == LAYERS
    shapes = {n: tuple(params["params"][n]["kernel"].shape) for n in LAYERS}
    assert shapes == {"Dense_0": (520, 128), "Dense_1": (128, 128), "Dense_2": (128, 6),
                      "Dense_3": (520, 128), "Dense_4": (128, 128), "Dense_5": (128, 1)}
    for name, scale in (("Dense_0", 2 ** 0.5), ("Dense_1", 2 ** 0.5), ("Dense_2", 0.01), ("Dense_5", 1.0)):
        k = params["params"][name]["kernel"].double()
        gram = k.T @ k if k.shape[0] >= k.shape[1] else k @ k.T
        np.testing.assert_allclose(gram.numpy(), scale * scale * np.eye(gram.shape[0]), atol=1e-5)
        assert float(params["params"][name]["bias"].abs().max()) == 0.0
    with pytest.raises(ValueError):                              # CPU tensors are not silently computed on the host
        net.apply(params, torch.zeros((4, 520), dtype=torch.uint8))


def test_public_header_is_plain_c():
    """The drop-in boundary is a C ABI: the header must compile as C99 with no torch / CUDA / C++ types in sight."""
    import subprocess
    import tempfile
    src = ('#include "oc_b200.h"\nint main(void){ oc_policy_desc d; oc_step_extras e; oc_layout_info i; oc_state s;'
           ' (void)d; (void)e; (void)i; (void)s; return oc_version() ? 0 : 1; }\n')
    with tempfile.NamedTemporaryFile("w", suffix=".c", delete=False) as f:
        f.write(src)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-fsyntax-only", f.name], capture_output=True, text=True)
    os.unlink(f.name)
    assert r.returncode == 0, r.stderr


def test_xla_ffi_shim_compiles_against_the_stub():
    """csrc/xla_ffi_shim.cc cannot meet the real xla/ffi/api/ffi.h here (no JAX in the image).  Compile-only check against a
    stand-in header that static_asserts every handler's parameter list against its binding, position by position, and
    type-checks every call into include/oc_b200.h; and every FFI target jax_ffi.py registers is a symbol the shim defines."""
    import re
    import subprocess
    shim = os.path.join(ROOT, "jaxovercooked_b200", "csrc", "xla_ffi_shim.cc")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "c_abi", "xla_ffi_stub"),
                        "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", shim], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", open(shim).read()))
    src = open(os.path.join(ROOT, "jaxovercooked_b200", "jax_ffi.py")).read()
    registered = set(re.findall(r'\("oc_\w+", "(Oc\w+)"\)', src))
    assert registered and registered <= defined, (registered, defined)
    called = set(re.findall(r'ffi_call\(\s*"(oc_\w+)"', src)) | {"oc_step_reset_state"}
    names = set(re.findall(r'\("(oc_\w+)", "Oc\w+"\)', src))
    assert called <= names, (called, names)


def test_unpack_results8_numpy():
    """The packed result byte of oc_rollout (reward/20 | class(shaped0) << 2 | class(shaped1) << 4 | done << 6), decoded on the host."""
    import numpy as np
    from jaxovercooked_b200.env import Overcooked
    codes = np.array([0, 1, 2, 1 << 2, 2 << 2, 1 << 4, 2 << 4, 1 << 6, 2 | (2 << 2) | (1 << 4) | (1 << 6)], np.uint8)
    r, s0, s1, d = Overcooked.unpack_results8(codes)
    assert r.tolist() == [0, 20, 40, 0, 0, 0, 0, 0, 40] and r.dtype == np.float32
    assert s0.tolist() == [0, 0, 0, 3, 5, 0, 0, 0, 5] and s1.tolist() == [0, 0, 0, 0, 0, 3, 5, 0, 3]
    assert d.tolist() == [False] * 7 + [True, True]
