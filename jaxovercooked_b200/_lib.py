"""ctypes binding of liboc_b200.so (the C ABI declared in include/oc_b200.h).

There is NO CPU fallback: if the shared library is missing or fails to load, every entry point raises.
Build it with `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# OC_B200_LIB: alternative build of the same library (A/B tuning of kernel variants); never a different backend
LIB_PATH = os.environ.get("OC_B200_LIB") or os.path.join(_PKG, "liboc_b200.so")

OC_OK = 0
PRNG_LEGACY = 0
PRNG_PARTITIONABLE = 1

STATE_FIELDS = ("agent_pos", "agent_dir", "agent_dir_idx", "agent_inv", "goal_pos", "pot_pos", "wall_map",
                "maze_map", "time", "terminal", "task_id")
EXTRA_FIELDS = ("episode_returns", "episode_lengths", "returned_episode_returns", "returned_episode_lengths",
                "returned_episode", "stats")


class LayoutDesc(C.Structure):
    _fields_ = [("height", C.c_int32), ("width", C.c_int32),
                ("n_wall", C.c_int32), ("n_goal", C.c_int32), ("n_pot", C.c_int32),
                ("n_onion_pile", C.c_int32), ("n_plate_pile", C.c_int32),
                ("wall_idx", C.POINTER(C.c_int32)), ("agent_idx", C.POINTER(C.c_int32)),
                ("goal_idx", C.POINTER(C.c_int32)), ("pot_idx", C.POINTER(C.c_int32)),
                ("onion_pile_idx", C.POINTER(C.c_int32)), ("plate_pile_idx", C.POINTER(C.c_int32)),
                ("random_reset", C.c_int32), ("max_steps", C.c_int32), ("task_id", C.c_int32),
                ("prng_mode", C.c_int32)]


class StatePtrs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in STATE_FIELDS]


class StepExtras(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in EXTRA_FIELDS] + [("split_key", C.c_void_p), ("split_num", C.c_int64),
                                                          ("split_first", C.c_int64), ("cells", C.c_void_p),
                                                          ("obs_interleaved", C.c_int32)]


class LayoutInfo(C.Structure):
    _fields_ = [("height", C.c_int32), ("width", C.c_int32), ("n_goal", C.c_int32), ("n_pot", C.c_int32),
                ("map_bytes_per_env", C.c_int64), ("obs_bytes_per_agent", C.c_int64),
                ("algorithmic_bytes", C.c_int64),
                ("envs_per_warp", C.c_int32), ("warps_per_cta", C.c_int32), ("smem_bytes_per_cta", C.c_int32),
                ("n_scan", C.c_int32), ("cell_stride", C.c_int32), ("cell_agents", C.c_int32)]


class RolloutIO(C.Structure):
    _fields_ = [("act0", C.c_void_p), ("act1", C.c_void_p), ("actions8", C.c_void_p), ("obs0", C.c_void_p),
                ("obs1", C.c_void_p), ("obs_step_stride", C.c_int64), ("reward", C.c_void_p), ("shaped0", C.c_void_p),
                ("shaped1", C.c_void_p), ("done", C.c_void_p), ("results8", C.c_void_p), ("obs_interleaved", C.c_int32),
                ("actions8_scratch", C.c_void_p)]


WIRE_REFERENCE, WIRE_PACKED = 0, 1


class PolicyDesc(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("action_dim", C.c_int32), ("activation", C.c_int32), ("prng_mode", C.c_int32),
                ("kernel", C.c_void_p * 6), ("bias", C.c_void_p * 6), ("obs_template", C.c_void_p),
                ("layout", C.c_void_p)]


EXPORTS = ("oc_version", "oc_status_string", "oc_layout_create", "oc_layout_destroy", "oc_layout_query",
           "oc_reset", "oc_step", "oc_rollout", "oc_rollout_host", "oc_step_host", "oc_host_stepper_create", "oc_host_stepper_step",
           "oc_host_stepper_wait", "oc_host_stepper_destroy", "oc_get_obs", "oc_get_obs_all", "oc_prng_split", "oc_prng_chain",
           "oc_policy_create", "oc_policy_update", "oc_policy_set_shard", "oc_policy_destroy", "oc_policy_forward", "oc_policy_forward_cells", "oc_policy_sample",
           "oc_layout_obs_template", "oc_get_obs_cells", "oc_layout_cell_tables")

_lib = None


class OcError(RuntimeError):
    pass


def load():
    """Load the CUDA library or raise -- never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OcError(f"{LIB_PATH} is missing: the CUDA extension has not been built "
                      "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise OcError(f"{LIB_PATH} does not export {name}")
    lib.oc_version.restype = C.c_int
    lib.oc_status_string.restype = C.c_char_p
    lib.oc_status_string.argtypes = [C.c_int]
    lib.oc_layout_create.restype = C.c_int
    lib.oc_layout_create.argtypes = [C.POINTER(LayoutDesc), C.POINTER(C.c_void_p)]
    lib.oc_layout_destroy.restype = None
    lib.oc_layout_destroy.argtypes = [C.c_void_p]
    lib.oc_layout_query.restype = C.c_int
    lib.oc_layout_query.argtypes = [C.c_void_p, C.POINTER(LayoutInfo)]
    lib.oc_reset.restype = C.c_int
    lib.oc_reset.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, StatePtrs, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oc_get_obs.restype = C.c_int
    lib.oc_get_obs.argtypes = [C.c_void_p, C.c_int64, StatePtrs, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oc_get_obs_all.restype = C.c_int
    lib.oc_get_obs_all.argtypes = [C.c_void_p, C.c_int64, StatePtrs, C.c_void_p, C.c_void_p]
    lib.oc_step.restype = C.c_int
    lib.oc_step.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, StatePtrs, C.c_void_p, C.c_void_p,
                            C.POINTER(StatePtrs), StatePtrs, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.POINTER(StepExtras), C.c_void_p]
    lib.oc_rollout.restype = C.c_int
    lib.oc_rollout.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, StatePtrs, C.POINTER(RolloutIO),
                               C.POINTER(StepExtras), C.c_void_p]
    lib.oc_rollout_host.restype = C.c_int
    lib.oc_rollout_host.argtypes = [C.c_void_p, C.c_int64, C.c_int32, StatePtrs, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(StepExtras), C.c_void_p]
    lib.oc_step_host.restype = C.c_int
    lib.oc_step_host.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, StatePtrs, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.POINTER(StepExtras), C.c_void_p]
    lib.oc_host_stepper_create.restype = C.c_int
    lib.oc_host_stepper_create.argtypes = [C.c_void_p, C.c_int64, StatePtrs, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.POINTER(StepExtras), C.c_void_p, C.POINTER(C.c_void_p)]
    lib.oc_host_stepper_step.restype = C.c_int
    lib.oc_host_stepper_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oc_host_stepper_wait.restype = C.c_int
    lib.oc_host_stepper_wait.argtypes = [C.c_void_p]
    lib.oc_host_stepper_destroy.restype = None
    lib.oc_host_stepper_destroy.argtypes = [C.c_void_p]
    lib.oc_prng_split.restype = C.c_int
    lib.oc_prng_split.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.oc_prng_chain.restype = C.c_int
    lib.oc_prng_chain.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.oc_policy_create.restype = C.c_int
    lib.oc_policy_create.argtypes = [C.POINTER(PolicyDesc), C.POINTER(C.c_void_p)]
    lib.oc_policy_update.restype = C.c_int
    lib.oc_policy_update.argtypes = [C.c_void_p, C.POINTER(PolicyDesc), C.c_void_p]
    lib.oc_policy_set_shard.restype = C.c_int
    lib.oc_policy_set_shard.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    lib.oc_policy_destroy.restype = None
    lib.oc_policy_destroy.argtypes = [C.c_void_p]
    lib.oc_policy_forward.restype = C.c_int
    lib.oc_policy_forward.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7
    lib.oc_policy_forward_cells.restype = C.c_int
    lib.oc_policy_forward_cells.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7
    lib.oc_policy_sample.restype = C.c_int
    lib.oc_policy_sample.argtypes = [C.c_int, C.c_int64, C.c_int32] + [C.c_void_p] * 5
    lib.oc_get_obs_cells.restype = C.c_int
    lib.oc_get_obs_cells.argtypes = [C.c_void_p, C.c_int64, StatePtrs, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oc_layout_cell_tables.restype = C.c_int
    lib.oc_layout_cell_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oc_layout_obs_template.restype = C.c_int
    lib.oc_layout_obs_template.argtypes = [C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def check(status, what):
    if status != OC_OK:
        msg = load().oc_status_string(status).decode()
        raise OcError(f"{what} failed: {msg} (status {status})")
