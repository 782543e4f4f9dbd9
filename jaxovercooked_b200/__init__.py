"""jaxovercooked_b200 -- the batched Overcooked environment step of JAXOvercooked, rebuilt for NVIDIA B200.

Public surface (mirrors the reference's `jax_marl` for this path; see DESIGN.md / INTEGRATION.md):
    make("overcooked", layout=..., random_reset=False, max_steps=400, task_id=0)   -> Overcooked
    env.reset(key) / env.step(key, state, actions[, reset_state]) / env.get_obs(state)
    State, LogWrapper, overcooked_layouts, layout_grid_to_dict, pad_layouts, random.PRNGKey / random.split
Importing the package is cheap and CUDA-free; constructing an env loads liboc_b200.so and raises if it is
missing -- there is no CPU fallback.
"""
from .layouts import (Layout, overcooked_layouts, layout_grid_to_dict, pad_layouts, cl_sequence_padded,  # noqa: F401
                      hard_layouts, medium_layouts, easy_layouts, padded_layouts)

__version__ = "0.1.2"

_LAZY = {"Overcooked": "env", "State": "env", "make": "env", "vmap": "env", "registered_envs": "env",
         "Actions": "env", "HostStepper": "env", "HostRollout": "env", "LogWrapper": "wrappers", "LogEnvState": "wrappers", "CTRolloutManager": "wrappers", "random": None, "env": None,
         "wrappers": None, "dist": None, "rollout": None, "RolloutBuffer": "rollout", "PolicyRollout": "rollout", "policy": None,
         "ActorCritic": "policy"}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod = _LAZY[name]
        if mod is None:
            return importlib.import_module(f".{name}", __name__)
        return getattr(importlib.import_module(f".{mod}", __name__), name)
    raise AttributeError(name)
