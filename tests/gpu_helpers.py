"""Helpers for the -m gpu parity tests: run the CUDA path through the host mirror (ctypes -> C ABI) and compare
with the CPU oracle / golden fixtures field by field, bit-exact."""
import numpy as np
import torch

from tests.helpers import DYN_FIELDS, STATIC_FIELDS, assert_same


def state_to_numpy(st):
    return {f: getattr(st, f).cpu().numpy() for f in DYN_FIELDS + STATIC_FIELDS}


def to_dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def keys_to_dev(k):
    return torch.from_numpy(np.ascontiguousarray(k).view(np.int32)).cuda().view(torch.uint32)


def compare_state(st, ref, what, fields=DYN_FIELDS + STATIC_FIELDS):
    got = state_to_numpy(st)
    for f in fields:
        assert_same(got[f], ref[f], f"{what} State.{f}")


def numpy_state_to_dev(env, ref):
    import jaxovercooked_b200 as ocb
    from jaxovercooked_b200.env import _STATE_DTYPES
    kw = {}
    for f, dt in _STATE_DTYPES.items():
        a = np.ascontiguousarray(ref[f])
        if dt == torch.uint32:
            kw[f] = torch.from_numpy(a.view(np.int32)).cuda().view(torch.uint32)
        else:
            kw[f] = torch.from_numpy(a).cuda()
    return ocb.State(**kw)
