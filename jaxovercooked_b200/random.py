"""The slice of `jax.random` the callers of the env use, on the device: PRNGKey and split.

`split` runs the threefry2x32 kernel behind `oc_prng_split` (include/oc_b200.h); keys are raw uint32[2] CUDA
tensors as in JAX's legacy key representation.  Callers' recipe (baselines/IPPO_CL.py:477,535):
    rng, _rng = split(rng);  step_keys = split(_rng, num_envs)
`split(key, num, first=a, count=b)` returns rows [a, a+b) only -- a GPU that owns that shard of a global batch
of `num` environments gets exactly the keys the single-device program would use for them."""
import ctypes as C

import torch

from . import _lib

_MODES = {"threefry_legacy": _lib.PRNG_LEGACY, "threefry_partitionable": _lib.PRNG_PARTITIONABLE}


def PRNGKey(seed, device=None):
    """jax.random.PRNGKey(seed) with x64 disabled: [0, seed mod 2**32]."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    t = torch.tensor([0, int(seed) & 0xFFFFFFFF], dtype=torch.int64)
    return t.to(torch.uint32).to(dev)


def split(key, num=2, *, first=0, count=None, impl="threefry_legacy", out=None):
    if key.dtype != torch.uint32 or tuple(key.shape) != (2,) or not key.is_cuda:
        raise ValueError("key must be a uint32 CUDA tensor of shape [2]")
    count = num - first if count is None else count
    key = key.contiguous()
    if out is None:
        out = torch.empty((count, 2), dtype=torch.uint32, device=key.device)
    lib = _lib.load()
    with torch.cuda.device(key.device):
        _lib.check(lib.oc_prng_split(_MODES[impl], key.data_ptr(), num, first, count, out.data_ptr(),
                                     C.c_void_p(torch.cuda.current_stream(key.device).cuda_stream)), "oc_prng_split")
    return out


class LazySplit:
    """`split(key, num)[first : first+count]` that is never materialised: pass it as the `key` of `env.step`
    and the kernel derives each env's key itself, only when that env resets (oc_step_extras.split_key)."""

    def __init__(self, key, num, first=0, count=None):
        if key.dtype != torch.uint32 or tuple(key.shape) != (2,) or not key.is_cuda:
            raise ValueError("key must be a uint32 CUDA tensor of shape [2]")
        self.key, self.num, self.first = key, int(num), int(first)
        self.count = self.num - self.first if count is None else int(count)
        if self.first < 0 or self.count < 0 or self.first + self.count > self.num:
            raise ValueError("shard [first, first+count) outside [0, num)")

    @property
    def shape(self):
        return (self.count, 2)

    def materialize(self, impl="threefry_legacy"):
        return split(self.key, self.num, first=self.first, count=self.count, impl=impl)


def split_chain(key, n, *, impl="threefry_legacy", out=None):
    """n times `key, sub = split(key)`: returns the n subkeys u32[n,2] (in `out` if given) and advances `key` IN PLACE."""
    if key.dtype != torch.uint32 or tuple(key.shape) != (2,) or not key.is_cuda or not key.is_contiguous():
        raise ValueError("key must be a contiguous uint32 CUDA tensor of shape [2]")
    subs = torch.empty((n, 2), dtype=torch.uint32, device=key.device) if out is None else out
    if subs.dtype != torch.uint32 or tuple(subs.shape) != (n, 2) or not subs.is_contiguous() or subs.device != key.device:
        raise ValueError("out must be a contiguous uint32 tensor of shape [n, 2] on the key's device")
    lib = _lib.load()
    with torch.cuda.device(key.device):
        _lib.check(lib.oc_prng_chain(_MODES[impl], key.data_ptr(), n, subs.data_ptr(),
                                     C.c_void_p(torch.cuda.current_stream(key.device).cuda_stream)), "oc_prng_chain")
    return subs
