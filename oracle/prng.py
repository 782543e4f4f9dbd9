"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the parts of `jax.random` the Overcooked hot path calls.

Nothing under `oracle/` is product code: only `tests/`, `__graft_entry__.smoke()`, `bench.py`'s
`cpu_baseline` / `--impl reference` legs and `tests/golden/make_golden.py` may import it.

The algorithm lives in a third-party dependency of the reference that is NOT vendored under
/root/reference: **jax** (requirements.txt:4 `jax[cuda12]`, unpinned; requirements.txt:3 hints
jaxlib==0.4.25).  It is restated here from the published algorithm (jax/_src/prng.py,
jax/_src/random.py as of 0.4.25): threefry2x32 (Random123, 20 rounds), `split`, `random_bits`,
`randint`, `uniform`, `gumbel`, `choice`.  Call sites in the reference that pin what we need:
    jax_marl/environments/multi_agent_env.py:52            split(key)
    overcooked_environment/overcooked.py:164-166           split, choice(p=..., replace=False)
    overcooked_environment/overcooked.py:173-174           split, choice(arange(4))  == randint(0, 4)
    overcooked_environment/overcooked.py:197-200           split, randint(0, 24)
    overcooked_environment/overcooked.py:225-228           split, choice(4 items)    == randint(0, 4)
    baselines/IPPO_CL.py:477,535                           split(rng, num_envs)

PARITY STATUS: no JAX is installed in this image.  Pinned by external known-answer vectors only
(tests/test_prng.py): Random123 threefry2x32 vectors; the `split(PRNGKey(0))` values printed in JAX's
documentation (both modes); and `random.normal(PRNGKey(0), (10,))` as printed in JAX's quickstart
(both modes), which pins `random_bits` and the bits->float mapping.  `randint`'s two-draw combination and
`choice`'s Gumbel/argsort route have no external vector: they stay "parity unpinned" against real JAX.

Two modes, selected by `partitionable`:
  False (default)  jax_threefry_partitionable=False, the default for jax < 0.5.0 (matches the 0.4.25 hint)
  True             jax_threefry_partitionable=True,  the default for jax >= 0.5.0
"""
import numpy as np

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Random123 Threefry-2x32, 20 rounds; all operands uint32 (scalars or arrays), wrap-around."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=U32)
        k1 = np.asarray(k1, dtype=U32)
        x0 = np.array(x0, dtype=U32, copy=True)
        x1 = np.array(x1, dtype=U32, copy=True)
        ks = (k0, k1, k0 ^ k1 ^ U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + U32(i + 1)
    return x0, x1


def PRNGKey(seed):
    """jax.random.PRNGKey for a 32-bit seed with x64 disabled: [0, seed]."""
    return np.array([0, int(seed) & 0xFFFFFFFF], dtype=U32)


def _bits_legacy(key, n):
    """`threefry_2x32(key, iota(n))`: counters split in halves, odd n padded with one 0."""
    if n == 0:
        return np.zeros((0,), U32)
    c = np.arange(n, dtype=U32)
    if n % 2:
        c = np.concatenate([c, np.zeros(1, U32)])
    m = c.size // 2
    y0, y1 = threefry2x32(key[0], key[1], c[:m], c[m:])
    return np.concatenate([y0, y1])[:n]


def _bits_partitionable(key, n):
    """`_threefry_random_bits_partitionable` for 32-bit words, flat index j < 2**32: y0 ^ y1 of block (0, j)."""
    j = np.arange(n, dtype=U32)
    y0, y1 = threefry2x32(key[0], key[1], np.zeros(n, U32), j)
    return y0 ^ y1


def random_bits(key, n, partitionable=False):
    return _bits_partitionable(key, n) if partitionable else _bits_legacy(key, n)


def split(key, num=2, partitionable=False):
    """jax.random.split(key, num) -> uint32[num, 2]."""
    key = np.asarray(key, dtype=U32)
    if partitionable:
        i = np.arange(num, dtype=U32)
        y0, y1 = threefry2x32(key[0], key[1], np.zeros(num, U32), i)
        return np.stack([y0, y1], axis=1)
    return _bits_legacy(key, 2 * num).reshape(num, 2)


def randint(key, n, lo, hi, partitionable=False):
    """jax.random.randint(key, (n,), lo, hi) for int32 and hi > lo, both in int32 range -> int32[n]."""
    k1, k2 = split(key, 2, partitionable)
    hb = random_bits(k1, n, partitionable)
    lb = random_bits(k2, n, partitionable)
    span = U32(hi - lo)
    with np.errstate(over="ignore"):
        mult = U32((1 << 16) % int(span))
        mult = U32((int(mult) * int(mult)) % int(span))
        off = (hb % span) * mult + (lb % span)  # uint32 wrap-around, as lax.mul/lax.add on uint32
        off = off % span
    return (np.int32(lo) + off.astype(np.int32)).astype(np.int32)


def uniform_f32(key, n, minval, maxval, partitionable=False):
    """jax.random.uniform(key, (n,), float32, minval, maxval): mantissa trick, then max(minval, f*(hi-lo)+lo)."""
    bits = random_bits(key, n, partitionable)
    fb = (bits >> U32(9)) | U32(0x3F800000)
    f = fb.view(np.float32) - np.float32(1.0)
    lo = np.float32(minval)
    hi = np.float32(maxval)
    return np.maximum(lo, f * (hi - lo) + lo)


def gumbel_f32(key, n, partitionable=False):
    u = uniform_f32(key, n, np.finfo(np.float32).tiny, 1.0, partitionable)
    with np.errstate(divide="ignore"):
        return -np.log(-np.log(u))


def choice_no_replace_p(key, n_inputs, n_draws, p, partitionable=False):
    """jax.random.choice(key, arange(n_inputs), (n_draws,), replace=False, p=p): Gumbel top-k in float32."""
    p = np.asarray(p, dtype=np.float32)
    with np.errstate(divide="ignore"):
        g = -gumbel_f32(key, n_inputs, partitionable) - np.log(p)
    return np.argsort(g, kind="stable")[:n_draws]


def choice_no_replace_p_int(key, n_inputs, n_draws, free_mask, partitionable=False):
    """Integer-only equivalent of `choice_no_replace_p` for p in {0,1}: the draws are the
    allowed cells with the largest 23-bit mantissas `bits >> 9`, ties to the lower index.  This is what the C
    oracle and the CUDA kernel compute; tests/test_prng.py checks it against the float32 form above."""
    bits = random_bits(key, n_inputs, partitionable)
    mant = (bits >> U32(9)).astype(np.int64)
    mant = np.where(np.asarray(free_mask, dtype=bool), mant, -1)
    order = np.argsort(-mant, kind="stable")
    return order[:n_draws]
