"""PRNG restatement (oracle/prng.py NumPy, oracle/oc_oracle.c C): known-answer vectors and cross-checks.

KAT sources: Random123 threefry2x32 test vectors (20 rounds); the `split(PRNGKey(0))` values printed in JAX's
PRNG documentation for the legacy (`jax_threefry_partitionable=False`) and the partitionable mode.
Everything above `split` (randint / choice) has no external KAT -- parity unpinned against real JAX."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import prng

KAT = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
       ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
       ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]


@pytest.mark.parametrize("key,ctr,exp", KAT)
def test_threefry_kat(key, ctr, exp):
    y0, y1 = prng.threefry2x32(key[0], key[1], ctr[0], ctr[1])
    assert (int(y0), int(y1)) == exp
    assert orc.threefry2x32(key[0], key[1], ctr[0], ctr[1]) == exp


def test_split_kat_legacy():
    assert prng.split(prng.PRNGKey(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert prng.split(prng.PRNGKey(42)).tolist() == [[2465931498, 3679230171], [255383827, 267815257]]
    assert prng.random_bits(prng.PRNGKey(0), 3).tolist() == [4146024105, 1351547692, 2718843009]


def test_split_kat_partitionable():
    assert prng.split(prng.PRNGKey(0), 2, True).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]


@pytest.mark.parametrize("part", [False, True])
@pytest.mark.parametrize("num", [1, 2, 3, 16, 17, 1000])
def test_c_split_matches_numpy(part, num):
    for seed in (0, 30, 12345):
        k = prng.PRNGKey(seed)
        assert np.array_equal(orc.split(k, num, part), prng.split(k, num, part))


def test_split_row_closed_form():
    """Row i of legacy split(key, N) depends only on (key, i, N): word j=2i+c is y0 of block (j, j+N) if j < N
    else y1 of block (j-N, j).  The multi-GPU sharding relies on this to derive keys from GLOBAL env indices."""
    key = prng.PRNGKey(7)
    N = 37
    full = prng.split(key, N)
    for i in (0, 1, 17, 18, 19, 36):
        for c in (0, 1):
            j = 2 * i + c
            if j < N:
                w = prng.threefry2x32(key[0], key[1], j, j + N)[0]
            else:
                w = prng.threefry2x32(key[0], key[1], j - N, j)[1]
            assert int(w) == int(full[i, c])


@pytest.mark.parametrize("part", [False, True])
def test_randint_reduces_to_low_bits_for_span_4(part):
    rng = np.random.default_rng(0)
    for _ in range(50):
        key = rng.integers(0, 2**32, 2, dtype=np.uint32)
        k1, k2 = prng.split(key, 2, part)
        lb = prng.random_bits(k2, 2, part)
        assert prng.randint(key, 2, 0, 4, part).tolist() == (lb & 3).astype(np.int32).tolist()


@pytest.mark.parametrize("part", [False, True])
def test_integer_choice_equals_float_gumbel_topk(part):
    """with p in {0,1} the float32 Gumbel top-2 equals 'largest mantissa first, ties low index'."""
    rng = np.random.default_rng(1)
    for n in (4, 16, 20, 45, 99, 100, 110, 225, 256):
        for _ in range(200):
            key = rng.integers(0, 2**32, 2, dtype=np.uint32)
            free = rng.random(n) < 0.4
            free[rng.integers(0, n, 3)] = True
            if free.sum() < 2:
                continue
            f = prng.choice_no_replace_p(key, n, 2, free.astype(np.float32), part)
            i = prng.choice_no_replace_p_int(key, n, 2, free, part)
            assert f.tolist() == i.tolist()


# jax.random.normal(jax.random.PRNGKey(0), (10,)) as printed in JAX's own quickstart: the older docs (legacy threefry,
# jax < 0.5) and the current docs (jax_threefry_partitionable=True).  normal = sqrt(2) * erfinv(uniform(-1, 1)), and
# uniform is the mantissa trick on random_bits(key, 10): an external known-answer for `random_bits` in both modes
# (counter layout of the legacy mode, y0 ^ y1 of the partitionable mode) and for the bits -> float32 mapping.
JAX_QUICKSTART_NORMAL = {
    False: [-0.3721109, 0.26423115, -0.18252768, -0.7368197, -0.44030377, -0.1521442, -0.67135346, -0.5908641,
            0.73168886, 0.5673026],
    True: [1.6226422, 2.0252647, -0.43359444, -0.07861735, 0.1760909, -0.97208923, -0.49529874, 0.4943786,
           0.6643493, -0.9501635],
}


@pytest.mark.parametrize("part", [False, True])
def test_random_bits_against_jax_quickstart_normal(part):
    from scipy.special import erfinv
    bits = prng.random_bits(prng.PRNGKey(0), 10, part)
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo, hi = np.nextafter(np.float32(-1), np.float32(0)), np.float32(1)
    u = np.maximum(lo, f * (hi - lo) + lo)                       # jax.random.uniform(key, (10,), minval=lo, maxval=1)
    x = np.sqrt(2.0) * erfinv(u.astype(np.float64))
    assert np.allclose(x, JAX_QUICKSTART_NORMAL[part], rtol=0, atol=3e-7)
