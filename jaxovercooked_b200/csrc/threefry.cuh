// threefry.cuh: jax.random on the device (threefry2x32, split, random_bits, randint) in both threefry modes -- part of the sm_100a Overcooked kernels (see oc_b200.cu for the overview)
#pragma once
#include <cstdint>

namespace {

// ------------------------------------------------------------------------------------------------ PRNG

__device__ __forceinline__ void tf2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t &y0, uint32_t &y1) {
    const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
#define OC_MIX(r) x0 += x1; x1 = __funnelshift_l(x1, x1, r); x1 ^= x0;
    x0 += k0; x1 += k1;
    OC_MIX(13) OC_MIX(15) OC_MIX(26) OC_MIX(6)
    x0 += k1; x1 += k2 + 1u;
    OC_MIX(17) OC_MIX(29) OC_MIX(16) OC_MIX(24)
    x0 += k2; x1 += k0 + 2u;
    OC_MIX(13) OC_MIX(15) OC_MIX(26) OC_MIX(6)
    x0 += k0; x1 += k1 + 3u;
    OC_MIX(17) OC_MIX(29) OC_MIX(16) OC_MIX(24)
    x0 += k1; x1 += k2 + 4u;
    OC_MIX(13) OC_MIX(15) OC_MIX(26) OC_MIX(6)
    x0 += k2; x1 += k0 + 5u;
#undef OC_MIX
    y0 = x0; y1 = x1;
}

// key, sub = jax.random.split(key)
__device__ __forceinline__ void split2(int part, uint32_t &k0, uint32_t &k1, uint32_t &s0, uint32_t &s1) {
    uint32_t a, b, c, d;
    if (part) {
        tf2x32(k0, k1, 0u, 0u, a, b);
        tf2x32(k0, k1, 0u, 1u, c, d);
        k0 = a; k1 = b; s0 = c; s1 = d;
    } else {
        tf2x32(k0, k1, 0u, 2u, a, b);
        tf2x32(k0, k1, 1u, 3u, c, d);
        k0 = a; k1 = c; s0 = b; s1 = d;
    }
}

// word j of random_bits(key, n) (32-bit)
__device__ __forceinline__ uint32_t bits_at(int part, uint32_t k0, uint32_t k1, int n, int j) {
    uint32_t a, b;
    if (part) { tf2x32(k0, k1, 0u, (uint32_t)j, a, b); return a ^ b; }
    const int m = (n + 1) >> 1;
    if (j < m) { tf2x32(k0, k1, (uint32_t)j, (j + m < n) ? (uint32_t)(j + m) : 0u, a, b); return a; }
    tf2x32(k0, k1, (uint32_t)(j - m), (uint32_t)j, a, b);
    return b;
}

// element j of jax.random.randint(key, (n,), 0, span)
__device__ __forceinline__ int randint_at(int part, uint32_t k0, uint32_t k1, int n, int j, uint32_t span) {
    uint32_t s0, s1;
    split2(part, k0, k1, s0, s1);          // k = first row (higher bits), s = second row (lower bits)
    uint32_t mult = 65536u % span;
    mult = (mult * mult) % span;
    const uint32_t lb = bits_at(part, s0, s1, n, j);
    uint32_t off = lb % span;
    if (mult != 0u) off += (bits_at(part, k0, k1, n, j) % span) * mult;
    return (int)(off % span);
}

// row i of jax.random.split(key, num): depends only on (key, i, num)
__device__ __forceinline__ void split_row(int part, uint32_t k0, uint32_t k1, long long num, long long i,
                                          uint32_t &r0, uint32_t &r1) {
    uint32_t a, b;
    if (part) { tf2x32(k0, k1, 0u, (uint32_t)i, r0, r1); return; }
    // legacy: flat word j of threefry_2x32(key, iota(2*num)); the counters are split in two halves of `num`
    long long j = 2 * i;
    if (j < num) { tf2x32(k0, k1, (uint32_t)j, (uint32_t)(j + num), a, b); r0 = a; }
    else { tf2x32(k0, k1, (uint32_t)(j - num), (uint32_t)j, a, b); r0 = b; }
    j += 1;
    if (j < num) { tf2x32(k0, k1, (uint32_t)j, (uint32_t)(j + num), a, b); r1 = a; }
    else { tf2x32(k0, k1, (uint32_t)(j - num), (uint32_t)j, a, b); r1 = b; }
}

}  // namespace
