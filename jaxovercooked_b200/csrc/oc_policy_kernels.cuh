// oc_policy_kernels.cuh -- device side of the actor-critic MLP forward of the rollout path (architectures/mlp.py:11-59 as
// called at baselines/IPPO_MLP.py:435-439) for sm_100a: scope-table row f3, the consumer of the observation stream.
// The C ABI around it is in oc_policy.cu.
//
//   logits = Dense_2(act(Dense_1(act(Dense_0(x)))))          actor   K -> 128 -> 128 -> A
//   value  = Dense_5(act(Dense_4(act(Dense_3(x)))))[:, 0]    critic  K -> 128 -> 128 -> 1
//   x      = one uint8 observation row (h*w*26 bytes, > 95 % of them equal to the layout's static image)
//
// One persistent CTA per SM (16 warps; 32 for the compact source below) walks 128-row tiles:
//   layer 1  (sparse, CUDA cores, fp32): each warp owns a row.  It XORs the row against the observation template
//            in 16-byte pieces, turns the few differing bytes into (k, delta) entries in shared memory and
//            accumulates base + sum(delta * W1[k, :]) with lane j holding hidden units 8j..8j+7 of the 256
//            (actor | critic); `base` = bias + template . W1 is precomputed.  Any input is handled (a dense row just
//            produces many entries); only the speed depends on how close rows are to the template.
//   layer 2  (dense, tensor cores): the activated hidden tile is written as fp16 in the 128-byte-swizzled K-major
//            layout tcgen05 reads; one thread issues 2 x 8 tcgen05.mma (M128 N128 K16, fp32 accumulators in TMEM
//            columns 0..127 actor / 128..255 critic) against the resident fp16 W2 images, then tcgen05.commit.
//   heads    all warps read their TMEM quadrant / column group with tcgen05.ld, apply bias + activation and the
//            128 -> A / 128 -> 1 heads on CUDA cores; 4-8 lanes per row finish it: logits, value and, when a key is
//            given, the categorical draw (threefry gumbel, argmax) and its log-probability.
//   Rows produced by the env kernel need not be scanned at all: oc_policy_forward_cells reads the env kernel's compact
//   observations (oc_step_extras.cells) and layer 1 becomes base + a few rows of a precomputed (cell, code) table.
//
// Layer 1 is 1.5 kFMA per row when done sparsely against 133 kMAC densely; layer 2 has no sparsity, which is where
// the tensor cores pay (DESIGN.md section 9 has the numbers).
#pragma once
#include "oc_b200.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "oc_common.cuh"
#include "threefry.cuh"

namespace {

constexpr int kHid = OC_POLICY_HIDDEN;        // 128
constexpr int kHid2 = 2 * kHid;               // actor | critic
constexpr int kTileM = 128;
constexpr int kObsWarps = 16;                 // warps per CTA when rows are uint8 observations (register-hungry scan)
#ifndef OC_POLICY_CELL_WARPS
#define OC_POLICY_CELL_WARPS 32               // tuning: 16 leaves half of the SM's registers to a concurrently running env kernel
#endif
constexpr int kCellWarps = OC_POLICY_CELL_WARPS;   // ... and when rows are compact observations (more warps hide the gather latency)
constexpr int kListCap = 512;                 // (k, delta) entries per warp between two flushes
constexpr int kListWords = kListCap + 8;      // + padding to a multiple of 8 entries
constexpr int kNetBytes = kTileM * kHid * 2;  // one fp16 128x128 operand image: 32 KB = 2 swizzle atoms of 16 KB
constexpr int kTmemCols = 512;               // two accumulator buffers of 256 columns (actor | critic)

struct PolicySmall {
    float base[kHid2];          // Dense_0/Dense_3 bias + template . kernel
    float b2[kHid2];            // Dense_1 | Dense_4 bias
    float hw[kHid * 8];         // Dense_2 kernel, row c = 8 floats (A used, rest 0; with A <= 6 slot 6 holds b2[c])
    float cw[kHid];             // Dense_5 kernel
    float b3[8];                // Dense_2 bias (padded)
    float b3c;                  // Dense_5 bias
    float pad[3];
};
static_assert(sizeof(PolicySmall) % 16 == 0, "PolicySmall is copied in 16-byte pieces");

// shared-memory carve-up (offsets from a 1024-byte aligned base)
constexpr int kOffA = 0;
constexpr int kOffW2 = kOffA + 2 * kNetBytes;
constexpr int kOffSmall = kOffW2 + 2 * kNetBytes;
constexpr int kOffHpA = kOffSmall + (int)sizeof(PolicySmall);
// everything behind the operands depends on the row source: the compact source (32 warps, 4 column groups per net) needs no
// entry lists and no template, which leaves ~90 KB of the SM's 256 KB to L1 for the embedding-table gathers
__host__ __device__ constexpr int off_hpc(bool cells) { return kOffHpA + kTileM * (cells ? 4 : 2) * 8 * 4; }     // actor head partials [128][NG][8]
__host__ __device__ constexpr int off_list(bool cells) { return off_hpc(cells) + kTileM * (cells ? 4 : 2) * 4; } // critic head partials [128][NG]
__host__ __device__ constexpr int off_scan(bool cells) { return off_list(cells) + (cells ? kCellWarps * 72 * 4 : kObsWarps * kListWords * 4); }
__host__ __device__ constexpr int off_bar(bool cells) { return off_scan(cells) + 64; }          // scan slot -> cell table of the compact source
__host__ __device__ constexpr int off_tmpl(bool cells) { return off_bar(cells) + 48; }          // two accumulator barriers, TMEM base address, W2 barrier
static_assert(kOffHpA % 16 == 0 && off_list(true) % 16 == 0 && off_list(false) % 16 == 0 && off_tmpl(true) % 16 == 0 &&
              off_tmpl(false) % 16 == 0, "alignment");

struct PParams {
    const __half *w1;           // [K][256] fp16
    const uint8_t *w2img;       // 2 x 32 KB fp16 images (swizzled, K-major, row = output unit)
    const PolicySmall *small;
    const uint8_t *tmpl;        // [K]
    const uint4 *obs16;         // observation base rounded down to 16 bytes
    int a0;                     // ... and the bytes that were rounded off
    long long M;
    int K, A, part;
    int nvar, gshift, tv_stride;
    int n_tiles;
    int rpt;                    // rows per tile: 128, or 64 / 32 when that is what it takes to give every SM a tile
    float *logits, *value, *log_prob;
    const uint32_t *key;
    int32_t *action;
    // compact-observation source (oc_policy_forward_cells): rows [0, N) are agent_0's views, [N, 2N) agent_1's
    const uint8_t *cells;       // [N][cell_stride]
    const __half *emb;          // [C*68 + 2][256] fp16: layer-1 contribution of (cell, code as the viewer sees it); then a zero row, the urgency row
    const uint8_t *scan_cells;  // [n_scan] cell index of every scan slot
    long long N;
    int cell_stride, cell_agents, n_scan, C;
    // sharded draws (oc_policy_set_shard): local unit i (an env for the compact source, a row otherwise) is unit noise_first + i
    // of noise_global units -- the gumbel noise is indexed as the single-device program over the whole batch would index it
    long long noise_first, noise_global;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// OC_POLICY_EXACT_TANH=1 (build option): libdevice's tanhf (<= 2 ulp) instead of the hardware approximation (max relative
// error 2^-11); DESIGN.md section 9 states what each reaches against the float32 oracle
#ifndef OC_POLICY_EXACT_TANH
#define OC_POLICY_EXACT_TANH 0
#endif
__device__ __forceinline__ float act_tanh(float x) {
#if OC_POLICY_EXACT_TANH
    return tanhf(x);
#else
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}
template <int ACT>
__device__ __forceinline__ float activate(float x) {
    return ACT == OC_ACT_RELU ? fminf(fmaxf(x, 0.0f), 65504.0f) : act_tanh(x);
}

// mask with bit 7 of every non-zero byte set
__device__ __forceinline__ uint32_t nz_bytes(uint32_t x) { return (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u; }

// tcgen05 shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 bytes apart
// (cute/arch/mma_sm100_desc.hpp SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout_type=2 [61,64))
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D=f32 [4,6)=1, A=B=f16 (0), both K-major, N>>3 [17,23), M>>4 [24,29)
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kHid >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Categorical(logits).sample(seed=key) and .log_prob(action) for one row (distrax over jax.random.categorical):
// action = argmax_a(logits[a] + gumbel(key, (1, M, A))[0, row, a]), first maximum wins; log_prob = log_softmax(logits)[action]
// A few consecutive lanes share a row: each draws the noise of its actions, a butterfly over those lanes picks the winner.
__device__ __forceinline__ float gumbel_at(int part, uint32_t k0, uint32_t k1, int n, int j) {
    // jax.random.gumbel: -log(-log(uniform(key, shape, minval=tiny, maxval=1)))
    const uint32_t bits = bits_at(part, k0, k1, n, j);
    const float f = __uint_as_float((bits >> 9) | 0x3f800000u) - 1.0f;
    const float u = fmaxf(1.17549435e-38f, f + 1.17549435e-38f);
    return -logf(-logf(u));
}

// TPR consecutive lanes share a row (4 or 8); lane part `s` owns the logits of actions s, s + TPR, ... in l[]: it draws their
// noise, butterflies over the TPR lanes pick the winner and form the log-sum-exp.
template <int TPR>
__device__ __forceinline__ void sample_lanes(int part, uint32_t k0, uint32_t k1, int n, long long row, int A, int s,
                                             const float (&l)[OC_POLICY_MAX_ACTIONS / TPR], int &arg, float &lp) {
    constexpr int NA = OC_POLICY_MAX_ACTIONS / TPR;
    float best = -INFINITY, mx = -INFINITY;
    arg = OC_POLICY_MAX_ACTIONS;
#pragma unroll
    for (int h = 0; h < NA; ++h) {
        const int a = s + TPR * h;
        if (a < A) {
            mx = fmaxf(mx, l[h]);
            const float z = l[h] + gumbel_at(part, k0, k1, n, (int)(row * A + a));
            if (arg == OC_POLICY_MAX_ACTIONS || z > best) { best = z; arg = a; }
        }
    }
#pragma unroll
    for (int d = 1; d < TPR; d <<= 1) {
        const float zo = __shfl_xor_sync(0xffffffffu, best, d);
        const int ao = __shfl_xor_sync(0xffffffffu, arg, d);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        if (ao < OC_POLICY_MAX_ACTIONS && (arg == OC_POLICY_MAX_ACTIONS || zo > best || (zo == best && ao < arg))) { best = zo; arg = ao; }
    }
    float sum = 0.f, la = 0.f;
#pragma unroll
    for (int h = 0; h < NA; ++h) {
        const int a = s + TPR * h;
        if (a < A) { sum += expf(l[h] - mx); if (a == arg) la = l[h]; }
    }
#pragma unroll
    for (int d = 1; d < TPR; d <<= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        la += __shfl_xor_sync(0xffffffffu, la, d);
    }
    lp = la - (mx + logf(sum));
}

__global__ void oc_policy_sample_kernel(int part, long long M, int A, const float *__restrict__ logits,
                                        const uint32_t *__restrict__ key, int32_t *action, float *log_prob) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = t >> 2;
    const int s = (int)(t & 3);
    const bool live = row < M;
    float l[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) l[h] = (live && s + 4 * h < A) ? logits[row * A + s + 4 * h] : 0.f;
    int arg;
    float lp;
    sample_lanes<4>(part, __ldg(key), __ldg(key + 1), (int)(M * A), live ? row : 0, A, s, l, arg, lp);
    if (live && s == 0) {
        if (action) action[row] = arg;
        if (log_prob) log_prob[row] = lp;
    }
}

// acc0 += lo(w) * d, acc1 += hi(w) * d with fp16 operands and fp32 accumulation (one FHFMA each on sm_100)
__device__ __forceinline__ void fma_h2(float &acc0, float &acc1, uint32_t w, unsigned short d) {
    asm("{\n\t.reg .f16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tfma.rn.f32.f16 %0, lo, %3, %0;\n\tfma.rn.f32.f16 %1, hi, %3, %1;\n\t}\n"
        : "+f"(acc0), "+f"(acc1) : "r"(w), "h"(d));
}

// accumulate `n` entries (k << 16 | fp16 delta) of this warp's list: acc[0..7] += delta * W1[k, 8*lane .. 8*lane+7]
// (W1 is stored as fp16 rows of 512 bytes: one 16-byte load per lane and entry, eight entries in flight)
__device__ __forceinline__ void gather_rows(const __half *__restrict__ w1, uint32_t *list, int n, int lane, float (&acc)[8]) {
    if (lane < 7) list[n + lane] = 0u;       // pad to a multiple of 8 with (k=0, delta=0)
    __syncwarp();
    const __half *wl = w1 + 8 * lane;
    for (int i = 0; i < n; i += 8) {
        const uint4 ea = *reinterpret_cast<const uint4 *>(list + i), eb = *reinterpret_cast<const uint4 *>(list + i + 4);
        const uint32_t e[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
        uint4 w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = __ldg(reinterpret_cast<const uint4 *>(wl + (e[u] >> 16) * kHid2));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned short d = (unsigned short)(e[u] & 0xffffu);
            fma_h2(acc[0], acc[1], w[u].x, d);
            fma_h2(acc[2], acc[3], w[u].y, d);
            fma_h2(acc[4], acc[5], w[u].z, d);
            fma_h2(acc[6], acc[7], w[u].w, d);
        }
    }
    __syncwarp();
}

// this lane's 16-byte pieces of one observation row (piece 32*g + lane), issued together so that one DRAM round trip
// covers the row; rows past the end load nothing.  Row `row_l` of tile `tile` starts at byte a0 + (tile*rpt + row_l)*K of
// the 16-byte aligned stream; tile*rpt*K is a multiple of 16 (rpt >= 32), so everything but the tile offset is 32-bit arithmetic.
template <int PL>
__device__ __forceinline__ void load_row(const PParams &p, long long tile, int row_l, int lane, uint4 (&pf)[PL]) {
    if (tile < p.n_tiles && row_l < p.rpt && tile * p.rpt + row_l < p.M) {
        const int local = p.a0 + row_l * p.K;
        const int np = ((local & 15) + p.K + 15) >> 4;
        const uint4 *src = p.obs16 + tile * (long long)(p.rpt / 16 * p.K) + (local >> 4);
#pragma unroll
        for (int g = 0; g < PL; ++g) {
            const int q = g * 32 + lane;
            pf[g] = q < np ? __ldg(src + q) : make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// append the differing bytes of one 16-byte piece to the list: entry = k << 16 | fp16(obs - template); bytes that belong
// to a neighbouring row (k outside [0, K)) become the neutral entry 0
__device__ __forceinline__ int emit_piece(uint32_t *list, int pos, const uint4 &o, const uint4 &t, const uint32_t (&mw)[4], int kbase, int K) {
    const uint32_t ow[4] = {o.x, o.y, o.z, o.w}, tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t mm = mw[i];
        while (mm) {
            const int sh = __ffs(mm) - 8;                 // bit 7 of byte b -> shift 8*b
            mm &= mm - 1;
            const int k = kbase + 4 * i + (sh >> 3);
            const int dlt = (int)((ow[i] >> sh) & 255u) - (int)((tw[i] >> sh) & 255u);
            const uint32_t e = ((uint32_t)k << 16) | (uint32_t)__half_as_ushort(__int2half_rn(dlt));
            list[pos++] = (unsigned)k < (unsigned)K ? e : 0u;
        }
    }
    return pos;
}

// this lane's bytes of one env's compact observation (slots lane and lane + 32).  In the compact source a tile holds
// rpt/2 environments: rows [0, rpt/2) are their agent_0 views, rows [rpt/2, rpt) the agent_1 views of the same envs.
__device__ __forceinline__ void load_cells(const PParams &p, long long tile, int e_l, int lane, uint32_t (&cb)[2]) {
    const int ept = p.rpt >> 1;
    if (tile < p.n_tiles && e_l < ept) {
        const int env = (int)tile * ept + e_l;
        if (env < (int)p.N) {
            const uint8_t *src = p.cells + (unsigned)(env * p.cell_stride + lane);
            cb[0] = lane < p.cell_stride ? __ldg(src) : 0xffu;
            if (p.cell_stride > 32) cb[1] = lane + 32 < p.cell_stride ? __ldg(src + 32) : 0xffu;
        }
    }
}

// PL > 0: rows are uint8 observations, PL 16-byte pieces per lane and row.  PL == 0: rows come from the env kernel's
// compact observations (cells) and layer 1 is a sum of precomputed (view, cell, code) embeddings.
template <int ACT, int PL>
__global__ void __launch_bounds__((PL > 0 ? kObsWarps : kCellWarps) * 32, (PL == 0 && kCellWarps == 16) ? 2 : 1) oc_policy_kernel(const PParams p) {
    constexpr int PLA = PL > 0 ? PL : 1;
    constexpr int NW = PL > 0 ? kObsWarps : kCellWarps;     // warps per CTA
    constexpr int kPolThreads = NW * 32;
    constexpr int RND = PL > 0 ? kTileM / NW : kTileM / 2 / NW;   // rows (uint8 source) or envs (compact source) per warp and tile
    constexpr int RSH = RND == 8 ? 3 : RND == 4 ? 2 : 1;          // log2 of it
    constexpr int NG = NW / 8;                               // 64- or 32-column groups per net in the heads phase (2 or 4)
    constexpr int TPR = kPolThreads / kTileM;                // threads per row in the finishing phase (4 or 8)
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = sm + kOffA;
    uint8_t *sW2 = sm + kOffW2;
    PolicySmall *sS = reinterpret_cast<PolicySmall *>(sm + kOffSmall);
    float *hpA = reinterpret_cast<float *>(sm + kOffHpA);      // actor head partials  [128 rows][NG column groups][8]
    constexpr bool kCells = PL == 0;
    float *hpC = reinterpret_cast<float *>(sm + off_hpc(kCells));      // critic head partials [128 rows][NG]
    uint32_t *lists = reinterpret_cast<uint32_t *>(sm + off_list(kCells));
    uint64_t *mbar = reinterpret_cast<uint64_t *>(sm + off_bar(kCells));
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(sm + off_bar(kCells) + 16);
    uint8_t *tmplv = sm + off_tmpl(kCells);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K;

    // Programmatic dependent launch: the next kernel in the stream (the env step) may start its prologue now, and this
    // kernel's own set-up -- parameters only -- runs while the previous kernel (the env step that writes the rows) drains.
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    // ---- one-time setup: resident W2 images, small parameters, template variants, TMEM, mbarrier
    {
        const uint4 *s2 = reinterpret_cast<const uint4 *>(p.small);
        uint4 *d2 = reinterpret_cast<uint4 *>(sS);
        for (int i = tid; i < (int)sizeof(PolicySmall) / 16; i += kPolThreads) d2[i] = __ldg(s2 + i);
        // template variants: variant v is the template repeated periodically, shifted so that a row whose first byte sits at
        // offset a_v inside its first 16-byte piece lines up with 16-byte reads (bytes of the neighbouring rows cancel too)
        if (PL > 0) {
            const int g = 1 << p.gshift;
            for (int i = tid; i < p.nvar * p.tv_stride; i += kPolThreads) {
                const int v = i / p.tv_stride, j = i - v * p.tv_stride;
                const int a = (p.a0 + v * g) & 15;
                tmplv[i] = __ldg(p.tmpl + (j - a + K) % K);
            }
        } else {
            for (int i = tid; i < p.n_scan; i += kPolThreads) sm[off_scan(kCells) + i] = __ldg(p.scan_cells + i);
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_addr(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_addr(mbar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_addr(mbar) + 8) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_addr(mbar) + 24) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        // the resident layer-2 weights (two 32 KB fp16 images) arrive through the bulk-copy engine while the first tile's
        // layer 1 runs; only this thread -- the one that issues the MMAs -- ever waits for them
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr(mbar) + 24), "r"(2 * kNetBytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(smem_addr(sW2)), "l"(p.w2img), "r"(2 * kNetBytes), "r"(smem_addr(mbar) + 24) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // W2 image: generic-proxy writes -> tensor-core reads
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const uint32_t sA_u = smem_addr(sA), sW2_u = smem_addr(sW2), bar_u = smem_addr(mbar);
    asm volatile("griddepcontrol.wait;\n" ::: "memory");           // the rows (and the key) are complete from here on
    uint32_t *list = lists + (warp & (kObsWarps - 1)) * kListWords;     // (k, delta) entries of this warp (uint8 source only)

    // Each warp walks the rows rr*16 + warp (rr = 0..7) of its CTA's tiles; the observation pieces of the next two
    // rows are always in flight in registers (pfA / pfB), also across the tile boundary and the layer-2 phase.
    int item = 0;
    uint4 pfA[PLA], pfB[PLA];
    uint32_t cbA[2] = {0xffu, 0xffu}, cbB[2] = {0xffu, 0xffu};
    if (PL > 0) {
        load_row<PLA>(p, blockIdx.x, warp, lane, pfA);
        load_row<PLA>(p, blockIdx.x, NW + warp, lane, pfB);
    } else {
        load_cells(p, blockIdx.x, warp, lane, cbA);
        load_cells(p, blockIdx.x, NW + warp, lane, cbB);
    }
    uint8_t *s_sc = reinterpret_cast<uint8_t *>(sm + off_scan(kCells));          // [n_scan] scan slot -> cell (cells source only)

    auto finish_row = [&](const int row_l, float (&acc)[8]) {
        // activation, fp16, swizzled K-major store: lane -> net = lane >> 4, 16-byte chunk j = lane & 15 of the row
        uint4 pk;
        {
            const __half2 h0 = __floats2half2_rn(activate<ACT>(acc[0]), activate<ACT>(acc[1]));
            const __half2 h1 = __floats2half2_rn(activate<ACT>(acc[2]), activate<ACT>(acc[3]));
            const __half2 h2 = __floats2half2_rn(activate<ACT>(acc[4]), activate<ACT>(acc[5]));
            const __half2 h3 = __floats2half2_rn(activate<ACT>(acc[6]), activate<ACT>(acc[7]));
            pk.x = *reinterpret_cast<const uint32_t *>(&h0); pk.y = *reinterpret_cast<const uint32_t *>(&h1);
            pk.z = *reinterpret_cast<const uint32_t *>(&h2); pk.w = *reinterpret_cast<const uint32_t *>(&h3);
        }
        const int j = lane & 15;
        uint8_t *dst = sA + (lane >> 4) * kNetBytes + (j >> 3) * (kNetBytes / 2) + row_l * 128 + (((j & 7) ^ (row_l & 7)) << 4);
        *reinterpret_cast<uint4 *>(dst) = pk;
    };

    // Layer 1 of BOTH agents' rows of one env from its compact observation.  The two views share everything but the agent
    // cells: shared = base + embeddings of the deviating scan cells (+ urgency row); view v adds the two agent embeddings
    // with the viewer bit of the codes as agent v sees them.  One decode and one set of shared gathers serve two rows.
    // What a slot means depends only on the lane, so it is decoded once: kind 1 = scan cell (table row sbase + code),
    // 2 = an agent's code (its cell is in the previous slot), 3 = urgency flag, 0 = padding.
    int kind[2], sbase[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const int slot = g * 32 + lane, ag = slot - p.cell_agents;
        kind[g] = (PL == 0 && slot < p.n_scan) ? 1 : (ag == 1 || ag == 3) ? 2 : ag == 4 ? 3 : 0;
        sbase[g] = (PL == 0 && slot < p.n_scan) ? (int)s_sc[slot] * OC_NUM_CELL_CODES : 0;
    }
    const int ag_slot0 = p.cell_agents + 1, ag_slot1 = p.cell_agents + 3;      // where the two agent codes sit
    int *wl = reinterpret_cast<int *>(lists) + warp * 72;        // this warp's compacted shared table rows (<= 60 + padding)
    auto do_env_cells = [&](uint32_t (&cb)[2], const long long tile, const int rr) {
        const int ept = p.rpt >> 1;
        const int e_l = rr * NW + warp;
        const int env = e_l < ept ? (int)tile * ept + e_l : (int)p.N;
        const bool valid_env = env < (int)p.N;
        float acc0[8], acc1[8];
        {
            const float4 b0 = *reinterpret_cast<const float4 *>(sS->base + 8 * lane);
            const float4 b1 = *reinterpret_cast<const float4 *>(sS->base + 8 * lane + 4);
            acc0[0] = b0.x; acc0[1] = b0.y; acc0[2] = b0.z; acc0[3] = b0.w;
            acc0[4] = b1.x; acc0[5] = b1.y; acc0[6] = b1.z; acc0[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc1[i] = acc0[i];
        }
        int cnt = 0;
        int a0v0 = 0, a1v0 = 0, a0v1 = 0, a1v1 = 0;             // table rows of agent 0 / 1 as seen by agent_0 / agent_1
        const int idx_zero = p.C * OC_NUM_CELL_CODES;
        if (valid_env) {
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (g == 1 && p.cell_stride <= 32) break;      // most layouts: one byte per lane covers the env
                const int b = (int)cb[g];
                const int prev = (int)__shfl_up_sync(0xffffffffu, cb[g], 1);     // an agent's cell sits one slot before its code
                const int ixa = prev * OC_NUM_CELL_CODES + b;                    // agent row as agent_0 sees it
                const int ixb = prev * OC_NUM_CELL_CODES + kCodeAgent + ((b - kCodeAgent) ^ 16);   // ... as agent_1 sees it
                if ((ag_slot0 >> 5) == g) {
                    a0v0 = __shfl_sync(0xffffffffu, ixa, ag_slot0 & 31);
                    a0v1 = __shfl_sync(0xffffffffu, ixb, ag_slot0 & 31);
                }
                if ((ag_slot1 >> 5) == g) {
                    a1v0 = __shfl_sync(0xffffffffu, ixa, ag_slot1 & 31);
                    a1v1 = __shfl_sync(0xffffffffu, ixb, ag_slot1 & 31);
                }
                const int ix = kind[g] == 1 ? sbase[g] + b : idx_zero + 1;
                const bool valid = kind[g] == 1 ? b != 0xff : kind[g] == 3 ? b != 0 : false;
                const unsigned bal = __ballot_sync(0xffffffffu, valid);
                if (valid) wl[cnt + __popc(bal & lt)] = ix;      // slot order: the sums below are deterministic
                cnt += __popc(bal);
            }
            if (lane == 0) wl[cnt] = idx_zero;                   // pad to a multiple of 2 with the zero row
            __syncwarp();
        }
        {   // this slot is free again: fetch the env after next
            const int it2 = item + 2;
            load_cells(p, (long long)blockIdx.x + (long long)(it2 >> RSH) * gridDim.x, (it2 & (RND - 1)) * NW + warp, lane, cb);
            ++item;
        }
        if (valid_env) {
            const __half *el = p.emb + 8 * lane;
            uint4 w[4];
            w[0] = __ldg(reinterpret_cast<const uint4 *>(el + (unsigned)a0v0 * kHid2));
            w[1] = __ldg(reinterpret_cast<const uint4 *>(el + (unsigned)a1v0 * kHid2));
            w[2] = __ldg(reinterpret_cast<const uint4 *>(el + (unsigned)a0v1 * kHid2));
            w[3] = __ldg(reinterpret_cast<const uint4 *>(el + (unsigned)a1v1 * kHid2));
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                fma_h2(acc0[0], acc0[1], w[u].x, (unsigned short)0x3c00); fma_h2(acc0[2], acc0[3], w[u].y, (unsigned short)0x3c00);
                fma_h2(acc0[4], acc0[5], w[u].z, (unsigned short)0x3c00); fma_h2(acc0[6], acc0[7], w[u].w, (unsigned short)0x3c00);
                fma_h2(acc1[0], acc1[1], w[2 + u].x, (unsigned short)0x3c00); fma_h2(acc1[2], acc1[3], w[2 + u].y, (unsigned short)0x3c00);
                fma_h2(acc1[4], acc1[5], w[2 + u].z, (unsigned short)0x3c00); fma_h2(acc1[6], acc1[7], w[2 + u].w, (unsigned short)0x3c00);
            }
            for (int i = 0; i < cnt; i += 2) {                   // shared rows, two in flight, added to both views
                const int2 ix = *reinterpret_cast<const int2 *>(wl + i);
                const uint4 s0 = __ldg(reinterpret_cast<const uint4 *>(el + (unsigned)ix.x * kHid2));
                const uint4 s1 = __ldg(reinterpret_cast<const uint4 *>(el + (unsigned)ix.y * kHid2));
                fma_h2(acc0[0], acc0[1], s0.x, (unsigned short)0x3c00); fma_h2(acc0[2], acc0[3], s0.y, (unsigned short)0x3c00);
                fma_h2(acc0[4], acc0[5], s0.z, (unsigned short)0x3c00); fma_h2(acc0[6], acc0[7], s0.w, (unsigned short)0x3c00);
                fma_h2(acc1[0], acc1[1], s0.x, (unsigned short)0x3c00); fma_h2(acc1[2], acc1[3], s0.y, (unsigned short)0x3c00);
                fma_h2(acc1[4], acc1[5], s0.z, (unsigned short)0x3c00); fma_h2(acc1[6], acc1[7], s0.w, (unsigned short)0x3c00);
                if (i + 1 < cnt) {
                    fma_h2(acc0[0], acc0[1], s1.x, (unsigned short)0x3c00); fma_h2(acc0[2], acc0[3], s1.y, (unsigned short)0x3c00);
                    fma_h2(acc0[4], acc0[5], s1.z, (unsigned short)0x3c00); fma_h2(acc0[6], acc0[7], s1.w, (unsigned short)0x3c00);
                    fma_h2(acc1[0], acc1[1], s1.x, (unsigned short)0x3c00); fma_h2(acc1[2], acc1[3], s1.y, (unsigned short)0x3c00);
                    fma_h2(acc1[4], acc1[5], s1.z, (unsigned short)0x3c00); fma_h2(acc1[6], acc1[7], s1.w, (unsigned short)0x3c00);
                }
            }
            __syncwarp();
            finish_row(e_l, acc0);
            finish_row(ept + e_l, acc1);
        }
    };

    auto do_row = [&](uint4 (&pf)[PLA], const long long tile, const int rr) {
        const int row_l = rr * NW + warp;
        float acc[8];
        {
            const float4 b0 = *reinterpret_cast<const float4 *>(sS->base + 8 * lane);
            const float4 b1 = *reinterpret_cast<const float4 *>(sS->base + 8 * lane + 4);
            acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
            acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
        }
        int n = 0;
        if (row_l < p.rpt && tile * p.rpt + row_l < p.M) {
            const int local = p.a0 + row_l * K;
            const int a = local & 15;
            const int np = (a + K + 15) >> 4;
            const uint8_t *tv = tmplv + (((a - p.a0) & 15) >> p.gshift) * p.tv_stride;
            uint4 t[PLA];
            uint32_t mw[PLA][4];
            int cnt = 0;
#pragma unroll
            for (int g = 0; g < PLA; ++g) {
                const int q = g * 32 + lane;
                t[g] = make_uint4(0u, 0u, 0u, 0u);
                if (q < np) t[g] = *reinterpret_cast<const uint4 *>(tv + 16 * q);
                mw[g][0] = nz_bytes(pf[g].x ^ t[g].x); mw[g][1] = nz_bytes(pf[g].y ^ t[g].y);
                mw[g][2] = nz_bytes(pf[g].z ^ t[g].z); mw[g][3] = nz_bytes(pf[g].w ^ t[g].w);
                cnt += __popc(mw[g][0]) + __popc(mw[g][1]) + __popc(mw[g][2]) + __popc(mw[g][3]);
            }
            if (__ballot_sync(0xffffffffu, cnt != 0) != 0u) {
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += up;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (total <= kListCap) {                 // the usual case: a handful of entries for the whole row
                    int pos = incl - cnt;
                    if (cnt) {
#pragma unroll
                        for (int g = 0; g < PLA; ++g) pos = emit_piece(list, pos, pf[g], t[g], mw[g], 16 * (g * 32 + lane) - a, K);
                    }
                    n = total;
                } else {                                  // a dense row: one piece per lane at a time, draining the list as it fills
#pragma unroll
                    for (int g = 0; g < PLA; ++g) {
                        const int c = __popc(mw[g][0]) + __popc(mw[g][1]) + __popc(mw[g][2]) + __popc(mw[g][3]);
                        int inc2 = c;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const int up = __shfl_up_sync(0xffffffffu, inc2, d);
                            if (lane >= d) inc2 += up;
                        }
                        const int tot2 = __shfl_sync(0xffffffffu, inc2, 31);
                        if (n + tot2 > kListCap) {
                            gather_rows(p.w1, list, n, lane, acc);
                            n = 0;
                        }
                        emit_piece(list, n + inc2 - c, pf[g], t[g], mw[g], 16 * (g * 32 + lane) - a, K);
                        n += tot2;
                        __syncwarp();
                    }
                }
            }
        }
        {   // this slot is free again: fetch the row after next
            const int it2 = item + 2;
            load_row<PLA>(p, (long long)blockIdx.x + (long long)(it2 >> RSH) * gridDim.x, (it2 & (RND - 1)) * NW + warp, lane, pf);
            ++item;
        }
        if (n) gather_rows(p.w1, list, n, lane, acc);
        finish_row(row_l, acc);
    };

    auto mbar_wait = [&](uint32_t bar, uint32_t parity) {
        uint32_t done = 0, spins = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            if (!done && ++spins > (1u << 26)) __trap();      // a lost MMA completion must not hang the GPU
        }
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    };
    // heads: TMEM quadrant (warp & 3), 64-column group (warp >> 2) of accumulator buffer `buf`
    // heads: TMEM lane quadrant (warp & 3) x column group (warp >> 2; NG groups of 128/NG columns per net) of accumulator `buf`
    auto do_heads = [&](const int buf) {
        constexpr int GC = kHid / NG;                        // columns per group
        const int qd = warp & 3, grp = warp >> 2, net = grp / NG, part = grp % NG;
        if (32 * qd >= p.rpt) return;                         // short tiles: these TMEM lanes hold no rows
        const int row_l = 32 * qd + lane;
        const uint32_t taddr = tmem + ((uint32_t)(32 * qd) << 16) + (uint32_t)(buf * kHid2 + net * kHid + part * GC);
        if (net == 0) {
            float pa[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (p.A <= 6) {      // the usual case (Overcooked: 6 actions): row c of hw is {w[c][0..5], b2[c], 0}
#pragma unroll 1
                for (int cc = 0; cc < GC / 16; ++cc) {
                    float v[16];
                    tmem_ld16(taddr + cc * 16, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = part * GC + cc * 16 + i;
                        const float4 wlo = *reinterpret_cast<const float4 *>(sS->hw + 8 * c);
                        const float4 whi = *reinterpret_cast<const float4 *>(sS->hw + 8 * c + 4);
                        const float h = activate<ACT>(v[i] + whi.z);
                        pa[0] = fmaf(h, wlo.x, pa[0]); pa[1] = fmaf(h, wlo.y, pa[1]);
                        pa[2] = fmaf(h, wlo.z, pa[2]); pa[3] = fmaf(h, wlo.w, pa[3]);
                        pa[4] = fmaf(h, whi.x, pa[4]); pa[5] = fmaf(h, whi.y, pa[5]);
                    }
                }
            } else {
#pragma unroll 1
                for (int cc = 0; cc < GC / 16; ++cc) {
                    float v[16];
                    tmem_ld16(taddr + cc * 16, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = part * GC + cc * 16 + i;
                        const float h = activate<ACT>(v[i] + sS->b2[c]);
                        const float4 wlo = *reinterpret_cast<const float4 *>(sS->hw + 8 * c);
                        const float4 whi = *reinterpret_cast<const float4 *>(sS->hw + 8 * c + 4);
                        pa[0] = fmaf(h, wlo.x, pa[0]); pa[1] = fmaf(h, wlo.y, pa[1]);
                        pa[2] = fmaf(h, wlo.z, pa[2]); pa[3] = fmaf(h, wlo.w, pa[3]);
                        pa[4] = fmaf(h, whi.x, pa[4]); pa[5] = fmaf(h, whi.y, pa[5]);
                        pa[6] = fmaf(h, whi.z, pa[6]); pa[7] = fmaf(h, whi.w, pa[7]);
                    }
                }
            }
            float4 *o = reinterpret_cast<float4 *>(hpA + (row_l * NG + part) * 8);
            o[0] = make_float4(pa[0], pa[1], pa[2], pa[3]);
            o[1] = make_float4(pa[4], pa[5], pa[6], pa[7]);
        } else {
            float pc = 0.f;
#pragma unroll 1
            for (int cc = 0; cc < GC / 16; ++cc) {
                float v[16];
                tmem_ld16(taddr + cc * 16, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = part * GC + cc * 16 + i;
                    pc = fmaf(activate<ACT>(v[i] + sS->b2[kHid + c]), sS->cw[c], pc);
                }
            }
            hpC[row_l * NG + part] = pc;
        }
    };
    // finish the rows of `tile`: logits, value, categorical draw (TPR threads per row)
    auto do_finish = [&](const long long tile) {
        constexpr int NA = OC_POLICY_MAX_ACTIONS / TPR;
        const int row_l = tid / TPR, s4 = tid % TPR;
        long long row = (long long)tile * p.rpt + row_l;
        bool live = row_l < p.rpt && row < p.M;
        if (PL == 0) {       // compact source: tile rows are (view, env-in-tile); global rows are agent-major
            const int ept = p.rpt >> 1, view = row_l >= ept ? 1 : 0;
            const long long env = (long long)tile * ept + (row_l - view * ept);
            row = view * p.N + env;
            live = row_l < p.rpt && env < p.N;
        }
        // the row whose noise this row draws, and the size of the noise array, in the single-device program
        long long nrow = p.noise_first + row, ncount = p.noise_global;
        if (PL == 0) {
            const int ept = p.rpt >> 1, view = row_l >= ept ? 1 : 0;
            nrow = view * p.noise_global + p.noise_first + ((long long)tile * ept + (row_l - view * ept));
            ncount = 2 * p.noise_global;
        }
        const int A = p.A;
        float l[NA];
#pragma unroll
        for (int h = 0; h < NA; ++h) {
            const int a = s4 + TPR * h;
            float v = sS->b3[a];
#pragma unroll
            for (int g = 0; g < NG; ++g) v += hpA[(row_l * NG + g) * 8 + a];      // fixed order: deterministic
            l[h] = v;
            if (live && a < A && p.logits) p.logits[row * A + a] = v;
        }
        if (live && s4 == TPR - 1 && p.value) {
            float v = sS->b3c;
#pragma unroll
            for (int g = 0; g < NG; ++g) v += hpC[row_l * NG + g];
            p.value[row] = v;
        }
        if (p.key) {
            int arg;
            float lp;
            sample_lanes<TPR>(p.part, __ldg(p.key), __ldg(p.key + 1), (int)(ncount * A), live ? nrow : 0, A, s4, l, arg, lp);
            if (live && s4 == 0) {
                if (p.action) p.action[row] = arg;
                if (p.log_prob) p.log_prob[row] = lp;
            }
        }
    };

    // Software pipeline over this CTA's tiles: while the tensor cores run layer 2 of tile t into one half of TMEM, the
    // warps do the heads of tile t-1 from the other half and then layer 1 of tile t+1.
    uint32_t phase0 = 0, phase1 = 0;
    bool w2_ready = false;
    long long prev = -1;
    int buf = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, buf ^= 1) {
        if (prev >= 0) {          // layer 2 of the previous tile has finished reading the A tile (and its accumulators are ready)
            if (buf) { mbar_wait(bar_u, phase0); phase0 ^= 1u; } else { mbar_wait(bar_u + 8, phase1); phase1 ^= 1u; }
        }
        // ------------------------------------------------------------------ layer 1: one row per warp, 8 rounds
#pragma unroll 1
        for (int rr = 0; rr < RND; rr += 2) {
            if (PL > 0) {
                do_row(pfA, tile, rr);
                do_row(pfB, tile, rr + 1);
            } else {
                do_env_cells(cbA, tile, rr);
                do_env_cells(cbB, tile, rr + 1);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();

        // ------------------------------------------------------------------ layer 2: 2 nets x 8 K-steps on the tensor cores
        if (tid == 0) {
            if (!w2_ready) { mbar_wait(bar_u + 24, 0u); w2_ready = true; }
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int net = 0; net < 2; ++net) {
#pragma unroll
                for (int ks = 0; ks < kHid / 16; ++ks) {
                    const uint32_t off = net * kNetBytes + (ks >> 2) * (kNetBytes / 2) + (ks & 3) * 32;
                    umma_f16(tmem + buf * kHid2 + net * kHid, umma_desc(sA_u + off), umma_desc(sW2_u + off), ks > 0 ? 1u : 0u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar_u + 8 * buf) : "memory");
        }
        if (prev >= 0) do_heads(buf ^ 1);
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (prev >= 0) do_finish(prev);
        prev = tile;
    }
    if (prev >= 0) {              // drain: the last tile's accumulators are in buffer buf ^ 1
        if (buf) { mbar_wait(bar_u, phase0); } else { mbar_wait(bar_u + 8, phase1); }
        do_heads(buf ^ 1);
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        do_finish(prev);
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// Repack the reference's parameter pytree for the kernel above.
struct PackArgs {
    const float *kernel[6];
    const float *bias[6];
    const uint8_t *tmpl;
    __half *w1;
    uint8_t *w2img;
    PolicySmall *small;
    int K, A;
};

__global__ void oc_policy_pack_kernel(const PackArgs a) {
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)gridDim.x * blockDim.x;
    // layer 1: [K][actor 128 | critic 128]
    for (long long i = gtid; i < (long long)a.K * kHid2; i += gstride) {
        const int k = (int)(i / kHid2), j = (int)(i % kHid2);
        a.w1[i] = __float2half_rn(j < kHid ? a.kernel[0][(size_t)k * kHid + j] : a.kernel[3][(size_t)k * kHid + (j - kHid)]);
    }
    // layer 2: B[n][k] = kernel[k][n] as fp16 in the swizzled K-major image
    __half *img = reinterpret_cast<__half *>(a.w2img);
    for (long long i = gtid; i < 2ll * kHid * kHid; i += gstride) {
        const int net = (int)(i / (kHid * kHid)), r = (int)(i % (kHid * kHid));
        const int n = r / kHid, k = r % kHid;
        const float w = a.kernel[net ? 4 : 1][(size_t)k * kHid + n];
        const int off = net * kNetBytes + (k >> 6) * (kNetBytes / 2) + n * 128 + ((((k & 63) >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
        img[off >> 1] = __float2half_rn(w);
    }
    if (blockIdx.x == 0) {
        for (int j = threadIdx.x; j < kHid2; j += blockDim.x) {
            const float *kern = j < kHid ? a.kernel[0] : a.kernel[3];
            const int jj = j < kHid ? j : j - kHid;
            float s = (j < kHid ? a.bias[0] : a.bias[3])[jj];
            for (int k = 0; k < a.K; ++k) {
                const int t = a.tmpl[k];
                if (t) s = fmaf((float)t, kern[(size_t)k * kHid + jj], s);
            }
            a.small->base[j] = s;
            a.small->b2[j] = (j < kHid ? a.bias[1] : a.bias[4])[jj];
        }
        for (int i = threadIdx.x; i < kHid * 8; i += blockDim.x) {
            const int c = i >> 3, q = i & 7;
            // up to 6 actions: slot 6 of the row carries the layer-2 bias of hidden unit c (one shared-memory read less per column)
            a.small->hw[i] = q < a.A ? a.kernel[2][(size_t)c * a.A + q] : (q == 6 && a.A <= 6) ? a.bias[1][c] : 0.f;
        }
        for (int c = threadIdx.x; c < kHid; c += blockDim.x) a.small->cw[c] = a.kernel[5][c];
        if (threadIdx.x < 8) a.small->b3[threadIdx.x] = threadIdx.x < a.A ? a.bias[2][threadIdx.x] : 0.f;
        if (threadIdx.x == 0) { a.small->b3c = a.bias[5][0]; a.small->pad[0] = a.small->pad[1] = a.small->pad[2] = 0.f; }
    }
}

// Layer-1 contribution of every (cell, code): sum_ch (record[code][ch] - template[cell][ch]) * W1[cell*26 + ch, :] in fp32
// from the fp32 kernels, rounded once to fp16 (records are "as the viewer sees them": agent_1's rows look an agent cell up
// with the viewer bit of its code flipped).  One CTA per table row, one thread per hidden unit.
// Rows C*68 and C*68 + 1: zeros (padding entry) and the urgency plane (channel 25 of every cell).
__global__ void oc_policy_embed_kernel(const float *__restrict__ k_actor, const float *__restrict__ k_critic,
                                       const uint8_t *__restrict__ tmpl, const unsigned *__restrict__ rec, int C, __half *emb) {
    const int row = blockIdx.x, j = threadIdx.x;
    const float *kern = j < kHid ? k_actor : k_critic;
    const int jj = j < kHid ? j : j - kHid;
    const int n_rows = C * OC_NUM_CELL_CODES;
    float s = 0.f;
    if (row < n_rows) {
        const int cell = row / OC_NUM_CELL_CODES, code = row % OC_NUM_CELL_CODES;
        for (int ch = 0; ch < OC_OBS_CHANNELS; ++ch) {
            const int rv = (int)((rec[code * kRecWords + (ch >> 2)] >> (8 * (ch & 3))) & 255u);
            const int d = rv - (int)tmpl[cell * OC_OBS_CHANNELS + ch];
            if (d) s = fmaf((float)d, kern[(size_t)(cell * OC_OBS_CHANNELS + ch) * kHid + jj], s);
        }
    } else if (row == n_rows + 1) {
        for (int cell = 0; cell < C; ++cell) s += kern[(size_t)(cell * OC_OBS_CHANNELS + 25) * kHid + jj];
    }
    emb[(size_t)row * kHid2 + j] = __float2half_rn(s);
}

}  // namespace
