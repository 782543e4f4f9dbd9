// oc_b200.cu -- hand-written sm_100a kernels + the C ABI (include/oc_b200.h) for the batched Overcooked step.
//
// What the kernels restate (reference file:line, relative to jax_marl/environments/):
//   step        multi_agent_env.py:41-69 (key split, auto-reset select) -> overcooked_environment/overcooked.py
//               :106-134 step_env, :366-514 step_agents, :516-642 process_interact, :488-500 pot timers
//   reset       overcooked.py:136-248 + common.py:76-170 make_overcooked_map
//   get_obs     overcooked.py:250-364 (26-channel per-agent observation)
//   LogWrapper  ../wrappers/baselines.py:75-107 (optional epilogue)
//   PRNG        jax.random.split / randint / choice on threefry2x32 (jax is not vendored in the reference;
//               restated from the published algorithm, both `jax_threefry_partitionable` settings)
//
// Design (DESIGN.md has the long form):
//   * one WARP owns a tile of `epw` consecutive environments; lanes 2e / 2e+1 are agent_0 / agent_1 of env e.
//     Movement, collision/swap and the Alice-before-Bob interact ordering are resolved with __shfl_xor_sync
//     between the two lanes of a pair; whole-warp votes skip the reset, conflict and urgency paths.  The grid is
//     persistent and launched with programmatic stream serialization (griddepcontrol), so the next step's
//     prologue overlaps this step's tail.
//   * the only part of maze_map that can ever change is the h x w interior; the warp stages the byte span that
//     covers the interior rows with 16-byte cp.async into double-buffered shared memory (prefetching the next
//     tile), updates it there and mirrors every changed cell to HBM in place.  The constant wall ring is never
//     touched; whole spans are written back only for envs that reset.
//   * observations are the HBM roofline term.  Each cell's 26 channel bytes are a pure function of a small
//     "cell code" (object / pot status / agent x direction x inventory; 68 records), and almost every cell equals
//     the layout's static image.  Per tile the staging buffer is filled with that image (lane copy from shared
//     memory for small images, one TMA bulk load completing on an mbarrier for large ones), the few deviating
//     cells (pots, items on counters, the two agents, urgency) are re-staged, and the tile leaves the SM as one
//     bulk asynchronous copy per agent (cp.async.bulk.global.shared::cta; SASS UBLKCP) -- no per-thread stores.
//   * auto-reset is fused: envs whose episode ended rebuild their interior from a per-layout template and draw
//     direction / position / pot status / inventory with an in-kernel threefry2x32; per-env keys can be derived
//     in the kernel from split(key, num)[first + i] for exactly the envs that reset.
#include "oc_b200.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace {

constexpr int kNumCodes = 68;        // 0..10 objects, 11..34 pot status 0..23, 35..66 agent(which,dir,inv), 67 spare
constexpr int kRecWords = 16;        // [0..7] record at byte offset 0, [8..15] same record shifted by 2 bytes
constexpr int kCodePot = 11;
constexpr int kCodeAgent = 35;
constexpr int kMaxSpanBytes = ((OC_MAX_DIM - 1) * (OC_MAX_DIM + 2 * OC_PAD) + OC_MAX_DIM) * 3;   // 1128
constexpr int kMaxMapBytes = (OC_MAX_DIM + 2 * OC_PAD) * (OC_MAX_DIM + 2 * OC_PAD) * 3;          // 1728
constexpr int kWarpsPerCta = 4;   // default (and maximum) CTA size in warps; very large grids fall back to 2 or 1
constexpr int kMaxTmplBytes = 8 * OC_MAX_CELLS * OC_OBS_CHANNELS;        // worst case: odd cell count -> 8 copies
constexpr int kHdrFixed = 64 + 32;   // wall/blocked bitmaps, pot cells; then the scan list and the observation template

struct DevLayout {
    int h, w, C, WP;
    int map_bytes, span_off, span_len, span_stride;
    int n_pot, n_goal;
    int max_steps, random_reset, task_id, prng_mode;
    int agent_cell[2];
    unsigned magic_C, magic_w;                // ceil(2^32 / d) for d = C, w
    unsigned wall_bits[8], blocked_bits[8];   // wall_map; wall_map | goal cells (movement blocking)
    unsigned short pot_cell[OC_MAX_SPECIAL];
    unsigned short goal_cell[OC_MAX_SPECIAL];
    alignas(16) unsigned rec[kNumCodes][kRecWords];
    unsigned char tmpl_span[kMaxSpanBytes + 8];   // interior span right after a reset, without agents, pots = 23
    unsigned char tmpl_full[kMaxMapBytes];        // whole padded map, same contents (oc_reset initialisation)
    int n_scan, tmpl_k, tmpl_chunks;
    unsigned scan[OC_MAX_CELLS];                  // cells that can deviate from the base image: span offset | cell << 16
    alignas(16) unsigned char obs_tmpl[kMaxTmplBytes];        // tmpl_k back-to-back copies of the layout's static observation
};

struct KParams {
    const DevLayout *L;
    long long N;
    long long n_tiles;
    const uint32_t *keys;
    oc_state st;
    oc_state rs;
    int has_rs;
    const int32_t *act0, *act1;
    const unsigned char *tmpl_tile;   // epw back-to-back copies of the layout's static observation (device memory)
    uint8_t *obs0, *obs1;
    float *reward, *shaped0, *shaped1;
    uint8_t *done;
    oc_step_extras ex;
    int epw;        // environments per warp tile: 16, 8, 4 or 2
    int bulk_ok;    // obs base pointers are 16-byte aligned -> full tiles leave through the bulk-copy engine
    int warp_bytes; // shared memory per warp
    // scalar copies of the layout, so the hot loop reads them from the constant bank
    int h, w, C, WP, map_bytes, span_off, span_len, span_stride, n_pot, max_steps, random_reset, part;
    int agent_cell0, agent_cell1;
    unsigned magic_C, magic_w;
    int n_scan, tmpl_k, tmpl_chunks, tmpl_bytes, scan_bytes, nch, pf_invariant, pf_bytes;
    unsigned magic_nch;
    unsigned magic_ns, magic_tc;
};

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_OBS = 2 };

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

// ------------------------------------------------------------------------------------------ small helpers

__device__ __forceinline__ unsigned fastdiv(unsigned x, unsigned magic) { return __umulhi(x, magic); }
__device__ __forceinline__ int dir_x(int d) { return (d == 2) - (d == 3); }
__device__ __forceinline__ int dir_y(int d) { return (d == 1) - (d == 0); }
__device__ __forceinline__ int obj_colour(int o) { return (int)((0x06716644500ull >> (4 * o)) & 0xFull); }  // common.py:51-63
__device__ __forceinline__ int inv_code(int inv) { return inv == 9 ? 3 : ((inv >> 1) & 3); }               // 1,3,5,9 -> 0..3
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Interact { int obj, status, inv, delivery, shaped, write; };

// overcooked.py:516-642 for one agent facing a cell holding (obj, status); `wall` = wall_map at that cell.
__device__ __forceinline__ Interact interact_cell(int obj, int status, int inv, bool wall) {
    const bool is_pot = obj == 8, is_goal = obj == 7, is_agent = obj == 10;
    const bool is_pile = (obj == 6) | (obj == 4);
    const bool pickable = (obj == 5) | (obj == 3) | (obj == 9);
    const bool is_table = wall & !is_pot;
    const bool table_empty = (obj == 2) | (obj == 1);
    const bool case1 = is_pot & (status > 20) & (inv == 3);      // onion into a pot that is not full
    const bool case2 = is_pot & (status == 0) & (inv == 5);      // plate a finished soup
    const bool pickup = is_table & !table_empty & (inv == 1) & (is_pile | pickable);
    const bool drop = is_table & table_empty & (inv != 1);
    const bool delivery = is_table & is_goal & (inv == 9);
    Interact r;
    r.status = case1 ? status - 1 : (case2 ? 23 : status);
    int ninv = case1 ? 1 : (case2 ? 9 : inv);
    r.obj = obj;
    if (pickup) { r.obj = is_pile ? obj : 2; ninv = (obj == 6) ? 5 : ((obj == 4) ? 3 : obj); }
    if (drop) { r.obj = inv; ninv = 1; }
    if (delivery) ninv = 1;
    r.inv = ninv;
    r.delivery = delivery;
    // :568-569; the plate-pickup bonus (:616-628) is identically zero: the ring's colour byte is 5
    r.shaped = case1 ? 3 : (case2 ? 5 : 0);
    r.write = !is_agent;
    return r;
}

// ------------------------------------------------------------------------------------------------ kernel

template <int MODE>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 8) oc_env_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const DevLayout *__restrict__ L = p.L;
    // ---- CTA-shared header: record table, wall/blocked bitmaps, pot cells, scan list, observation template
    // (the 68-entry record table stays in global memory: it is only touched for the few cells per env that deviate
    //  from the static image, and 4.3 KB of it live in L1)
    const uint32_t *__restrict__ g_rec = &L->rec[0][0];
    uint32_t *s_bits = reinterpret_cast<uint32_t *>(smem);                // [0..7] wall, [8..15] blocked
    unsigned short *s_pot = reinterpret_cast<unsigned short *>(s_bits + 16);
    uint32_t *s_scan = reinterpret_cast<uint32_t *>(smem + kHdrFixed);   // [n_scan] span offset | cell << 16
    uint32_t *s_pf = reinterpret_cast<uint32_t *>(smem + kHdrFixed + p.scan_bytes);   // [epw*nch] prefetch plan
    // small static images are kept in shared memory and copied by the lanes (tmpl_bytes > 0); large ones are pulled
    // per tile from global memory by the bulk-copy engine
    uint4 *s_tmpl = reinterpret_cast<uint4 *>(smem + kHdrFixed + p.scan_bytes + p.pf_bytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = p.h, w = p.w, C = p.C, WP = p.WP;
    const int epw = p.epw;
    const int span_stride = p.span_stride;
    const int part = p.part;
    const int env_bytes = C * OC_OBS_CHANNELS;
    const int view_bytes = epw * env_bytes;
    unsigned char *wbase = smem + kHdrFixed + p.scan_bytes + p.pf_bytes + p.tmpl_bytes + (size_t)warp * p.warp_bytes;
    unsigned char *stage = wbase;                                              // agent_0's view of the tile
    unsigned char *smap_buf = wbase + view_bytes;                              // 2 x [epw][span_stride]
    uint32_t *s_ap = reinterpret_cast<uint32_t *>(smap_buf + 2 * epw * span_stride);   // [2*epw] urgency flags
    const uint32_t tbar = smem_u32(s_ap + 2 * epw);                                    // this warp's mbarrier (8 bytes)

    const int e = lane >> 1, ai = lane & 1;   // env within the tile, agent index (0 = "Alice", 1 = "Bob")
    bool bulk_pending = false;
    // agent_1's view differs from agent_0's only in the two agent cells (overcooked.py:356-359): the staged image is
    // sent to obs0, then -- as soon as the copy engine has READ it -- the two agent cells are re-staged from
    // agent_1's perspective and the same buffer is sent to obs1.
    // write one cell's 26-byte record into the staged image at flat cell index g (2-byte aligned offset g*26)
    auto stage_record = [&](int g, int code) {
        const uint32_t *r = g_rec + code * kRecWords + ((g & 1) ? 8 : 0);
        const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(r)), v1 = __ldg(reinterpret_cast<const uint4 *>(r + 4));
        unsigned char *dst = stage + g * OC_OBS_CHANNELS;
        if (g & 1) {   // offset = 2 (mod 4): shifted record -- upper half of word 0, then words 1..6
            *reinterpret_cast<unsigned short *>(dst) = (unsigned short)(v0.x >> 16);
            uint32_t *d = reinterpret_cast<uint32_t *>(dst + 2);
            d[0] = v0.y; d[1] = v0.z; d[2] = v0.w; d[3] = v1.x; d[4] = v1.y; d[5] = v1.z;
        } else {       // word aligned: words 0..5, then the lower half of word 6
            uint32_t *d = reinterpret_cast<uint32_t *>(dst);
            d[0] = v0.x; d[1] = v0.y; d[2] = v0.z; d[3] = v0.w; d[4] = v1.x; d[5] = v1.y;
            *reinterpret_cast<unsigned short *>(dst + 24) = (unsigned short)(v1.z & 0xffffu);
        }
    };

    // cp.async (LDGSTS) of one tile's interior spans into span buffer `b`, 16 bytes per lane-op: env `el`'s slot is
    // filled from the 16-byte-aligned window that covers its span, so the span itself starts `mis` bytes into the
    // slot, with mis = (address of the span) mod 16.  The few over-read bytes belong to the env's own wall ring.
    const int cta_warps = (int)(blockDim.x >> 5), cta_threads = (int)blockDim.x;
    const long long tile_stride = (long long)gridDim.x * cta_warps;
    auto prefetch_spans = [&](long long tl, int b) {
        const long long e0 = tl * epw;
        const int nv = (int)min((long long)epw, p.N - e0);
        const unsigned char *g0 = p.st.maze_map + e0 * (long long)p.map_bytes + p.span_off;
        const int m0 = (int)(reinterpret_cast<uintptr_t>(g0) & 15);
        const uint32_t dst0 = smem_u32(smap_buf + b * epw * span_stride);
        const int total = nv * p.nch;
        if (p.pf_invariant) {   // every tile has the same alignment: replay the per-CTA copy plan
            for (int f = lane; f < total; f += 32) {
                const uint32_t d = s_pf[f];
                if (d != 0xffffffffu)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                                 :: "r"(dst0 + ((d & 0xffffu) << 4)), "l"(g0 + ((int)(d >> 16) - 16)) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            return;
        }
        for (int f = lane; f < total; f += 32) {
            const int el = (int)fastdiv((unsigned)f, p.magic_nch);
            const int c = f - el * p.nch;
            const int o = el * p.map_bytes;
            const int mis = (m0 + o) & 15;
            if (16 * c < mis + p.span_len)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                             :: "r"(dst0 + (unsigned)(el * span_stride + 16 * c)), "l"(g0 + (o - mis + 16 * c)) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    struct Fields { uint2 pos; unsigned short dir; int di, inv, t, act; };   // raw, unpacked only when consumed
    auto load_fields = [&](long long tl) {
        Fields f;
        f.pos = make_uint2(0u, 0u); f.dir = 0; f.di = 0; f.inv = 1; f.t = 0; f.act = 4;
        const long long e0 = tl * epw;
        const int nv = (int)min((long long)epw, p.N - e0);
        if (e < nv && MODE != MODE_RESET) {
            const long long en = e0 + e;
            f.pos = reinterpret_cast<const uint2 *>(p.st.agent_pos)[en * 2 + ai];
            f.di = p.st.agent_dir_idx[en * 2 + ai];
            f.inv = p.st.agent_inv[en * 2 + ai];
            f.t = p.st.time[en];
            if (MODE == MODE_STEP) {
                f.dir = reinterpret_cast<const unsigned short *>(p.st.agent_dir)[en * 2 + ai];
                f.act = __ldg((ai ? p.act1 : p.act0) + en);
            }
        }
        return f;
    };
    // Programmatic dependent launch: let the next kernel in the stream start its own prologue now, fetch the
    // CTA-shared tables (read-only layout data) while the previous kernel may still be draining, and only then
    // wait for the previous kernel's memory to be visible before touching any environment state.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    {
        if (p.tmpl_bytes) {
            const uint4 *tsrc = reinterpret_cast<const uint4 *>(L->obs_tmpl);
            for (int i = tid; i < p.tmpl_chunks; i += cta_threads) s_tmpl[i] = __ldg(tsrc + i);
        }
        for (int i = tid; i < p.n_scan; i += cta_threads) s_scan[i] = L->scan[i];
        if (p.pf_invariant) {
            const int m0 = (int)((reinterpret_cast<uintptr_t>(p.st.maze_map) + (unsigned)p.span_off) & 15);
            for (int f = tid; f < epw * p.nch; f += cta_threads) {
                const int el = (int)fastdiv((unsigned)f, p.magic_nch);
                const int c = f - el * p.nch;
                const int o = el * p.map_bytes;
                const int mis = (m0 + o) & 15;
                s_pf[f] = (16 * c < mis + p.span_len)
                              ? ((unsigned)(el * (span_stride >> 4) + c) | ((unsigned)(o - mis + 16 * c + 16) << 16))
                              : 0xffffffffu;
            }
        }
        if (tid < 8) { s_bits[tid] = L->wall_bits[tid]; s_bits[8 + tid] = L->blocked_bits[tid]; }
        if (tid < OC_MAX_SPECIAL) s_pot[tid] = L->pot_cell[tid];
    }
    // The static observation image of a tile lives in global memory (L2-resident, tile-sized); each tile's staging
    // buffer is (re)filled from it by ONE bulk asynchronous copy that completes on this warp's mbarrier while the
    // step logic runs.
    uint32_t tphase = 0;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();   // the tables (incl. the prefetch plan) are complete before any warp uses them
    asm volatile("griddepcontrol.wait;" ::: "memory");
    long long tile = (long long)blockIdx.x * cta_warps + warp;
    int buf = 0;
    Fields nf;
    if (tile < p.n_tiles) { prefetch_spans(tile, 0); nf = load_fields(tile); }

    for (; tile < p.n_tiles; tile += tile_stride, buf ^= 1) {
        const long long env0 = tile * epw;
        const int n_valid = (int)min((long long)epw, p.N - env0);
        const bool active = e < n_valid;
        const int es = active ? e : 0;            // idle lanes shadow env 0 of the tile (reads only)
        const long long env = env0 + es;
        // first interior byte of the tile in maze_map; a span sits `mis` bytes into its shared-memory slot
        unsigned char *g_tile = p.st.maze_map + env0 * (long long)p.map_bytes + p.span_off;
        const int m0 = (int)(reinterpret_cast<uintptr_t>(g_tile) & 15);
        unsigned char *smap = smap_buf + buf * epw * span_stride;
        unsigned char *my_span = smap + es * span_stride + ((m0 + es * p.map_bytes) & 15);
        unsigned char *g_span = g_tile + es * p.map_bytes;                                   // same bytes in HBM

        // ------------------------------------------------ (0) refill the staging buffer with the static image
        if (!p.tmpl_bytes) {
            if (lane == 0) {
                if (bulk_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // obs1 of the previous tile
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(tbar), "r"(view_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(stage)), "l"(p.tmpl_tile), "r"(view_bytes), "r"(tbar) : "memory");
            }
            bulk_pending = false;
        }
        // ------------------------------------------------ (1)+(2) this tile's state; prefetch the next tile's
        const Fields cf = nf;
        if (tile + tile_stride < p.n_tiles) {
            prefetch_spans(tile + tile_stride, buf ^ 1);
            nf = load_fields(tile + tile_stride);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        const int px = (int)cf.pos.x, py = (int)cf.pos.y, di = cf.di, inv = cf.inv, t = cf.t;
        const int dx = (int)(signed char)(cf.dir & 0xffu), dy = (int)(signed char)(cf.dir >> 8);
        const int act = clampi(cf.act, 0, 5);
        __syncwarp();

        int nx = px, ny = py, ndi = di, ninv = inv, t_out = t, term_out = 0;
        bool done = (MODE == MODE_RESET);
        int rew = 0, shaped = 0;
        // a cell is updated in the staged span AND, in place, in HBM (only cells that change are ever written back)
        auto put_cell = [&](int off, int a, int b, int c) {
            my_span[off] = (unsigned char)a; my_span[off + 1] = (unsigned char)b; my_span[off + 2] = (unsigned char)c;
            g_span[off] = (unsigned char)a; g_span[off + 1] = (unsigned char)b; g_span[off + 2] = (unsigned char)c;
        };

        if (MODE == MODE_STEP) {
            // ------------------------------------------------------------ movement (overcooked.py:370-431)
            const bool is_move = act < 4;
            const int cx = clampi(px + (is_move ? dir_x(act) : dx), 0, w - 1);
            const int cy = clampi(py + (is_move ? dir_y(act) : dy), 0, h - 1);
            const int cc = cy * w + cx;
            const bool blocked = (s_bits[8 + (cc >> 5)] >> (cc & 31)) & 1u;
            const bool bounced = blocked | !is_move;
            const int tx = bounced ? px : cx, ty = bounced ? py : cy;
            const int otx = __shfl_xor_sync(0xffffffffu, tx, 1), oty = __shfl_xor_sync(0xffffffffu, ty, 1);
            const int opx = __shfl_xor_sync(0xffffffffu, px, 1), opy = __shfl_xor_sync(0xffffffffu, py, 1);
            const bool collision = (tx == otx) & (ty == oty);
            const bool swapped = (tx == opx) & (ty == opy) & (otx == px) & (oty == py);
            const bool stay = collision | swapped;
            nx = stay ? px : tx; ny = stay ? py : ty;
            ndi = is_move ? act : di;                                       // :434 turn even when bounced

            // ------------------------------------------------------------ interact (:437-468, :516-642)
            const int fx = clampi(px + dx, 0, w - 1), fy = clampi(py + dy, 0, h - 1);   // old pos + old dir
            const int fc = fy * w + fx;
            const int foff = (fy * WP + fx) * 3;
            const bool wallbit = (s_bits[fc >> 5] >> (fc & 31)) & 1u;
            const bool do_int = active & (act == 5);
            Interact r = interact_cell(my_span[foff], my_span[foff + 2], inv, wallbit);
            if (do_int && ai == 0 && r.write) put_cell(foff, r.obj, obj_colour(r.obj), r.status);
            __syncwarp();
            const int a_fc = __shfl_sync(0xffffffffu, fc, lane & ~1);
            const int a_did = __shfl_sync(0xffffffffu, (int)do_int, lane & ~1);
            if (do_int && ai == 1) {
                if (a_did && a_fc == fc) r = interact_cell(my_span[foff], my_span[foff + 2], inv, wallbit);   // Bob sees Alice's update
                if (r.write) put_cell(foff, r.obj, obj_colour(r.obj), r.status);
            }
            __syncwarp();
            if (do_int) { ninv = r.inv; rew = r.delivery ? 20 : 0; shaped = r.shaped; }

            // ------------------------------------------------------------ redraw agents (:470-486)
            if (active) put_cell((py * WP + px) * 3, 1, 0, 0);
            __syncwarp();
            if (active) put_cell((ny * WP + nx) * 3, 10, 2 * ai, ndi);
            __syncwarp();
            // ------------------------------------------------------------ pot timers (:488-500)
            if (active) {
                for (int q = ai; q < p.n_pot; q += 2) {
                    const int pc = s_pot[q];
                    const int pyy = (int)fastdiv((unsigned)pc, p.magic_w), pxx = pc - pyy * w;
                    const int off = (pyy * WP + pxx) * 3 + 2;
                    const int s = my_span[off];
                    if (s >= 1 && s <= 20) { my_span[off] = (unsigned char)(s - 1); g_span[off] = (unsigned char)(s - 1); }
                }
            }
            rew += __shfl_xor_sync(0xffffffffu, rew, 1);                   // :502 reward = alice + bob
            t_out = t + 1;                                                  // :118
            done = t_out >= p.max_steps;                                    // :120, :644-647
            __syncwarp();
        }

        // -------------------------------------------------------------------- (3) reset (overcooked.py:136-248)
        bool full_writeback = (MODE == MODE_RESET);
        if (MODE != MODE_OBS) {
            const unsigned done_mask = __ballot_sync(0xffffffffu, done && active);
            if (done_mask) {
                full_writeback = true;
                if (MODE == MODE_STEP && p.has_rs) {
                    // multi_agent_env.py:55-59: caller-supplied reset_state replaces reset(key_reset)
                    const unsigned char *rt = p.rs.maze_map + env0 * (long long)p.map_bytes + p.span_off;
                    for (int el = 0; el < n_valid; ++el) {
                        if (!((done_mask >> (2 * el)) & 1u)) continue;
                        unsigned char *dst = smap + el * span_stride + ((m0 + el * p.map_bytes) & 15);
                        for (int k = lane; k < p.span_len; k += 32) dst[k] = __ldg(rt + el * p.map_bytes + k);
                    }
                    if (done && active) {
                        const uint2 pp = __ldg(reinterpret_cast<const uint2 *>(p.rs.agent_pos) + env * 2 + ai);
                        nx = (int)pp.x; ny = (int)pp.y;
                        ndi = __ldg(p.rs.agent_dir_idx + env * 2 + ai);
                        ninv = __ldg(p.rs.agent_inv + env * 2 + ai);
                        t_out = __ldg(p.rs.time + env);
                        term_out = __ldg(p.rs.terminal + env);
                    }
                    __syncwarp();
                } else {
                    // template interior for every finished env (warp-cooperative byte copy)
                    for (int el = 0; el < n_valid; ++el) {
                        if (!((done_mask >> (2 * el)) & 1u)) continue;
                        unsigned char *dst = smap + el * span_stride + ((m0 + el * p.map_bytes) & 15);
                        for (int k = lane; k < p.span_len; k += 32) dst[k] = __ldg(L->tmpl_span + k);
                    }
                    __syncwarp();
                    if (done && active) {
                        uint32_t k0, k1;
                        if (p.keys) { k0 = __ldg(p.keys + env * 2); k1 = __ldg(p.keys + env * 2 + 1); }
                        else {   // the caller's split(key, num_envs) fused in: derive this env's row on demand
                            split_row(part, __ldg(p.ex.split_key), __ldg(p.ex.split_key + 1), p.ex.split_num,
                                      p.ex.split_first + env, k0, k1);
                        }
                        uint32_t s0, s1;
                        if (MODE == MODE_STEP) { split2(part, k0, k1, s0, s1); k0 = s0; k1 = s1; }   // key_reset
                        uint32_t a0, a1, d0, d1, q0, q1, i0, i1;
                        split2(part, k0, k1, a0, a1);      // :164 agent cells
                        split2(part, k0, k1, d0, d1);      // :173 directions
                        split2(part, k0, k1, q0, q1);      // :197 pot status
                        split2(part, k0, k1, i0, i1);      // :225 inventories
                        ndi = randint_at(part, d0, d1, 2, ai, 4u);                      // :174
                        int cell = ai ? p.agent_cell1 : p.agent_cell0;
                        ninv = 1;
                        if (p.random_reset) {
                            // :165-169 choice(p = free cells, replace=False): the two largest 23-bit mantissas,
                            // ties to the lower cell index
                            unsigned long long b1 = 0ull, b2 = 0ull;
                            const int m = (C + 1) >> 1;
                            const int nblk = part ? C : m;
                            for (int j = ai; j < nblk; j += 2) {
                                uint32_t y0, y1;
                                int c0, c1;
                                if (part) { tf2x32(a0, a1, 0u, (uint32_t)j, y0, y1); y0 ^= y1; c0 = j; c1 = -1; }
                                else { tf2x32(a0, a1, (uint32_t)j, (j + m < C) ? (uint32_t)(j + m) : 0u, y0, y1); c0 = j; c1 = (j + m < C) ? j + m : -1; }
                                #pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    const int c = u ? c1 : c0;
                                    const uint32_t bits = u ? y1 : y0;
                                    if (c < 0) continue;
                                    if ((s_bits[c >> 5] >> (c & 31)) & 1u) continue;   // walls have p = 0
                                    const unsigned long long key = ((unsigned long long)((bits >> 9) + 1u) << 32) | (uint32_t)(~(uint32_t)c);
                                    if (key > b1) { b2 = b1; b1 = key; } else if (key > b2) { b2 = key; }
                                }
                            }
                            const unsigned long long o1 = __shfl_xor_sync(done_mask, b1, 1);
                            const unsigned long long o2 = __shfl_xor_sync(done_mask, b2, 1);
                            const unsigned long long first = b1 > o1 ? b1 : o1;
                            const unsigned long long second = b1 > o1 ? (b2 > o1 ? b2 : o1) : (o2 > b1 ? o2 : b1);
                            cell = (int)(~(uint32_t)((ai ? second : first) & 0xffffffffull));
                            const int it = randint_at(part, i0, i1, 2, ai, 4u);         // :226-230: [1,3,5,9][it]
                            ninv = it == 3 ? 9 : 2 * it + 1;
                            for (int q = ai; q < p.n_pot; q += 2) {                     // :200
                                const int pc = s_pot[q];
                                const int pyy = (int)fastdiv((unsigned)pc, p.magic_w), pxx = pc - pyy * w;
                                my_span[(pyy * WP + pxx) * 3 + 2] = (unsigned char)randint_at(part, q0, q1, p.n_pot, q, 24u);
                            }
                        }
                        ny = (int)fastdiv((unsigned)cell, p.magic_w); nx = cell - ny * w;    // :170
                        unsigned char *c1 = my_span + (ny * WP + nx) * 3;                    // common.py:99-105
                        c1[0] = 10; c1[1] = (unsigned char)(2 * ai); c1[2] = (unsigned char)ndi;
                        t_out = 0;
                    }
                    __syncwarp();
                }
            }
        }

        // LogWrapper trackers of this agent: requested here so the loads overlap the observation phase
        float lw_er = 0.f, lw_el = 0.f, lw_rr = 0.f, lw_rl = 0.f;
        if (MODE == MODE_STEP && p.ex.episode_returns && active) {
            const long long j = env * 2 + ai;
            lw_er = p.ex.episode_returns[j]; lw_el = p.ex.episode_lengths[j];
            lw_rr = p.ex.returned_episode_returns[j]; lw_rl = p.ex.returned_episode_lengths[j];
        }

        // ---------------------------------------------------------------- (4) observations (overcooked.py:250-364)
        bool obs1_pending = false;
        int code_v1 = 0, my_cell = 0, urg_lane = 0;
        const size_t obs_off = (size_t)env0 * env_bytes;
        {
            const int urg = (p.max_steps - t_out) < 40;                                       // :311
            if (active) s_ap[lane] = (uint32_t)urg;
            // (4a) base image
            if (p.tmpl_bytes) {   // copied by the lanes from shared memory, 16 bytes per lane-step
                if (bulk_pending) {   // the previous tile's obs1 copy must have finished READING the staging buffer
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    bulk_pending = false;
                    __syncwarp();
                }
                const int n_chunks = (n_valid * env_bytes + 15) >> 4;
                uint4 *dst = reinterpret_cast<uint4 *>(stage);
                int src = lane;                                                         // tiles start on a template period
                src -= (int)fastdiv((unsigned)src, p.magic_tc) * p.tmpl_chunks;
                const int step = 32 - (int)fastdiv(32u, p.magic_tc) * p.tmpl_chunks;      // 32 mod tmpl_chunks
                for (int f = lane; f < n_chunks; f += 32) {
                    dst[f] = s_tmpl[src];
                    src += step;
                    if (src >= p.tmpl_chunks) src -= p.tmpl_chunks;
                }
            } else {              // wait for the bulk copy issued at the top of the tile
                uint32_t ok = 0;
                while (!ok) {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(tbar), "r"(tphase) : "memory");
                }
                tphase ^= 1u;
            }
            __syncwarp();
            // (4b) cells that can deviate from the base image: counters next to a walkable cell (loose items) and pots
            {
                const int total = n_valid * p.n_scan;
                for (int f = lane; f < total; f += 32) {
                    const int el = (int)fastdiv((unsigned)f, p.magic_ns);
                    const uint32_t sc = s_scan[f - el * p.n_scan];
                    const unsigned char *c = smap + el * span_stride + ((m0 + el * p.map_bytes) & 15) + (int)(sc & 0xffffu);
                    const int o = c[0], st = c[2];
                    if (o == 8 ? (st != 23) : (o != 2 && o != 10))
                        stage_record(el * C + (int)(sc >> 16), (o == 8) ? kCodePot + min(st, 23) : min(o, 10));
                }
            }
            __syncwarp();
            // (4c) agent cells show the agent, its direction and what it holds (:313-323, :345-353)
            if (active) stage_record(es * C + ny * w + nx, kCodeAgent + 16 * ai + (ndi | (inv_code(ninv) << 2)));
            __syncwarp();
            // (4d) urgency plane (:311, :341): channel 25 of every cell of an env with < 40 steps left
            {
                const unsigned urg_mask = __ballot_sync(0xffffffffu, active && urg && ai == 0);
                if (urg_mask) {
                    const int n_cells = n_valid * C;
                    for (int g = lane; g < n_cells; g += 32) {
                        const int el = (int)fastdiv((unsigned)g, p.magic_C);
                        if (s_ap[2 * el]) stage[g * OC_OBS_CHANNELS + 25] = 1;
                    }
                }
            }
            // this lane's agent cell as agent_1 sees it: "which" flipped (agent_1 is the viewer)
            code_v1 = kCodeAgent + (ai ? 0 : 16) + (ndi | (inv_code(ninv) << 2));
            my_cell = es * C + ny * w + nx;
            urg_lane = urg;
            if (p.bulk_ok && n_valid == epw) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 :: "l"(p.obs0 + obs_off), "r"(smem_u32(stage)), "r"(view_bytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                obs1_pending = true;      // issued after the state write-back below, which hides the engine's read
            } else {
                // ragged last tile (or unaligned caller buffers): plain stores, 2-byte granular
                __syncwarp();
                const int nb = n_valid * env_bytes;   // even
                const unsigned short *h0 = reinterpret_cast<const unsigned short *>(stage);
                unsigned short *g0 = reinterpret_cast<unsigned short *>(p.obs0 + obs_off);
                unsigned short *g1 = reinterpret_cast<unsigned short *>(p.obs1 + obs_off);
                for (int k = lane; k < (nb >> 1); k += 32) g0[k] = h0[k];
                __syncwarp();
                if (active) stage_record(my_cell, code_v1);
                if (urg && active) stage[my_cell * OC_OBS_CHANNELS + 25] = 1;
                __syncwarp();
                for (int k = lane; k < (nb >> 1); k += 32) g1[k] = h0[k];
            }
        }

        // ---------------------------------------------------------------- (5) write the state back, in place
        if (MODE != MODE_OBS) {
            __syncwarp();
            if (full_writeback) {   // a reset rewrote whole interiors: store the staged spans (rare: 1 step in max_steps)
                for (int el = 0; el < n_valid; ++el) {
                    const int mis = (m0 + el * p.map_bytes) & 15;
                    const unsigned char *src = smap + el * span_stride;          // slot and HBM agree modulo 16
                    unsigned char *dstg = g_tile + el * p.map_bytes - mis;
                    for (int o = (mis & ~3) + 4 * lane; o < mis + p.span_len; o += 128)
                        *reinterpret_cast<uint32_t *>(dstg + o) = *reinterpret_cast<const uint32_t *>(src + o);
                }
            }
            if (active) {
                reinterpret_cast<uint2 *>(p.st.agent_pos)[env * 2 + ai] = make_uint2((uint32_t)nx, (uint32_t)ny);
                reinterpret_cast<char2 *>(p.st.agent_dir)[env * 2 + ai] = make_char2((signed char)dir_x(ndi), (signed char)dir_y(ndi));
                p.st.agent_dir_idx[env * 2 + ai] = ndi;
                p.st.agent_inv[env * 2 + ai] = ninv;
                if (ai == 0) {
                    p.st.time[env] = t_out;
                    p.st.terminal[env] = (uint8_t)term_out;   // stepped: done is False; reset: False (overcooked.py:242)
                }
            }
            if (MODE == MODE_STEP) {
                int st_rr = 0, st_rl = 0;
                if (active) {
                    if (ai == 0) { p.reward[env] = (float)rew; p.done[env] = (uint8_t)done; }
                    (ai ? p.shaped1 : p.shaped0)[env] = (float)shaped;
                    if (p.ex.episode_returns) {   // wrappers/baselines.py:90-106, this lane = agent ai
                        const long long j = env * 2 + ai;
                        // pin the four loaded trackers here: without this the compiler hoists their first use right
                        // behind the loads and the warp stalls on DRAM before the observation phase
                        asm volatile("" : "+f"(lw_er), "+f"(lw_el), "+f"(lw_rr), "+f"(lw_rl));
                        const float keep = done ? 0.f : 1.f, dn = done ? 1.f : 0.f;
                        const float nr = lw_er + (float)rew;
                        const float nl = lw_el + 1.f;
                        p.ex.episode_returns[j] = nr * keep;
                        p.ex.episode_lengths[j] = nl * keep;
                        const float rr = lw_rr * keep + nr * dn;
                        const float rl = lw_rl * keep + nl * dn;
                        p.ex.returned_episode_returns[j] = rr;
                        p.ex.returned_episode_lengths[j] = rl;
                        if (p.ex.returned_episode) p.ex.returned_episode[j] = (uint8_t)done;
                        if (ai == 0 && done) { st_rr = (int)rr; st_rl = (int)rl; }
                    } else if (p.ex.returned_episode) {
                        p.ex.returned_episode[env * 2 + ai] = (uint8_t)done;
                    }
                }
                if (p.ex.stats) {   // exact integer sums: warp-reduced per tile, one atomic per non-zero counter
                    int v_done = (active && ai == 0 && done) ? 1 : 0;
                    int v_rew = (active && ai == 0) ? rew : 0;
                    int v_s0 = (active && ai == 0) ? shaped : 0;
                    int v_s1 = (active && ai == 1) ? shaped : 0;
                    #pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        v_done += __shfl_xor_sync(0xffffffffu, v_done, o);
                        v_rew += __shfl_xor_sync(0xffffffffu, v_rew, o);
                        v_s0 += __shfl_xor_sync(0xffffffffu, v_s0, o);
                        v_s1 += __shfl_xor_sync(0xffffffffu, v_s1, o);
                        st_rr += __shfl_xor_sync(0xffffffffu, st_rr, o);
                        st_rl += __shfl_xor_sync(0xffffffffu, st_rl, o);
                    }
                    if (lane == 0) {
                        unsigned long long *sp = reinterpret_cast<unsigned long long *>(p.ex.stats);
                        if (v_done) atomicAdd(sp + 0, (unsigned long long)v_done);
                        if (v_rew) atomicAdd(sp + 1, (unsigned long long)v_rew);
                        if (v_s0) atomicAdd(sp + 2, (unsigned long long)v_s0);
                        if (v_s1) atomicAdd(sp + 3, (unsigned long long)v_s1);
                        if (st_rr) atomicAdd(sp + 4, (unsigned long long)st_rr);
                        if (st_rl) atomicAdd(sp + 5, (unsigned long long)st_rl);
                    }
                }
            }
        }
        if (obs1_pending) {   // agent_1's view: re-stage the two agent cells once the engine has read the image
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            if (active) stage_record(my_cell, code_v1);
            if (urg_lane && active) stage[my_cell * OC_OBS_CHANNELS + 25] = 1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(p.obs1 + obs_off), "r"(smem_u32(stage)), "r"(view_bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            bulk_pending = true;
        }
        __syncwarp();
    }
    if (MODE == MODE_STEP && p.ex.stats && blockIdx.x == 0 && tid == 0)   // every env stepped once
        atomicAdd(reinterpret_cast<unsigned long long *>(p.ex.stats) + 6, (unsigned long long)p.N);
    if (bulk_pending && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}


// oc_reset pre-pass: constant ring + static State fields (the reset kernel then fills interior and dynamics)
__global__ void oc_init_static_kernel(const DevLayout *__restrict__ L, long long N, oc_state st) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int mb = L->map_bytes, C = L->C, w = L->w;
    for (long long i = tid; i < N * mb; i += stride) st.maze_map[i] = L->tmpl_full[(int)(i % mb)];
    for (long long i = tid; i < N * C; i += stride) {
        const int c = (int)(i % C);
        st.wall_map[i] = (uint8_t)((L->wall_bits[c >> 5] >> (c & 31)) & 1u);
    }
    for (long long i = tid; i < N * L->n_goal; i += stride) {
        const int c = L->goal_cell[(int)(i % L->n_goal)];
        st.goal_pos[2 * i] = (uint32_t)(c % w); st.goal_pos[2 * i + 1] = (uint32_t)(c / w);
    }
    for (long long i = tid; i < N * L->n_pot; i += stride) {
        const int c = L->pot_cell[(int)(i % L->n_pot)];
        st.pot_pos[2 * i] = (uint32_t)(c % w); st.pot_pos[2 * i + 1] = (uint32_t)(c / w);
    }
    for (long long i = tid; i < N; i += stride) st.task_id[i] = L->task_id;
}

// rows [first, first+count) of jax.random.split(key, num)
__global__ void oc_split_kernel(int part, const uint32_t *__restrict__ key, long long num, long long first,
                                long long count, uint32_t *__restrict__ out) {
    const uint32_t k0 = __ldg(key), k1 = __ldg(key + 1);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < count; r += stride) {
        uint32_t a, b;
        split_row(part, k0, k1, num, first + r, a, b);
        out[2 * r] = a; out[2 * r + 1] = b;
    }
}

// n sequential `key, sub = split(key)` (the callers' per-step `rng, _rng = jax.random.split(rng)`), one thread
__global__ void oc_chain_kernel(int part, uint32_t *__restrict__ key, int n, uint32_t *__restrict__ subs) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t k0 = key[0], k1 = key[1];
    for (int i = 0; i < n; ++i) {
        uint32_t s0, s1;
        split2(part, k0, k1, s0, s1);
        subs[2 * i] = s0; subs[2 * i + 1] = s1;
    }
    key[0] = k0; key[1] = k1;
}

}  // namespace

// ================================================================================================ host side

struct oc_layout {
    DevLayout host;
    DevLayout *dev;
    int device;
    int epw;
    int smem_bytes;
    int warp_bytes;
    int warps;                 // warps per CTA: 4 unless the layout's tiles are too big for shared memory
    int pf_bytes;
    int tmpl_smem_bytes;       // > 0: the static image (tmpl_k copies) is kept in shared memory
    unsigned char *tmpl_tile;
    int num_sms;
    mutable int occ[3];   // resident CTAs per SM of each kernel mode (filled on first launch)
};

namespace {


unsigned magic_of(unsigned d) { return (unsigned)((0x100000000ull + d - 1) / d); }

// one cell's 26 observation bytes for a given cell code, from agent_0's... "viewer's" perspective
void fill_record(int code, unsigned char v[26]) {
    std::memset(v, 0, 26);
    auto env_obj = [&](int o) {   // overcooked.py:326-340 on o' (object id, or the inventory on an agent cell)
        if (o == 8) v[10] = 1;
        if (o == 2) v[11] = 1;
        if (o == 4) v[12] = 1;
        if (o == 6) v[14] = 1;
        if (o == 7) v[15] = 1;
        if (o == 5) v[22] = 1;
        if (o == 3) v[23] = 1;
    };
    if (code < kCodePot) {
        env_obj(code);
        if (code == 9) { v[18] = 3; v[21] = 1; }          // dish on a counter: soup_loc (:302, :308, :310)
    } else if (code < kCodeAgent) {
        const int s = code - kCodePot;                      // pot with status s
        v[10] = 1;
        if (s >= 20) v[16] = (unsigned char)((23 - s) < 3 ? (23 - s) : 3);   // :306
        else { v[18] = 3; v[20] = (unsigned char)s; v[21] = (s == 0); }       // :307-310
    } else if (code < kCodeAgent + 32) {
        const int k = code - kCodeAgent, which = k >> 4, dir = k & 3, ic = (k >> 2) & 3;
        static const int items[4] = {1, 3, 5, 9};
        v[which] = 1;                                       // :351 / :357
        v[2 + 4 * which + dir] = 1;                         // :345-347, :353 / :358
        env_obj(items[ic]);
        if (items[ic] == 9) { v[18] = 3; v[21] = 1; }       // held dish (:320-323)
    }
}

int fail_cuda() { return OC_ERR_CUDA; }

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool state_ok(const oc_state &s) {
    return s.agent_pos && s.agent_dir && s.agent_dir_idx && s.agent_inv && s.goal_pos && s.pot_pos && s.wall_map &&
           s.maze_map && s.time && s.terminal && s.task_id;
}

bool aligned_to(const void *p, unsigned a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1u)) == 0; }

// what the kernels' vector accesses need: agent_pos is read as uint2, maze_map as 32-bit words, agent_dir as char2
bool state_aligned(const oc_state &s) {
    return aligned_to(s.agent_pos, 8) && aligned_to(s.agent_dir, 2) && aligned_to(s.agent_dir_idx, 4) &&
           aligned_to(s.agent_inv, 4) && aligned_to(s.maze_map, 4) && aligned_to(s.time, 4) &&
           aligned_to(s.goal_pos, 4) && aligned_to(s.pot_pos, 4) && aligned_to(s.task_id, 4);
}

template <int MODE>
int launch_env(const oc_layout *lay, KParams &kp, cudaStream_t stream) {
    const DevLayout &H = lay->host;
    kp.L = lay->dev;
    kp.tmpl_tile = lay->tmpl_tile;
    kp.epw = lay->epw;
    kp.warp_bytes = lay->warp_bytes;
    kp.h = H.h; kp.w = H.w; kp.C = H.C; kp.WP = H.WP; kp.map_bytes = H.map_bytes; kp.span_off = H.span_off;
    kp.span_len = H.span_len; kp.span_stride = H.span_stride; kp.n_pot = H.n_pot;
    kp.max_steps = H.max_steps; kp.random_reset = H.random_reset; kp.part = H.prng_mode;
    kp.agent_cell0 = H.agent_cell[0]; kp.agent_cell1 = H.agent_cell[1];
    kp.magic_C = H.magic_C; kp.magic_w = H.magic_w;
    kp.n_scan = H.n_scan; kp.tmpl_k = H.tmpl_k; kp.tmpl_chunks = H.tmpl_chunks; kp.tmpl_bytes = lay->tmpl_smem_bytes;
    kp.scan_bytes = (H.n_scan * 4 + 15) & ~15;
    kp.nch = H.span_stride / 16; kp.magic_nch = magic_of((unsigned)kp.nch);
    kp.pf_bytes = lay->pf_bytes;
    kp.pf_invariant = (((long long)lay->epw * H.map_bytes) % 16 == 0 && (long long)lay->epw * H.map_bytes + 32 < 65536) ? 1 : 0;
    kp.magic_ns = magic_of((unsigned)(H.n_scan > 0 ? H.n_scan : 1)); kp.magic_tc = magic_of((unsigned)H.tmpl_chunks);
    kp.n_tiles = (kp.N + lay->epw - 1) / lay->epw;
    if (kp.n_tiles < 1) return OC_OK;
    if (lay->occ[MODE] == 0) {
        if (cudaFuncSetAttribute(oc_env_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return fail_cuda();
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, oc_env_kernel<MODE>, lay->warps * 32, lay->smem_bytes) != cudaSuccess)
            return fail_cuda();
        lay->occ[MODE] = nb < 1 ? 1 : nb;
    }
    // persistent grid: every resident warp gets the same number of tiles (k), so there is no ragged last wave
    const long long resident = (long long)lay->num_sms * lay->occ[MODE];
    const long long want = (kp.n_tiles + lay->warps - 1) / lay->warps;
    long long ctas = want;
    if (want > resident) {
        const long long k = (want + resident - 1) / resident;
        ctas = (want + k - 1) / k;
    }
    static const bool use_pdl = [] { const char *e = std::getenv("OC_B200_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3((unsigned)lay->warps * 32);
    cfg.dynamicSmemBytes = (size_t)lay->smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, oc_env_kernel<MODE>, kp) == cudaSuccess ? OC_OK : fail_cuda();
}

}  // namespace

extern "C" {

int oc_version(void) { return OC_B200_VERSION; }

const char *oc_status_string(int s) {
    switch (s) {
        case OC_OK: return "ok";
        case OC_ERR_INVALID_ARG: return "invalid argument";
        case OC_ERR_UNSUPPORTED_LAYOUT: return "unsupported layout";
        case OC_ERR_CUDA: return "CUDA runtime error";
        case OC_ERR_ALIGNMENT: return "pointer not aligned for its element type";
        case OC_ERR_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
        default: return "unknown status";
    }
}

int oc_layout_create(const oc_layout_desc *d, oc_layout **out) {
    if (!d || !out) return OC_ERR_INVALID_ARG;
    *out = nullptr;
    const int h = d->height, w = d->width;
    if (h < 1 || w < 1 || h > OC_MAX_DIM || w > OC_MAX_DIM) return OC_ERR_UNSUPPORTED_LAYOUT;
    const int C = h * w;
    if (d->n_pot < 1 || d->n_pot > OC_MAX_SPECIAL || d->n_goal < 0 || d->n_goal > OC_MAX_SPECIAL ||
        d->n_onion_pile < 0 || d->n_onion_pile > OC_MAX_SPECIAL || d->n_plate_pile < 0 ||
        d->n_plate_pile > OC_MAX_SPECIAL || d->n_wall < 0 || d->max_steps < 1)
        return OC_ERR_UNSUPPORTED_LAYOUT;
    if (!d->agent_idx || !d->pot_idx || (d->n_wall && !d->wall_idx) || (d->n_goal && !d->goal_idx) ||
        (d->n_onion_pile && !d->onion_pile_idx) || (d->n_plate_pile && !d->plate_pile_idx))
        return OC_ERR_INVALID_ARG;
    if (d->prng_mode != OC_PRNG_THREEFRY_LEGACY && d->prng_mode != OC_PRNG_THREEFRY_PARTITIONABLE)
        return OC_ERR_INVALID_ARG;

    std::vector<unsigned char> wall(C, 0);
    for (int i = 0; i < d->n_wall; ++i) {
        const int c = d->wall_idx[i];
        if (c < 0 || c >= C) return OC_ERR_UNSUPPORTED_LAYOUT;
        wall[c] = 1;
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            if ((y == 0 || y == h - 1 || x == 0 || x == w - 1) && !wall[y * w + x]) return OC_ERR_UNSUPPORTED_LAYOUT;
    auto special_ok = [&](const int32_t *idx, int n) {
        for (int i = 0; i < n; ++i)
            if (idx[i] < 0 || idx[i] >= C || !wall[idx[i]]) return false;
        return true;
    };
    if (!special_ok(d->goal_idx, d->n_goal) || !special_ok(d->pot_idx, d->n_pot) ||
        !special_ok(d->onion_pile_idx, d->n_onion_pile) || !special_ok(d->plate_pile_idx, d->n_plate_pile))
        return OC_ERR_UNSUPPORTED_LAYOUT;
    for (int i = 0; i < 2; ++i)
        if (d->agent_idx[i] < 0 || d->agent_idx[i] >= C || wall[d->agent_idx[i]]) return OC_ERR_UNSUPPORTED_LAYOUT;
    if (d->agent_idx[0] == d->agent_idx[1]) return OC_ERR_UNSUPPORTED_LAYOUT;
    int n_free = 0;
    for (int c = 0; c < C; ++c) n_free += !wall[c];
    if (n_free < 2) return OC_ERR_UNSUPPORTED_LAYOUT;

    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return OC_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return OC_ERR_NO_DEVICE;
    if (prop.major < 10) return OC_ERR_NO_DEVICE;

    oc_layout *lay = new (std::nothrow) oc_layout();
    if (!lay) return OC_ERR_INVALID_ARG;
    DevLayout &L = lay->host;
    std::memset(&L, 0, sizeof(L));
    L.h = h; L.w = w; L.C = C; L.WP = w + 2 * OC_PAD;
    L.map_bytes = (h + 2 * OC_PAD) * L.WP * 3;
    L.span_off = (OC_PAD * L.WP + OC_PAD) * 3;
    L.span_len = ((h - 1) * L.WP + w) * 3;
    L.span_stride = ((L.span_len + 15 + 15) / 16) * 16;   // shared-memory slot: span + up to 15 bytes of misalignment
    L.n_pot = d->n_pot; L.n_goal = d->n_goal;
    L.max_steps = d->max_steps; L.random_reset = d->random_reset ? 1 : 0; L.task_id = d->task_id;
    L.prng_mode = d->prng_mode;
    L.agent_cell[0] = d->agent_idx[0]; L.agent_cell[1] = d->agent_idx[1];
    L.magic_C = magic_of((unsigned)C); L.magic_w = magic_of((unsigned)w);
    for (int c = 0; c < C; ++c)
        if (wall[c]) { L.wall_bits[c >> 5] |= 1u << (c & 31); L.blocked_bits[c >> 5] |= 1u << (c & 31); }
    for (int i = 0; i < d->n_goal; ++i) {
        L.goal_cell[i] = (unsigned short)d->goal_idx[i];
        L.blocked_bits[d->goal_idx[i] >> 5] |= 1u << (d->goal_idx[i] & 31);   // overcooked.py:381-391
    }
    for (int i = 0; i < d->n_pot; ++i) L.pot_cell[i] = (unsigned short)d->pot_idx[i];

    // observation record table
    for (int code = 0; code < kNumCodes; ++code) {
        unsigned char v[32];
        std::memset(v, 0, sizeof(v));
        fill_record(code, v);
        unsigned char a[32], b[32];
        std::memset(a, 0, 32); std::memset(b, 0, 32);
        std::memcpy(a, v, 26);
        std::memcpy(b + 2, v, 26);
        std::memcpy(&L.rec[code][0], a, 32);
        std::memcpy(&L.rec[code][8], b, 32);
    }
    // post-reset map template (common.py:76-170), agents left out; pots at POT_EMPTY_STATUS
    {
        const int HP = h + 2 * OC_PAD, WP = L.WP;
        unsigned char *m = L.tmpl_full;
        for (int i = 0; i < HP * WP; ++i) { m[3 * i] = 2; m[3 * i + 1] = 5; m[3 * i + 2] = 0; }
        auto put = [&](int c, int o, int col, int st) {
            unsigned char *q = m + ((OC_PAD + c / w) * WP + OC_PAD + c % w) * 3;
            q[0] = (unsigned char)o; q[1] = (unsigned char)col; q[2] = (unsigned char)st;
        };
        for (int c = 0; c < C; ++c) if (!wall[c]) put(c, 1, 0, 0);
        for (int i = 0; i < d->n_goal; ++i) put(d->goal_idx[i], 7, 1, 0);
        for (int i = 0; i < d->n_onion_pile; ++i) put(d->onion_pile_idx[i], 4, 4, 0);
        for (int i = 0; i < d->n_plate_pile; ++i) put(d->plate_pile_idx[i], 6, 6, 0);
        for (int i = 0; i < d->n_pot; ++i) put(d->pot_idx[i], 8, 7, 23);
        std::memcpy(L.tmpl_span, m + L.span_off, (size_t)L.span_len);
        // cells whose observation can deviate from the static image without being an agent cell: pots, and plain
        // counters an agent can face (4-neighbours of a walkable cell) -- the only cells interact ever rewrites
        L.n_scan = 0;
        for (int c = 0; c < C; ++c) {
            const int y = c / w, x = c % w;
            const unsigned char *q = m + ((OC_PAD + y) * WP + OC_PAD + x) * 3;
            bool near_free = false;
            if (y > 0 && !wall[c - w]) near_free = true;
            if (y < h - 1 && !wall[c + w]) near_free = true;
            if (x > 0 && !wall[c - 1]) near_free = true;
            if (x < w - 1 && !wall[c + 1]) near_free = true;
            if (q[0] == 8 || (q[0] == 2 && wall[c] && near_free))
                L.scan[L.n_scan++] = (unsigned)((y * WP + x) * 3) | ((unsigned)c << 16);
        }
        // static observation of one env (agent_0's view == agent_1's view without agents), repeated tmpl_k times so
        // that the copy period is a whole number of 16-byte chunks
        const int env_bytes = C * OC_OBS_CHANNELS;
        int g = env_bytes, b16 = 16;
        while (b16) { const int t = g % b16; g = b16; b16 = t; }
        L.tmpl_k = 16 / g;
        L.tmpl_chunks = L.tmpl_k * env_bytes / 16;
        for (int k = 0; k < L.tmpl_k; ++k)
            for (int c = 0; c < C; ++c) {
                const unsigned char *q = m + ((OC_PAD + c / w) * WP + OC_PAD + c % w) * 3;
                unsigned char v[32];
                fill_record(q[0] == 8 ? kCodePot + 23 : q[0], v);
                std::memcpy(L.obs_tmpl + (size_t)k * env_bytes + (size_t)c * OC_OBS_CHANNELS, v, OC_OBS_CHANNELS);
            }
    }

    // tile shape: the largest envs-per-warp that still leaves >= 4 CTAs (16 warps) resident per SM, subject to the
    // 16-byte size rule of the bulk copy (epw * C * 26 must be a multiple of 16); OC_B200_EPW overrides (tuning)
    const int smem_cap = 227 * 1024;
    auto pf_bytes_for = [&](int e) { return ((e * (L.span_stride / 16) * 4) + 15) & ~15; };
    const int tmpl_smem = (L.tmpl_chunks * 16 <= 4096) ? L.tmpl_chunks * 16 : 0;
    const int kHdrBytesHost0 = kHdrFixed + ((L.n_scan * 4 + 15) & ~15) + tmpl_smem;
    auto warp_bytes_for = [&](int e) { return ((e * C * OC_OBS_CHANNELS + 2 * e * L.span_stride + 2 * e * 4 + 16) + 15) & ~15; };
    auto size_ok = [&](int e) { return ((e * C * OC_OBS_CHANNELS) % 16) == 0; };
    // (warps per CTA, envs per warp).  Resident warps per SM are limited by shared memory and by the 64 registers x
    // 1024 threads the kernel is compiled for.  Measured on B200: beyond ~24 resident warps more warps stop paying,
    // while larger tiles (full lanes in the step logic, fewer per-tile overheads) and 4-warp CTAs keep paying.  So:
    // among the shapes that reach 24 resident warps take the largest tile, then the largest CTA; if none does,
    // take the shape with the most resident warps.
    int epw = 0, W = kWarpsPerCta, best = 0;
    bool best_enough = false;
    for (int wv = kWarpsPerCta; wv >= 1; wv >>= 1)
        for (int e = 16; e >= 2; e >>= 1) {
            const int bytes = kHdrBytesHost0 + pf_bytes_for(e) + wv * warp_bytes_for(e);
            if (!size_ok(e) || bytes > smem_cap) continue;
            int ctas = smem_cap / (bytes + 1024);
            if (ctas > 32 / wv) ctas = 32 / wv;
            if (ctas < 1) ctas = 1;
            const int warps = ctas * wv;
            const bool enough = warps >= 24;
            bool better;
            if (enough != best_enough) better = enough;
            else if (enough) better = e > epw || (e == epw && wv > W);
            else better = warps > best || (warps == best && (e > epw || (e == epw && wv > W)));
            if (!epw || better) { best = warps; best_enough = enough; epw = e; W = wv; }
        }
    if (const char *ev = std::getenv("OC_B200_EPW")) {
        const int v = std::atoi(ev);
        if ((v == 2 || v == 4 || v == 8 || v == 16) && size_ok(v) &&
            kHdrBytesHost0 + pf_bytes_for(v) + W * warp_bytes_for(v) <= smem_cap) epw = v;
    }
    if (!epw) { delete lay; return OC_ERR_UNSUPPORTED_LAYOUT; }
    lay->epw = epw;
    lay->warp_bytes = warp_bytes_for(epw);
    lay->pf_bytes = pf_bytes_for(epw);
    lay->tmpl_smem_bytes = tmpl_smem;
    lay->warps = W;
    lay->smem_bytes = kHdrBytesHost0 + lay->pf_bytes + W * lay->warp_bytes;
    lay->num_sms = prop.multiProcessorCount;
    lay->occ[0] = lay->occ[1] = lay->occ[2] = 0;
    lay->device = device;
    {   // tile-sized static observation image
        const size_t eb = (size_t)C * OC_OBS_CHANNELS;
        std::vector<unsigned char> img((size_t)epw * eb);
        for (int k = 0; k < epw; ++k) std::memcpy(img.data() + (size_t)k * eb, L.obs_tmpl, eb);
        lay->tmpl_tile = nullptr;
        if (cudaMalloc(&lay->tmpl_tile, img.size()) != cudaSuccess) { delete lay; return fail_cuda(); }
        if (cudaMemcpy(lay->tmpl_tile, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(lay->tmpl_tile); delete lay; return fail_cuda();
        }
    }
    if (cudaMalloc(&lay->dev, sizeof(DevLayout)) != cudaSuccess) { cudaFree(lay->tmpl_tile); delete lay; return fail_cuda(); }
    if (cudaMemcpy(lay->dev, &L, sizeof(DevLayout), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(lay->dev); delete lay; return fail_cuda();
    }
    *out = lay;
    return OC_OK;
}

void oc_layout_destroy(oc_layout *lay) {
    if (!lay) return;
    if (lay->dev) cudaFree(lay->dev);
    if (lay->tmpl_tile) cudaFree(lay->tmpl_tile);
    delete lay;
}

int oc_layout_query(const oc_layout *lay, oc_layout_info *info) {
    if (!lay || !info) return OC_ERR_INVALID_ARG;
    const DevLayout &L = lay->host;
    info->height = L.h; info->width = L.w; info->n_goal = L.n_goal; info->n_pot = L.n_pot;
    info->map_bytes_per_env = L.map_bytes;
    info->obs_bytes_per_agent = (int64_t)L.C * OC_OBS_CHANNELS;
    info->algorithmic_bytes = 58ll * L.C + 117;
    info->envs_per_warp = lay->epw; info->warps_per_cta = lay->warps; info->smem_bytes_per_cta = lay->smem_bytes;
    return OC_OK;
}

int oc_reset(const oc_layout *lay, int64_t N, const uint32_t *keys, oc_state out, uint8_t *obs0, uint8_t *obs1,
             void *stream_) {
    if (!lay || N < 0 || !keys || !state_ok(out) || !obs0 || !obs1) return OC_ERR_INVALID_ARG;
    if (N == 0) return OC_OK;
    if (!state_aligned(out) || !aligned_to(keys, 4)) return OC_ERR_ALIGNMENT;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    long long blocks = (N * lay->host.map_bytes + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    oc_init_static_kernel<<<(unsigned)blocks, 256, 0, stream>>>(lay->dev, N, out);
    if (cudaGetLastError() != cudaSuccess) return fail_cuda();
    KParams kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.N = N; kp.keys = keys; kp.st = out; kp.obs0 = obs0; kp.obs1 = obs1;
    kp.bulk_ok = aligned16(obs0) && aligned16(obs1);
    return launch_env<MODE_RESET>(lay, kp, stream);
}

int oc_get_obs(const oc_layout *lay, int64_t N, oc_state in, uint8_t *obs0, uint8_t *obs1, void *stream_) {
    if (!lay || N < 0 || !state_ok(in) || !obs0 || !obs1) return OC_ERR_INVALID_ARG;
    if (N == 0) return OC_OK;
    if (!state_aligned(in)) return OC_ERR_ALIGNMENT;
    KParams kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.N = N; kp.st = in; kp.obs0 = obs0; kp.obs1 = obs1;
    kp.bulk_ok = aligned16(obs0) && aligned16(obs1);
    return launch_env<MODE_OBS>(lay, kp, static_cast<cudaStream_t>(stream_));
}

int oc_step(const oc_layout *lay, int64_t N, const uint32_t *keys, oc_state in, const int32_t *act0,
            const int32_t *act1, const oc_state *reset_state, oc_state out, uint8_t *obs0, uint8_t *obs1,
            float *reward, float *shaped0, float *shaped1, uint8_t *done, const oc_step_extras *extras,
            void *stream_) {
    if (!lay || N < 0 || !state_ok(in) || !state_ok(out) || !act0 || !act1 || !obs0 || !obs1 || !reward ||
        !shaped0 || !shaped1 || !done)
        return OC_ERR_INVALID_ARG;
    if (!keys && !(extras && extras->split_key && extras->split_num >= 1 && extras->split_first >= 0 &&
                   extras->split_first + N <= extras->split_num && extras->split_num <= 0x7fffffffll))
        return OC_ERR_INVALID_ARG;
    if (reset_state && !state_ok(*reset_state)) return OC_ERR_INVALID_ARG;
    if (N == 0) return OC_OK;
    if (!state_aligned(in) || !state_aligned(out) || (keys && !aligned_to(keys, 4)) || !aligned_to(act0, 4) ||
        !aligned_to(act1, 4) || !aligned_to(reward, 4) || !aligned_to(shaped0, 4) || !aligned_to(shaped1, 4) ||
        (reset_state && !state_aligned(*reset_state)))
        return OC_ERR_ALIGNMENT;
    if (extras) {
        const bool any = extras->episode_returns || extras->episode_lengths || extras->returned_episode_returns ||
                         extras->returned_episode_lengths;
        const bool all = extras->episode_returns && extras->episode_lengths && extras->returned_episode_returns &&
                         extras->returned_episode_lengths;
        if (any && !all) return OC_ERR_INVALID_ARG;
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const DevLayout &L = lay->host;
    // functional call: copy whatever is not aliased, then step `out` in place
#define OC_COPY(field, bytes_per_env)                                                                            \
    if ((const void *)out.field != (const void *)in.field) {                                                     \
        if (cudaMemcpyAsync(out.field, in.field, (size_t)N * (bytes_per_env), cudaMemcpyDeviceToDevice, stream)  \
            != cudaSuccess) return fail_cuda();                                                                  \
    }
    OC_COPY(agent_pos, 16) OC_COPY(agent_dir, 4) OC_COPY(agent_dir_idx, 8) OC_COPY(agent_inv, 8)
    OC_COPY(goal_pos, (size_t)8 * L.n_goal) OC_COPY(pot_pos, (size_t)8 * L.n_pot) OC_COPY(wall_map, (size_t)L.C)
    OC_COPY(maze_map, (size_t)L.map_bytes) OC_COPY(time, 4) OC_COPY(terminal, 1) OC_COPY(task_id, 4)
#undef OC_COPY
    KParams kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.N = N; kp.keys = keys; kp.st = out; kp.act0 = act0; kp.act1 = act1; kp.obs0 = obs0; kp.obs1 = obs1;
    kp.reward = reward; kp.shaped0 = shaped0; kp.shaped1 = shaped1; kp.done = done;
    if (reset_state) { kp.rs = *reset_state; kp.has_rs = 1; }
    if (extras) kp.ex = *extras;
    kp.bulk_ok = aligned16(obs0) && aligned16(obs1);
    return launch_env<MODE_STEP>(lay, kp, stream);
}

int oc_step_host(const oc_layout *lay, int64_t N, const uint32_t *keys, oc_state state, const int32_t *h_actions,
                 void *h_results, void *d_scratch, uint8_t *obs0, uint8_t *obs1, const oc_step_extras *extras,
                 void *stream_) {
    if (!lay || N < 0 || !h_actions || !h_results || !d_scratch) return OC_ERR_INVALID_ARG;
    if (N == 0) return OC_OK;
    if (!aligned_to(d_scratch, 16) || (N & 3)) return OC_ERR_ALIGNMENT;   // keeps every packed sub-array 4-byte aligned
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    unsigned char *d = static_cast<unsigned char *>(d_scratch);
    int32_t *d_act = reinterpret_cast<int32_t *>(d);                 // [2,N]
    float *d_res = reinterpret_cast<float *>(d + 8 * N);             // reward | shaped0 | shaped1 (f32[N] each) | done u8[N]
    if (cudaMemcpyAsync(d_act, h_actions, (size_t)8 * N, cudaMemcpyHostToDevice, stream) != cudaSuccess) return fail_cuda();
    const int rc = oc_step(lay, N, keys, state, d_act, d_act + N, nullptr, state, obs0, obs1, d_res, d_res + N,
                           d_res + 2 * N, reinterpret_cast<uint8_t *>(d_res + 3 * N), extras, stream_);
    if (rc != OC_OK) return rc;
    if (cudaMemcpyAsync(h_results, d_res, (size_t)13 * N, cudaMemcpyDeviceToHost, stream) != cudaSuccess) return fail_cuda();
    return OC_OK;
}

struct oc_host_stepper {
    const oc_layout *lay;
    int64_t N;
    oc_state state;
    void *h_results, *d_scratch;
    uint8_t *obs0, *obs1;
    oc_step_extras extras;
    void *stream;
};

int oc_host_stepper_create(const oc_layout *lay, int64_t N, oc_state state, void *h_results, void *d_scratch,
                           uint8_t *obs0, uint8_t *obs1, const oc_step_extras *extras, void *stream,
                           oc_host_stepper **out) {
    if (!lay || !out || N <= 0 || !state_ok(state) || !h_results || !d_scratch || !obs0 || !obs1 || !extras)
        return OC_ERR_INVALID_ARG;
    oc_host_stepper *h = new (std::nothrow) oc_host_stepper();
    if (!h) return OC_ERR_INVALID_ARG;
    h->lay = lay; h->N = N; h->state = state; h->h_results = h_results; h->d_scratch = d_scratch;
    h->obs0 = obs0; h->obs1 = obs1; h->extras = *extras; h->stream = stream;
    *out = h;
    return OC_OK;
}

int oc_host_stepper_step(oc_host_stepper *h, const uint32_t *split_key, const int32_t *h_actions) {
    if (!h || !split_key || !h_actions) return OC_ERR_INVALID_ARG;
    h->extras.split_key = split_key;
    return oc_step_host(h->lay, h->N, nullptr, h->state, h_actions, h->h_results, h->d_scratch, h->obs0, h->obs1,
                        &h->extras, h->stream);
}

int oc_host_stepper_wait(oc_host_stepper *h) {
    if (!h) return OC_ERR_INVALID_ARG;
    return cudaStreamSynchronize(static_cast<cudaStream_t>(h->stream)) == cudaSuccess ? OC_OK : fail_cuda();
}

void oc_host_stepper_destroy(oc_host_stepper *h) { delete h; }

int oc_prng_split(int prng_mode, const uint32_t *key, int64_t num, int64_t first, int64_t count, uint32_t *out,
                  void *stream_) {
    if (!key || !out || num < 1 || first < 0 || count < 0 || first + count > num || num > 0x7fffffffll)
        return OC_ERR_INVALID_ARG;
    if (prng_mode != OC_PRNG_THREEFRY_LEGACY && prng_mode != OC_PRNG_THREEFRY_PARTITIONABLE) return OC_ERR_INVALID_ARG;
    if (count == 0) return OC_OK;
    long long blocks = (count + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    oc_split_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream_)>>>(prng_mode, key, num, first, count, out);
    return cudaGetLastError() == cudaSuccess ? OC_OK : fail_cuda();
}

int oc_prng_chain(int prng_mode, uint32_t *key, int n, uint32_t *subkeys, void *stream_) {
    if (!key || !subkeys || n < 0) return OC_ERR_INVALID_ARG;
    if (prng_mode != OC_PRNG_THREEFRY_LEGACY && prng_mode != OC_PRNG_THREEFRY_PARTITIONABLE) return OC_ERR_INVALID_ARG;
    if (n == 0) return OC_OK;
    oc_chain_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream_)>>>(prng_mode, key, n, subkeys);
    return cudaGetLastError() == cudaSuccess ? OC_OK : fail_cuda();
}

}  // extern "C"
