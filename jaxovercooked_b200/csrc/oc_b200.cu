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
#include "oc_kernels.cuh"

// ================================================================================================ host side


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

// shared memory of a tile shape: the per-CTA copy plan and one warp's area (staging image, span buffer(s), urgency flags, mbarrier)
int pf_bytes_of(const DevLayout &L, int e) { return ((e * (L.span_stride / 16) * 4) + 15) & ~15; }
int warp_bytes_of(const DevLayout &L, int e, int span_bufs) {
    return ((e * L.C * OC_OBS_CHANNELS + span_bufs * e * L.span_stride + 2 * e * 4 + 16) + 15) & ~15;
}
bool tile_size_ok(const DevLayout &L, int e) { return ((e * L.C * OC_OBS_CHANNELS) % 16) == 0; }   // bulk-copy size rule
constexpr int kActRingBytes = 4 * 32 * 4;   // oc_kernels.cuh: ACT_RING slots of 32 lanes

// MODE_ROLLOUT keeps a tile for T steps, so its best tile shape depends on the batch.  Measured on B200 (cramped_room, us per
// step by envs per warp; profiles/r02_rollout_tile_shapes.txt):
//     envs      16    1024   4096   8192  16384  32768  65536
//     2        2.63   2.87   3.44   4.77   9.09   17.7   34.7      few environments: a step is one warp's latency chain, so
//     4        2.82   3.09   4.77   3.88   5.26   9.91   19.2      small tiles on many warps and SMs win;
//     8        3.41   3.74   3.84   5.94   6.01   7.17   15.0      once every SM has a full set of resident warps, 16-env tiles
//     16       4.19   4.73   4.90   6.24   6.26   6.90   14.4      win (every lane busy in the step logic, half the per-tile work)
// OC_B200_RO_EPW overrides (tuning).
int rollout_epw(const oc_layout *lay, long long N, int cells_smem, bool rollout) {
    const DevLayout &H = lay->host;
    const int span_bufs = rollout ? 1 : 2;
    auto fits = [&](int e) {
        return tile_size_ok(H, e) && lay->hdr_bytes + pf_bytes_of(H, e) +
               lay->warps * (warp_bytes_of(H, e, span_bufs) + (rollout ? kActRingBytes : 0) + cells_smem * e) <= 227 * 1024;
    };
    if (const char *ev = std::getenv(rollout ? "OC_B200_RO_EPW" : "OC_B200_STEP_EPW")) {
        const int v = std::atoi(ev);
        if ((v == 2 || v == 4 || v == 8 || v == 16) && fits(v)) return v;
    }
    const long long per_sm = (N + lay->num_sms - 1) / lay->num_sms;     // environments per SM
    if (rollout && per_sm > 112) {
        // large batch: the largest tile that still leaves enough resident warps per SM (shared memory decides).  Measured at
        // 65,536 envs (us per step, 16 / 8 / 4 envs per warp; gpurun_out/r03f.txt): 6x6 28.2 / 24.6 / 24.9, 5x9 29.8 / 32.9 / 30.0,
        // 7x9 56.3 / 43.7 / -, 8x10 76.9 / 50.1 / 55.8, 10x10 70.9 / 63.7 / 67.5, 9x11 83.0 / 65.3 / -: 8 resident warps are
        // enough when the static image arrives through the copy engine, 12 are wanted when the lanes copy it
        auto res_warps = [&](int e) {
            const int cta = lay->hdr_bytes + pf_bytes_of(H, e) + lay->warps * (warp_bytes_of(H, e, 1) + kActRingBytes + cells_smem * e) + 1024;
            int ctas = (227 * 1024) / cta;
            if (ctas > 7) ctas = 7;                 // 72 registers x 128 threads
            return ctas * lay->warps;
        };
        const int need = lay->tmpl_smem_bytes ? 12 : 8;
        for (int e = 16; e >= 8; e >>= 1)
            if (fits(e) && res_warps(e) >= need) return e;
        for (int e = 16; e >= 8; e >>= 1)
            if (fits(e) && res_warps(e) >= 8) return e;
        int best_e = 0, best_w = 0;
        for (int e = 16; e >= 2; e >>= 1)
            if (fits(e) && res_warps(e) > best_w) { best_w = res_warps(e); best_e = e; }
        if (best_e) return best_e;
    }
    // one launch per step (oc_step / oc_get_obs / oc_reset): the layout's shape (8 envs per warp for cramped_room: 16 measured 8 %
    // slower there) unless the batch is small, where the same latency argument as for rollouts holds
    const int want = per_sm <= 28 ? 2 : per_sm <= 112 ? 4 : (rollout ? 16 : lay->epw);
    if (!rollout && want >= lay->epw) return lay->epw;
    for (int e = want; e <= 16; e <<= 1)
        if (fits(e)) return e;
    for (int e = want >> 1; e >= 2; e >>= 1)
        if (fits(e)) return e;
    return lay->epw;
}

template <int MODE>
int launch_env(const oc_layout *lay, KParams &kp, cudaStream_t stream) {
    const DevLayout &H = lay->host;
    kp.L = lay->dev;
    kp.tmpl_tile = lay->tmpl_tile;
    const int epw = rollout_epw(lay, kp.N, kp.cells ? lay->cell_stride : 0, MODE == MODE_ROLLOUT);
    int warps = lay->warps;      // warps per CTA (OC_B200_RO_WARPS: tuning knob for the rollout mode)
    if (MODE == MODE_ROLLOUT) {
        static const int w_env = [] { const char *e = std::getenv("OC_B200_RO_WARPS"); return e ? std::atoi(e) : 0; }();
        if (w_env >= 1 && w_env <= lay->warps) warps = w_env;
    }
    kp.epw = epw;
    kp.h = H.h; kp.w = H.w; kp.C = H.C; kp.WP = H.WP; kp.map_bytes = H.map_bytes; kp.span_off = H.span_off;
    kp.span_len = H.span_len; kp.span_stride = H.span_stride; kp.n_pot = H.n_pot;
    kp.max_steps = H.max_steps; kp.random_reset = H.random_reset; kp.part = H.prng_mode;
    kp.agent_cell0 = H.agent_cell[0]; kp.agent_cell1 = H.agent_cell[1];
    kp.magic_C = H.magic_C; kp.magic_w = H.magic_w;
    kp.n_scan = H.n_scan; kp.tmpl_k = H.tmpl_k; kp.tmpl_chunks = H.tmpl_chunks; kp.tmpl_bytes = lay->tmpl_smem_bytes;
    kp.scan_bytes = (H.n_scan * 4 + 15) & ~15;
    kp.nch = H.span_stride / 16; kp.magic_nch = magic_of((unsigned)kp.nch);
    kp.pf_bytes = pf_bytes_of(H, epw);
    kp.pf_invariant = (((long long)epw * H.map_bytes) % 16 == 0 && (long long)epw * H.map_bytes + 32 < 65536) ? 1 : 0;
    kp.magic_ns = magic_of((unsigned)(H.n_scan > 0 ? H.n_scan : 1)); kp.magic_tc = magic_of((unsigned)H.tmpl_chunks);
    kp.cell_stride = lay->cell_stride; kp.cell_agents = lay->cell_agents;
    kp.n_tiles = (kp.N + epw - 1) / epw;
    if (kp.n_tiles < 1) return OC_OK;
    const bool with_cells = kp.cells != nullptr;
    if (with_cells && !aligned16(kp.cells)) return OC_ERR_ALIGNMENT;
    if (kp.interleaved && (with_cells || epw < 2 || MODE == MODE_RESET)) return OC_ERR_INVALID_ARG;
    const int var = kp.interleaved ? VAR_ALL : (with_cells ? VAR_CELLS : VAR_PLAIN);
    void (*kern)(const KParams) = kp.interleaved ? oc_env_kernel<MODE, (MODE == MODE_RESET) ? VAR_PLAIN : VAR_ALL>
                                  : (with_cells ? oc_env_kernel<MODE, VAR_CELLS> : oc_env_kernel<MODE, VAR_PLAIN>);
    // the compact observations of a tile are staged per warp, behind the warp's regular shared-memory area
    const int cells_smem = with_cells ? epw * lay->cell_stride : 0;
    // MODE_ROLLOUT keeps one span buffer per warp instead of two (oc_kernels.cuh: SPAN_BUFS)
    const int warp_bytes0 = warp_bytes_of(H, epw, MODE == MODE_ROLLOUT ? 1 : 2);
    const int act_smem = (MODE == MODE_ROLLOUT) ? kActRingBytes : 0;
    const int smem_bytes = lay->hdr_bytes + kp.pf_bytes + warps * (warp_bytes0 + act_smem + cells_smem);
    if (smem_bytes > 227 * 1024) return OC_ERR_UNSUPPORTED_LAYOUT;
    kp.act_off = warp_bytes0;
    kp.cells_off = warp_bytes0 + act_smem;
    kp.warp_bytes = warp_bytes0 + act_smem + cells_smem;
    const int le = epw == 16 ? 4 : epw == 8 ? 3 : epw == 4 ? 2 : epw == 2 ? 1 : 0;
    int &occ = lay->occ_ro[MODE][var][le];
    if (occ == 0) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return fail_cuda();
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, warps * 32, smem_bytes) != cudaSuccess)
            return fail_cuda();
        occ = nb < 1 ? 1 : nb;
    }
    // persistent grid: every resident warp gets the same number of tiles (k), so there is no ragged last wave
    const long long resident = (long long)lay->num_sms * occ;
    const long long want = (kp.n_tiles + warps - 1) / warps;
    long long ctas = want;
    if (want > resident) {
        const long long k = (want + resident - 1) / resident;
        ctas = (want + k - 1) / k;
    }
    static const bool use_pdl = [] { const char *e = std::getenv("OC_B200_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3((unsigned)warps * 32);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, kp) == cudaSuccess ? OC_OK : fail_cuda();
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
    // static image kept in shared memory (copied by the lanes) up to this size; larger ones are pulled per tile by the copy engine
    static const int tmpl_smem_max = [] { const char *e = std::getenv("OC_B200_TMPL_SMEM_MAX"); return e ? std::atoi(e) : 4096; }();
    const int tmpl_smem = (L.tmpl_chunks * 16 <= tmpl_smem_max) ? L.tmpl_chunks * 16 : 0;
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
    lay->hdr_bytes = kHdrBytesHost0;
    std::memset(lay->occ_ro, 0, sizeof(lay->occ_ro));
    lay->num_sms = prop.multiProcessorCount;
    lay->cell_agents = (L.n_scan + 1) & ~1;
    lay->cell_stride = (lay->cell_agents + 5 + 15) & ~15;
    lay->device = device;
    {   // tile-sized static observation image
        const size_t eb = (size_t)C * OC_OBS_CHANNELS;
        const int copies = 16;      // the largest tile any launch uses (MODE_ROLLOUT picks its shape per launch)
        std::vector<unsigned char> img((size_t)copies * eb);
        for (int k = 0; k < copies; ++k) std::memcpy(img.data() + (size_t)k * eb, L.obs_tmpl, eb);
        lay->tmpl_tile = nullptr;
        if (cudaMalloc(&lay->tmpl_tile, img.size()) != cudaSuccess) { delete lay; return fail_cuda(); }
        if (cudaMemcpy(lay->tmpl_tile, img.data(), img.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(lay->tmpl_tile); delete lay; return fail_cuda();
        }
    }
    if (cudaMalloc(&lay->dev, sizeof(DevLayout)) != cudaSuccess) { cudaFree(lay->tmpl_tile); delete lay; return fail_cuda(); }
    if (cudaMemcpy(lay->dev, &L, sizeof(DevLayout), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(lay->dev); cudaFree(lay->tmpl_tile); delete lay; return fail_cuda();
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
    info->n_scan = L.n_scan; info->cell_stride = lay->cell_stride; info->cell_agents = lay->cell_agents;
    return OC_OK;
}

int oc_layout_cell_tables(const oc_layout *lay, uint8_t *records, int32_t *scan_cells) {
    if (!lay) return OC_ERR_INVALID_ARG;
    if (records)
        for (int c = 0; c < OC_NUM_CELL_CODES; ++c) std::memcpy(records + c * OC_OBS_CHANNELS, lay->host.rec[c], OC_OBS_CHANNELS);
    if (scan_cells)
        for (int i = 0; i < lay->host.n_scan; ++i) scan_cells[i] = (int32_t)(lay->host.scan[i] >> 16);
    return OC_OK;
}

int oc_layout_obs_template(const oc_layout *lay, uint8_t *out) {
    if (!lay || !out) return OC_ERR_INVALID_ARG;
    std::memcpy(out, lay->host.obs_tmpl, (size_t)lay->host.C * OC_OBS_CHANNELS);
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
    return oc_get_obs_cells(lay, N, in, obs0, obs1, nullptr, stream_);
}

int oc_get_obs_all(const oc_layout *lay, int64_t N, oc_state in, uint8_t *obs_all, void *stream_) {
    if (!lay || N < 0 || !state_ok(in) || !obs_all) return OC_ERR_INVALID_ARG;
    if (N == 0) return OC_OK;
    if (!state_aligned(in)) return OC_ERR_ALIGNMENT;
    KParams kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.N = N; kp.st = in; kp.obs0 = obs_all; kp.obs1 = obs_all + (size_t)lay->host.C * OC_OBS_CHANNELS; kp.interleaved = 1;
    kp.bulk_ok = aligned16(obs_all);
    return launch_env<MODE_OBS>(lay, kp, static_cast<cudaStream_t>(stream_));
}

int oc_get_obs_cells(const oc_layout *lay, int64_t N, oc_state in, uint8_t *obs0, uint8_t *obs1, uint8_t *cells,
                     void *stream_) {
    if (!lay || N < 0 || !state_ok(in) || !obs0 || !obs1) return OC_ERR_INVALID_ARG;
    if (N == 0) return OC_OK;
    if (!state_aligned(in)) return OC_ERR_ALIGNMENT;
    KParams kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.N = N; kp.st = in; kp.obs0 = obs0; kp.obs1 = obs1; kp.cells = cells;
    kp.bulk_ok = aligned16(obs0) && aligned16(obs1);
    return launch_env<MODE_OBS>(lay, kp, static_cast<cudaStream_t>(stream_));
}

int oc_step(const oc_layout *lay, int64_t N, const uint32_t *keys, oc_state in, const int32_t *act0,
            const int32_t *act1, const oc_state *reset_state, oc_state out, uint8_t *obs0, uint8_t *obs1,
            float *reward, float *shaped0, float *shaped1, uint8_t *done, const oc_step_extras *extras,
            void *stream_) {
    const bool inter = extras && extras->obs_interleaved;
    if (!lay || N < 0 || !state_ok(in) || !state_ok(out) || !act0 || !act1 || !obs0 || (!obs1 && !inter) || !reward ||
        !shaped0 || !shaped1 || !done)
        return OC_ERR_INVALID_ARG;
    if (inter && ((extras->obs_interleaved != OC_OBS_INTERLEAVED) || extras->cells ||
                  (obs1 && obs1 != obs0 + (size_t)lay->host.C * OC_OBS_CHANNELS)))
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
    if (extras) { kp.ex = *extras; kp.cells = extras->cells; }
    kp.interleaved = inter ? 1 : 0;
    kp.bulk_ok = aligned16(obs0) && (inter || aligned16(obs1));
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

int oc_rollout(const oc_layout *lay, int64_t N, int32_t T, const uint32_t *keys, oc_state st, const oc_rollout_io *io,
               const oc_step_extras *extras, void *stream_) {
    if (!lay || N < 0 || T < 1 || T > 65535 || !io || !state_ok(st) || !io->obs0 || (!io->obs1 && !io->obs_interleaved))
        return OC_ERR_INVALID_ARG;
    if (io->obs_interleaved && (io->obs_interleaved != OC_OBS_INTERLEAVED || (extras && extras->cells) ||
                                (io->obs1 && io->obs1 != io->obs0 + (size_t)lay->host.C * OC_OBS_CHANNELS)))
        return OC_ERR_INVALID_ARG;
    if (!io->actions8 && !(io->act0 && io->act1)) return OC_ERR_INVALID_ARG;
    if (!io->results8 && !(io->reward && io->shaped0 && io->shaped1 && io->done)) return OC_ERR_INVALID_ARG;
    if (!keys && !(extras && extras->split_key && extras->split_num >= 1 && extras->split_first >= 0 &&
                   extras->split_first + N <= extras->split_num && extras->split_num <= 0x7fffffffll))
        return OC_ERR_INVALID_ARG;
    const long long view_bytes = (long long)N * lay->host.C * OC_OBS_CHANNELS * (io->obs_interleaved ? 2 : 1);
    if (T > 1 && io->obs_step_stride < view_bytes) return OC_ERR_INVALID_ARG;
    if (N == 0) return OC_OK;
    if (!state_aligned(st) || (keys && !aligned_to(keys, 4)) || (io->act0 && !aligned_to(io->act0, 4)) ||
        (io->act1 && !aligned_to(io->act1, 4)) || (io->reward && !aligned_to(io->reward, 4)) ||
        (io->shaped0 && !aligned_to(io->shaped0, 4)) || (io->shaped1 && !aligned_to(io->shaped1, 4)))
        return OC_ERR_ALIGNMENT;
    if (extras) {
        const bool any = extras->episode_returns || extras->episode_lengths || extras->returned_episode_returns ||
                         extras->returned_episode_lengths;
        const bool all = extras->episode_returns && extras->episode_lengths && extras->returned_episode_returns &&
                         extras->returned_episode_lengths;
        if (any && !all) return OC_ERR_INVALID_ARG;
    }
    KParams kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.N = N; kp.T = T; kp.keys = keys; kp.st = st;
    kp.act0 = io->act0; kp.act1 = io->act1; kp.act8 = io->actions8;
    // large batch of a small layout, int32 actions: pack them first (see oc_rollout_io; measured 13.7 vs 14.4 us per step at 65,536
    // cramped_room envs, nothing to gain at 16,384 envs or when an env's observations are several KB)
    if (!io->actions8 && io->actions8_scratch && N >= 32768 && lay->host.C <= 64) {
        static const bool use_pack = [] { const char *e = std::getenv("OC_B200_PACK_ACTIONS"); return !(e && e[0] == '0'); }();
        if (use_pack) {
            const long long n = (long long)T * N;
            long long blocks = (n / 4 + 255) / 256;
            if (blocks > 148 * 8) blocks = 148 * 8;
            oc_pack_actions_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream_)>>>(io->act0, io->act1, n, io->actions8_scratch);
            if (cudaGetLastError() != cudaSuccess) return fail_cuda();
            kp.act8 = io->actions8_scratch;
        }
    }
    kp.obs0 = io->obs0; kp.obs1 = io->obs1; kp.obs_step_stride = io->obs_step_stride;
    kp.reward = io->reward; kp.shaped0 = io->shaped0; kp.shaped1 = io->shaped1; kp.done = io->done; kp.res8 = io->results8;
    if (extras) { kp.ex = *extras; kp.cells = extras->cells; }
    kp.interleaved = io->obs_interleaved ? 1 : 0;
    kp.bulk_ok = aligned16(io->obs0) && (io->obs_interleaved || aligned16(io->obs1)) && (io->obs_step_stride % 16 == 0);
    return launch_env<MODE_ROLLOUT>(lay, kp, static_cast<cudaStream_t>(stream_));
}

int oc_rollout_host(const oc_layout *lay, int64_t N, int32_t T, oc_state state, int wire, const void *h_actions,
                    void *h_results, void *d_scratch, uint8_t *obs0, uint8_t *obs1, int64_t obs_step_stride,
                    const oc_step_extras *extras, void *stream_) {
    if (!lay || N < 0 || T < 1 || !h_actions || !h_results || !d_scratch || !extras || !extras->split_key)
        return OC_ERR_INVALID_ARG;
    if (wire != OC_WIRE_REFERENCE && wire != OC_WIRE_PACKED) return OC_ERR_INVALID_ARG;
    if (N == 0) return OC_OK;
    const long long TN = (long long)T * N;
    if (!aligned_to(d_scratch, 16) || (TN & 3)) return OC_ERR_ALIGNMENT;   // keeps every packed sub-array 4-byte aligned
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    unsigned char *d = static_cast<unsigned char *>(d_scratch);
    oc_rollout_io io;
    std::memset(&io, 0, sizeof(io));
    io.obs0 = obs0; io.obs1 = obs1; io.obs_step_stride = obs_step_stride;
    const size_t in_bytes = (size_t)(wire == OC_WIRE_PACKED ? 1 : 8) * TN, out_bytes = (size_t)(wire == OC_WIRE_PACKED ? 1 : 13) * TN;
    unsigned char *d_out = d + in_bytes;
    if (wire == OC_WIRE_PACKED) {
        io.actions8 = d; io.results8 = d_out;
    } else {
        io.act0 = reinterpret_cast<const int32_t *>(d); io.act1 = io.act0 + TN;
        io.reward = reinterpret_cast<float *>(d_out); io.shaped0 = io.reward + TN; io.shaped1 = io.reward + 2 * TN;
        io.done = reinterpret_cast<uint8_t *>(io.reward + 3 * TN);
    }
    if (cudaMemcpyAsync(d, h_actions, in_bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess) return fail_cuda();
    const int rc = oc_rollout(lay, N, T, nullptr, state, &io, extras, stream_);
    if (rc != OC_OK) return rc;
    if (cudaMemcpyAsync(h_results, d_out, out_bytes, cudaMemcpyDeviceToHost, stream) != cudaSuccess) return fail_cuda();
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
