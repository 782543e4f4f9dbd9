// oc_common.cuh: constants, the device-side layout descriptor and the kernel parameter block -- part of the sm_100a Overcooked kernels (see oc_b200.cu for the overview)
#pragma once
#include "oc_b200.h"

#include <cuda_runtime.h>

#include <cstdint>

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
    uint8_t *cells;       // compact observation (scan-cell codes, agent cells, urgency), or NULL
    int act_off;    // MODE_ROLLOUT: the warp's action ring inside its shared memory
    int cell_stride, cell_agents, cells_off;   // cells_off: staging area inside the warp's shared memory (CELLS launches)
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
    // MODE_ROLLOUT: T steps per launch.  Per-step arrays are [T, N] (keys [T, N, 2] or ex.split_key [T, 2]); the observations
    // of step t start obs_step_stride bytes after those of step t-1
    int T;
    long long obs_step_stride;
    const uint8_t *act8;   // or NULL: packed actions u8[T, N], agent_0 | agent_1 << 4 (replaces act0 / act1)
    uint8_t *res8;         // or NULL: packed results u8[T, N] (oc_rollout_io.results8)
    int interleaved;       // obs0 is uint8[N, 2, h, w, 26] (both views of an env side by side): launches the VAR_ALL instantiation
};

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_OBS = 2, MODE_ROLLOUT = 3, NUM_MODES = 4 };
enum { VAR_PLAIN = 0, VAR_CELLS = 1, VAR_ALL = 2, NUM_VARS = 3 };   // oc_env_kernel<MODE, VAR>

}  // namespace

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
    int cell_stride, cell_agents;   // compact observation: bytes per env, offset of the agent entries (oc_step_extras.cells)
    int hdr_bytes;             // CTA-shared header of the kernels' shared memory (bitmaps, scan list, static image)
    mutable int occ_ro[NUM_MODES][NUM_VARS][5];   // tile shapes are picked per launch: resident CTAs per SM by mode, variant, log2(envs per warp)   // resident CTAs per SM of each kernel mode (filled on first launch)
};

