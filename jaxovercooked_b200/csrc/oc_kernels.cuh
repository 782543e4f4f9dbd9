// oc_kernels.cuh: the fused step / reset / get_obs kernel and the small PRNG / initialisation kernels -- part of the sm_100a Overcooked kernels (see oc_b200.cu for the overview)
#pragma once
#include "oc_common.cuh"
#include "oc_rules.cuh"
#include "threefry.cuh"

#ifndef OC_OBS_EVICT_FIRST
#define OC_OBS_EVICT_FIRST 0
#endif
#ifndef OC_ROLLOUT_MIN_CTAS
#define OC_ROLLOUT_MIN_CTAS 7   // resident CTAs per SM the rollout instantiation is compiled for (72 registers)
#endif

namespace {

// one bulk asynchronous copy shared -> global (the copy engine; SASS UBLKCP).  OC_OBS_EVICT_FIRST (tuning build): the
// observation stream is marked evict-first in L2 so that it does not push the (re-read) action arrays out
__device__ __forceinline__ void bulk_store(unsigned char *dst, uint32_t src_smem, int bytes) {
#if OC_OBS_EVICT_FIRST
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(dst), "r"(src_smem), "r"(bytes), "l"(pol) : "memory");
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src_smem), "r"(bytes) : "memory");
#endif
}

// ------------------------------------------------------------------------------------------------ kernel

// CELLS: also emit the compact form of the observations (oc_step_extras.cells); a separate instantiation so that the
// plain kernel keeps its register budget
template <int MODE, int VAR>
__global__ void __launch_bounds__(kWarpsPerCta * 32, (MODE == MODE_ROLLOUT || VAR == VAR_ALL) ? OC_ROLLOUT_MIN_CTAS : 8) oc_env_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    // VAR_CELLS: also emit the compact observations.  VAR_ALL: the two views of an env are written side by side
    // (uint8[N, 2, h, w, 26], the CTRolloutManager's obs["__all__"], wrappers/baselines.py:196) instead of as two [N, h, w, 26] arrays
    constexpr bool CELLS = (VAR == VAR_CELLS), INTERLEAVED = (VAR == VAR_ALL);
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
    // the interior spans of the tile: double-buffered (the next tile is prefetched while this one is processed), except in
    // MODE_ROLLOUT, where a tile stays for T steps and one buffer is enough -- the shared memory goes to occupancy instead
    constexpr int SPAN_BUFS = (MODE == MODE_ROLLOUT) ? 1 : 2;
    unsigned char *smap_buf = wbase + view_bytes;                              // SPAN_BUFS x [epw][span_stride]
    uint32_t *s_ap = reinterpret_cast<uint32_t *>(smap_buf + SPAN_BUFS * epw * span_stride);   // [2*epw] urgency flags
    const uint32_t tbar = smem_u32(s_ap + 2 * epw);                                    // this warp's mbarrier (8 bytes)
    // MODE_ROLLOUT: ring of the next steps' actions, one 32-bit slot per lane and step, filled by 4-byte cp.async three steps
    // ahead (a register prefetch one step ahead measured 7 % slower at 65,536 envs: the int32 action loads miss L2)
    constexpr int ACT_RING = 4;
    uint32_t *s_act = reinterpret_cast<uint32_t *>(wbase + p.act_off);   // [ACT_RING][32]
    unsigned char *s_cells = wbase + p.cells_off;        // CELLS: [epw][cell_stride] staging of the compact observations

    const int e = lane >> 1, ai = lane & 1;   // env within the tile, agent index (0 = "Alice", 1 = "Bob")
    // MODE_ROLLOUT is MODE_STEP applied T times to the same tile: the state is loaded once, lives in shared memory and
    // registers for the T steps, and is written back once (multi_agent_env.py:41-69 under the callers' lax.scan)
    constexpr bool STEPPING = (MODE == MODE_STEP || MODE == MODE_ROLLOUT);
    const int T = (MODE == MODE_ROLLOUT) ? p.T : 1;
    auto load_act = [&](int ts, long long en) -> int {
        if (MODE == MODE_ROLLOUT) {
            const long long j = (long long)ts * p.N + en;
            if (p.act8) { const int v = __ldg(p.act8 + j); return ai ? (v >> 4) : (v & 15); }
            return __ldg((ai ? p.act1 : p.act0) + j);
        }
        return __ldg((ai ? p.act1 : p.act0) + en);
    };
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
            if (STEPPING) {
                f.dir = reinterpret_cast<const unsigned short *>(p.st.agent_dir)[en * 2 + ai];
                f.act = load_act(0, en);
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
    if (CELLS)         // padding bytes of the compact form stay 0xFF
        for (int i = lane; i < (epw * p.cell_stride) >> 2; i += 32) reinterpret_cast<uint32_t *>(s_cells)[i] = 0xffffffffu;
    __syncthreads();   // the tables (incl. the prefetch plan) are complete before any warp uses them
    asm volatile("griddepcontrol.wait;" ::: "memory");
    long long tile = (long long)blockIdx.x * cta_warps + warp;
    int buf = 0;
    Fields nf;
    if (tile < p.n_tiles) { prefetch_spans(tile, 0); nf = load_fields(tile); }

    for (; tile < p.n_tiles; tile += tile_stride, buf ^= (SPAN_BUFS - 1)) {
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
        auto refill_stage = [&]() {
            if (!p.tmpl_bytes) {
                if (lane == 0) {
                    if (bulk_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // obs1 of the previous tile / step
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(tbar), "r"(view_bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(smem_u32(stage)), "l"(p.tmpl_tile), "r"(view_bytes), "r"(tbar) : "memory");
                }
                bulk_pending = false;
            }
        };
        refill_stage();
        // ------------------------------------------------ (1)+(2) this tile's state; prefetch the next tile's
        const Fields cf = nf;
        if (SPAN_BUFS == 1) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else if (tile + tile_stride < p.n_tiles) {
            prefetch_spans(tile + tile_stride, buf ^ 1);
            nf = load_fields(tile + tile_stride);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        int px = (int)cf.pos.x, py = (int)cf.pos.y, di = cf.di, inv = cf.inv, t = cf.t;
        int dx = (int)(signed char)(cf.dir & 0xffu), dy = (int)(signed char)(cf.dir >> 8);
        int act = clampi(cf.act, 0, 5);
        float lw_er = 0.f, lw_el = 0.f, lw_rr = 0.f, lw_rl = 0.f;   // LogWrapper trackers of this agent
        int acc_done = 0, acc_rew = 0, acc_s0 = 0, acc_s1 = 0, acc_rr = 0, acc_rl = 0;   // MODE_ROLLOUT: statistics of the T steps
        __syncwarp();

        const bool act_ring = (MODE == MODE_ROLLOUT) && !p.act8;
        auto issue_act = [&](int ts2) {   // this lane's action of step ts2 -> its ring slot, asynchronously; one group per step
            if (ts2 < T && active)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
                             :: "r"(smem_u32(s_act + (ts2 & (ACT_RING - 1)) * 32 + lane)),
                                "l"((ai ? p.act1 : p.act0) + (long long)ts2 * p.N + env) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (act_ring) { issue_act(1); issue_act(2); }

      for (int ts = 0; ts < T; ++ts) {
        const bool last_ts = (MODE != MODE_ROLLOUT) || ts == T - 1;
        if (act_ring) issue_act(ts + 3);
        if (MODE == MODE_ROLLOUT && ts > 0) refill_stage();   // every step emits from a fresh static image
        int nact = 4;
        if (MODE == MODE_ROLLOUT && !act_ring && !last_ts && active) nact = load_act(ts + 1, env);   // packed actions: register prefetch
        int nx = px, ny = py, ndi = di, ninv = inv, t_out = t, term_out = 0;
        bool done = (MODE == MODE_RESET);
        int rew = 0, shaped = 0;
        // a cell is updated in the staged span AND, in place, in HBM (only cells that change are ever written back)
        // (MODE_ROLLOUT: the staged span alone; every span is stored once, after the last step)
        auto put_cell = [&](int off, int a, int b, int c) {
            my_span[off] = (unsigned char)a; my_span[off + 1] = (unsigned char)b; my_span[off + 2] = (unsigned char)c;
            if (MODE != MODE_ROLLOUT) {
                g_span[off] = (unsigned char)a; g_span[off + 1] = (unsigned char)b; g_span[off + 2] = (unsigned char)c;
            }
        };

        if (STEPPING) {
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
                    if (s >= 1 && s <= 20) {
                        my_span[off] = (unsigned char)(s - 1);
                        if (MODE != MODE_ROLLOUT) g_span[off] = (unsigned char)(s - 1);
                    }
                }
            }
            rew += __shfl_xor_sync(0xffffffffu, rew, 1);                   // :502 reward = alice + bob
            t_out = t + 1;                                                  // :118
            done = t_out >= p.max_steps;                                    // :120, :644-647
            __syncwarp();
        }

        // -------------------------------------------------------------------- (3) reset (overcooked.py:136-248)
        bool full_writeback = (MODE == MODE_RESET) || (MODE == MODE_ROLLOUT);
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
                        // MODE_ROLLOUT: step ts has its own keys -- row ts of keys[T, N, 2] or of split_key[T, 2]
                        const long long kj = (MODE == MODE_ROLLOUT) ? (long long)ts * p.N + env : env;
                        if (p.keys) { k0 = __ldg(p.keys + kj * 2); k1 = __ldg(p.keys + kj * 2 + 1); }
                        else {   // the caller's split(key, num_envs) fused in: derive this env's row on demand
                            const uint32_t *sk = p.ex.split_key + ((MODE == MODE_ROLLOUT) ? 2 * ts : 0);
                            split_row(part, __ldg(sk), __ldg(sk + 1), p.ex.split_num, p.ex.split_first + env, k0, k1);
                        }
                        uint32_t s0, s1;
                        if (STEPPING) { split2(part, k0, k1, s0, s1); k0 = s0; k1 = s1; }   // key_reset
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
        if ((MODE == MODE_STEP || (MODE == MODE_ROLLOUT && ts == 0)) && p.ex.episode_returns && active) {
            const long long j = env * 2 + ai;
            lw_er = p.ex.episode_returns[j]; lw_el = p.ex.episode_lengths[j];
            lw_rr = p.ex.returned_episode_returns[j]; lw_rl = p.ex.returned_episode_lengths[j];
        }

        // ---------------------------------------------------------------- (4) observations (overcooked.py:250-364)
        bool obs1_pending = false;
        int code_v1 = 0, my_cell = 0, urg_lane = 0;
        const size_t obs_off = (size_t)env0 * env_bytes * (INTERLEAVED ? 2 : 1) + ((MODE == MODE_ROLLOUT) ? (size_t)ts * (size_t)p.obs_step_stride : 0);
        if (INTERLEAVED) {
            // Both views of an env side by side: row e of the output is [agent_0's view | agent_1's view].  The staging image
            // holds epw / 2 envs x 2 views, so a tile leaves in two halves, each one bulk copy of the usual size; a deviating
            // scan cell is staged twice (once per view), the agent cells once per view with the viewer bit as that view sees it.
            const int urg = (p.max_steps - t_out) < 40;                                       // :311
            if (active) s_ap[lane] = (uint32_t)urg;
            const int half = epw >> 1;
            const int code_v0 = kCodeAgent + 16 * ai + (ndi | (inv_code(ninv) << 2));
            for (int hf = 0; hf < 2; ++hf) {
                const int e_lo = hf * half;
                const int nv_h = min(half, n_valid - e_lo);
                if (nv_h <= 0) break;
                if (p.tmpl_bytes) {
                    if (bulk_pending) {
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        bulk_pending = false;
                        __syncwarp();
                    }
                    const int n_chunks = (2 * nv_h * env_bytes + 15) >> 4;
                    uint4 *dst = reinterpret_cast<uint4 *>(stage);
                    int src = lane;
                    src -= (int)fastdiv((unsigned)src, p.magic_tc) * p.tmpl_chunks;
                    const int step = 32 - (int)fastdiv(32u, p.magic_tc) * p.tmpl_chunks;
                    for (int f = lane; f < n_chunks; f += 32) {
                        dst[f] = s_tmpl[src];
                        src += step;
                        if (src >= p.tmpl_chunks) src -= p.tmpl_chunks;
                    }
                } else {
                    if (hf) refill_stage();   // (the first half's image was requested at the top of the step)
                    uint32_t ok = 0;
                    while (!ok) {
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(tbar), "r"(tphase) : "memory");
                    }
                    tphase ^= 1u;
                }
                __syncwarp();
                {
                    const int total = nv_h * p.n_scan;
                    for (int f = lane; f < total; f += 32) {
                        const int el_l = (int)fastdiv((unsigned)f, p.magic_ns);
                        const int el = e_lo + el_l;
                        const uint32_t sc = s_scan[f - el_l * p.n_scan];
                        const unsigned char *c = smap + el * span_stride + ((m0 + el * p.map_bytes) & 15) + (int)(sc & 0xffffu);
                        const int o = c[0], st = c[2];
                        if (o == 8 ? (st != 23) : (o != 2 && o != 10)) {
                            const int code = (o == 8) ? kCodePot + min(st, 23) : min(o, 10);
                            stage_record((2 * el_l) * C + (int)(sc >> 16), code);
                            stage_record((2 * el_l + 1) * C + (int)(sc >> 16), code);
                        }
                    }
                }
                __syncwarp();
                const bool mine = active && es >= e_lo && es < e_lo + nv_h;
                if (mine) {
                    const int el_l = es - e_lo, cell = ny * w + nx;
                    stage_record((2 * el_l) * C + cell, code_v0);
                    stage_record((2 * el_l + 1) * C + cell, ai ? code_v0 - 16 : code_v0 + 16);   // agent_1 is the viewer: "which" flips
                }
                __syncwarp();
                {
                    const unsigned urg_mask = __ballot_sync(0xffffffffu, mine && urg && ai == 0);
                    if (urg_mask) {
                        const int n_cells = 2 * nv_h * C;
                        for (int g = lane; g < n_cells; g += 32) {
                            const int slot = (int)fastdiv((unsigned)g, p.magic_C);
                            if (s_ap[2 * (e_lo + (slot >> 1))]) stage[g * OC_OBS_CHANNELS + 25] = 1;
                        }
                    }
                }
                const size_t off_h = obs_off + (size_t)e_lo * 2 * env_bytes;
                if (p.bulk_ok && nv_h == half) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        bulk_store(p.obs0 + off_h, smem_u32(stage), view_bytes);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    bulk_pending = true;
                } else {
                    __syncwarp();
                    const int nb = 2 * nv_h * env_bytes;
                    const unsigned short *h0 = reinterpret_cast<const unsigned short *>(stage);
                    unsigned short *g0 = reinterpret_cast<unsigned short *>(p.obs0 + off_h);
                    for (int k = lane; k < (nb >> 1); k += 32) g0[k] = h0[k];
                    __syncwarp();
                }
            }
        } else {
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
                    if (CELLS)     // the compact form of this observation (include/oc_b200.h, oc_step_extras.cells)
                        s_cells[el * p.cell_stride + (f - el * p.n_scan)] =
                            (uint8_t)((o == 8 ? (st != 23) : (o != 2 && o != 10)) ? ((o == 8) ? kCodePot + min(st, 23) : min(o, 10)) : 0xff);
                }
            }
            __syncwarp();
            // (4c) agent cells show the agent, its direction and what it holds (:313-323, :345-353)
            if (active) stage_record(es * C + ny * w + nx, kCodeAgent + 16 * ai + (ndi | (inv_code(ninv) << 2)));
            if (CELLS) {
                if (active) {
                    unsigned char *cl = s_cells + es * p.cell_stride + p.cell_agents;
                    const int code = kCodeAgent + 16 * ai + (ndi | (inv_code(ninv) << 2));
                    *reinterpret_cast<unsigned short *>(cl + 2 * ai) = (unsigned short)((ny * w + nx) | (code << 8));
                    if (ai == 0) cl[4] = urg ? 1 : 0;
                }
                __syncwarp();
                // the tile's compact observations are contiguous in global memory: 16-byte stores
                const uint4 *src = reinterpret_cast<const uint4 *>(s_cells);
                uint4 *dst = reinterpret_cast<uint4 *>(p.cells + (env0 + ((MODE == MODE_ROLLOUT) ? (long long)ts * p.N : 0)) * p.cell_stride);
                for (int i = lane; i < (n_valid * p.cell_stride) >> 4; i += 32) dst[i] = src[i];
            }
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
                    bulk_store(p.obs0 + obs_off, smem_u32(stage), view_bytes);
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
            if (full_writeback && last_ts) {   // whole interiors: after a reset (rare: 1 step in max_steps) / after a rollout's last step
                for (int el = 0; el < n_valid; ++el) {
                    const int mis = (m0 + el * p.map_bytes) & 15;
                    const unsigned char *src = smap + el * span_stride;          // slot and HBM agree modulo 16
                    unsigned char *dstg = g_tile + el * p.map_bytes - mis;
                    for (int o = (mis & ~3) + 4 * lane; o < mis + p.span_len; o += 128)
                        *reinterpret_cast<uint32_t *>(dstg + o) = *reinterpret_cast<const uint32_t *>(src + o);
                }
            }
            if (active && last_ts) {
                reinterpret_cast<uint2 *>(p.st.agent_pos)[env * 2 + ai] = make_uint2((uint32_t)nx, (uint32_t)ny);
                reinterpret_cast<char2 *>(p.st.agent_dir)[env * 2 + ai] = make_char2((signed char)dir_x(ndi), (signed char)dir_y(ndi));
                p.st.agent_dir_idx[env * 2 + ai] = ndi;
                p.st.agent_inv[env * 2 + ai] = ninv;
                if (ai == 0) {
                    p.st.time[env] = t_out;
                    p.st.terminal[env] = (uint8_t)term_out;   // stepped: done is False; reset: False (overcooked.py:242)
                }
            }
            if (STEPPING) {
                int st_rr = 0, st_rl = 0;
                const long long oi = env + ((MODE == MODE_ROLLOUT) ? (long long)ts * p.N : 0);   // row of the per-step outputs
                if (MODE == MODE_ROLLOUT && p.res8) {
                    // packed results (oc_rollout_io.results8): reward / 20 | shaped0 class << 2 | shaped1 class << 4 | done << 6
                    const int sc = shaped == 0 ? 0 : (shaped == 3 ? 1 : 2);
                    const int osc = __shfl_xor_sync(0xffffffffu, sc, 1);
                    if (active && ai == 0) p.res8[oi] = (uint8_t)((rew / 20) | (sc << 2) | (osc << 4) | ((int)done << 6));
                } else if (active) {
                    if (ai == 0) { p.reward[oi] = (float)rew; p.done[oi] = (uint8_t)done; }
                    (ai ? p.shaped1 : p.shaped0)[oi] = (float)shaped;
                }
                if (active) {
                    if (p.ex.episode_returns) {   // wrappers/baselines.py:90-106, this lane = agent ai
                        const long long j = env * 2 + ai;
                        // pin the four loaded trackers here: without this the compiler hoists their first use right
                        // behind the loads and the warp stalls on DRAM before the observation phase
                        asm volatile("" : "+f"(lw_er), "+f"(lw_el), "+f"(lw_rr), "+f"(lw_rl));
                        const float keep = done ? 0.f : 1.f, dn = done ? 1.f : 0.f;
                        const float nr = lw_er + (float)rew;
                        const float nl = lw_el + 1.f;
                        const float rr = lw_rr * keep + nr * dn;
                        const float rl = lw_rl * keep + nl * dn;
                        lw_er = nr * keep; lw_el = nl * keep; lw_rr = rr; lw_rl = rl;   // MODE_ROLLOUT: carried to the next step
                        if (last_ts) {
                            p.ex.episode_returns[j] = lw_er;
                            p.ex.episode_lengths[j] = lw_el;
                            p.ex.returned_episode_returns[j] = rr;
                            p.ex.returned_episode_lengths[j] = rl;
                        }
                        if (p.ex.returned_episode) p.ex.returned_episode[oi * 2 + ai] = (uint8_t)done;
                        if (ai == 0 && done) { st_rr = (int)rr; st_rl = (int)rl; }
                    } else if (p.ex.returned_episode) {
                        p.ex.returned_episode[oi * 2 + ai] = (uint8_t)done;
                    }
                }
                if (p.ex.stats) {   // exact integer sums: warp-reduced per tile, one atomic per non-zero counter
                    int v_done = (active && ai == 0 && done) ? 1 : 0;
                    int v_rew = (active && ai == 0) ? rew : 0;
                    int v_s0 = (active && ai == 0) ? shaped : 0;
                    int v_s1 = (active && ai == 1) ? shaped : 0;
                    if (MODE == MODE_ROLLOUT) {   // summed per lane over the T steps, reduced once per tile
                        acc_done += v_done; acc_rew += v_rew; acc_s0 += v_s0; acc_s1 += v_s1; acc_rr += st_rr; acc_rl += st_rl;
                        v_done = acc_done; v_rew = acc_rew; v_s0 = acc_s0; v_s1 = acc_s1; st_rr = acc_rr; st_rl = acc_rl;
                    }
                    if (last_ts) {
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
        }
        if (obs1_pending) {   // agent_1's view: re-stage the two agent cells once the engine has read the image
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            if (active) stage_record(my_cell, code_v1);
            if (urg_lane && active) stage[my_cell * OC_OBS_CHANNELS + 25] = 1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                bulk_store(p.obs1 + obs_off, smem_u32(stage), view_bytes);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            bulk_pending = true;
        }
        __syncwarp();
        if (MODE == MODE_ROLLOUT) {   // the next step starts from what this one produced (agent_dir is DIR_TO_VEC[dir_idx], :434-435)
            px = nx; py = ny; di = ndi; inv = ninv; t = t_out;
            dx = dir_x(ndi); dy = dir_y(ndi);
            if (act_ring) {
                asm volatile("cp.async.wait_group 2;" ::: "memory");
                if (!last_ts && active) nact = (int)s_act[((ts + 1) & (ACT_RING - 1)) * 32 + lane];
            }
            act = clampi(nact, 0, 5);
        }
      }   // ts
        if (SPAN_BUFS == 1 && tile + tile_stride < p.n_tiles) {   // single span buffer: fetch the next tile once this one is stored
            __syncwarp();
            prefetch_spans(tile + tile_stride, 0);
            nf = load_fields(tile + tile_stride);
        }
    }
    if (STEPPING && p.ex.stats && blockIdx.x == 0 && tid == 0)   // every env stepped T times
        atomicAdd(reinterpret_cast<unsigned long long *>(p.ex.stats) + 6, (unsigned long long)p.N * (unsigned long long)T);
    if (bulk_pending && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}


// oc_rollout pre-pass: int32 actions of both agents -> one byte per env and step (the kernel's own clamp to 0..5 applied here)
__global__ void oc_pack_actions_kernel(const int32_t *__restrict__ act0, const int32_t *__restrict__ act1, long long n,
                                       uint8_t *__restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 3 < n && ((reinterpret_cast<uintptr_t>(act0 + i) | reinterpret_cast<uintptr_t>(act1 + i)) & 15) == 0 &&
            (reinterpret_cast<uintptr_t>(out + i) & 3) == 0) {
            const int4 a = __ldg(reinterpret_cast<const int4 *>(act0 + i)), b = __ldg(reinterpret_cast<const int4 *>(act1 + i));
            const uint32_t v = (uint32_t)(clampi(a.x, 0, 5) | (clampi(b.x, 0, 5) << 4)) |
                               ((uint32_t)(clampi(a.y, 0, 5) | (clampi(b.y, 0, 5) << 4)) << 8) |
                               ((uint32_t)(clampi(a.z, 0, 5) | (clampi(b.z, 0, 5) << 4)) << 16) |
                               ((uint32_t)(clampi(a.w, 0, 5) | (clampi(b.w, 0, 5) << 4)) << 24);
            *reinterpret_cast<uint32_t *>(out + i) = v;
        } else {
            for (long long j = i; j < n && j < i + 4; ++j)
                out[j] = (uint8_t)(clampi(__ldg(act0 + j), 0, 5) | (clampi(__ldg(act1 + j), 0, 5) << 4));
        }
    }
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
