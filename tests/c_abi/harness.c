/* Plain-C caller of the C ABI in include/oc_b200.h -- no Python between this program and liboc_b200.so.
 *
 *   harness IN OUT
 *
 * IN  (written by tests/test_gpu_c_harness.py from a golden fixture): int32 header
 *       magic 0x0C0C0C0C, h, w, n_wall, n_goal, n_pot, n_onion_pile, n_plate_pile, random_reset, max_steps, task_id,
 *       prng_mode, E, T
 *     then int32 wall_idx[n_wall], agent_idx[2], goal_idx[n_goal], pot_idx[n_pot], onion_pile_idx[], plate_pile_idx[];
 *     uint32 reset_keys[E,2]; uint32 step_keys[T,E,2]; int32 actions[T,E,2]
 * OUT section A (oc_layout_create -> oc_reset -> T x oc_step, state passed BY VALUE as the header declares it):
 *       reset: obs0, obs1, then the dynamic State fields; per step: obs0, obs1, reward, shaped0, shaped1, done, State fields
 *     section B (oc_reset again, then ONE oc_rollout over the same keys / actions): per step obs0, obs1, reward, shaped0,
 *       shaped1, done; then the final State fields; then oc_get_obs of the final state (obs0, obs1)
 * The test byte-compares OUT with the arrays of the golden fixture (recorded from the reference's own source). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "oc_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 3; } } while (0)
#define OC(x) do { int s_ = (x); if (s_ != OC_OK) { fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, oc_status_string(s_)); return 4; } } while (0)

static int32_t *read_i32(FILE *f, size_t n) {
    int32_t *p = (int32_t *)malloc((n ? n : 1) * 4);
    if (n && fread(p, 4, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
    return p;
}

static void *dmalloc(size_t n) {
    void *p = NULL;
    if (cudaMalloc(&p, n ? n : 16) != cudaSuccess) { fprintf(stderr, "cudaMalloc(%zu) failed\n", n); exit(3); }
    return p;
}

static int dump(FILE *out, const void *dptr, size_t n) {
    void *h = malloc(n ? n : 1);
    CK(cudaMemcpy(h, dptr, n, cudaMemcpyDeviceToHost));
    fwrite(h, 1, n, out);
    free(h);
    return 0;
}

typedef struct { size_t pos, dir, diri, inv, goal, pot, wall, maze, time, term, task; } sizes_t;

static int dump_state(FILE *out, const oc_state *s, const sizes_t *z) {
    if (dump(out, s->agent_pos, z->pos) || dump(out, s->agent_dir, z->dir) || dump(out, s->agent_dir_idx, z->diri) ||
        dump(out, s->agent_inv, z->inv) || dump(out, s->maze_map, z->maze) || dump(out, s->time, z->time) ||
        dump(out, s->terminal, z->term))
        return 1;
    return 0;
}

int main(int argc, char **argv) {
    if (argc != 3) { fprintf(stderr, "usage: harness IN OUT\n"); return 1; }
    FILE *in = fopen(argv[1], "rb");
    FILE *out = fopen(argv[2], "wb");
    if (!in || !out) { fprintf(stderr, "cannot open files\n"); return 1; }
    int32_t *hd = read_i32(in, 14);
    if ((uint32_t)hd[0] != 0x0C0C0C0Cu) { fprintf(stderr, "bad magic\n"); return 2; }
    oc_layout_desc d;
    memset(&d, 0, sizeof(d));
    d.height = hd[1]; d.width = hd[2]; d.n_wall = hd[3]; d.n_goal = hd[4]; d.n_pot = hd[5];
    d.n_onion_pile = hd[6]; d.n_plate_pile = hd[7]; d.random_reset = hd[8]; d.max_steps = hd[9]; d.task_id = hd[10];
    d.prng_mode = hd[11];
    const int64_t E = hd[12];
    const int T = hd[13];
    d.wall_idx = read_i32(in, (size_t)d.n_wall);
    d.agent_idx = read_i32(in, 2);
    d.goal_idx = read_i32(in, (size_t)d.n_goal);
    d.pot_idx = read_i32(in, (size_t)d.n_pot);
    d.onion_pile_idx = read_i32(in, (size_t)d.n_onion_pile);
    d.plate_pile_idx = read_i32(in, (size_t)d.n_plate_pile);
    uint32_t *reset_keys = (uint32_t *)read_i32(in, (size_t)E * 2);
    uint32_t *step_keys = (uint32_t *)read_i32(in, (size_t)T * E * 2);
    int32_t *actions = read_i32(in, (size_t)T * E * 2);
    fclose(in);

    CK(cudaSetDevice(0));
    if (oc_version() != OC_B200_VERSION) { fprintf(stderr, "header / library version mismatch\n"); return 5; }
    oc_layout *lay = NULL;
    OC(oc_layout_create(&d, &lay));
    oc_layout_info info;
    OC(oc_layout_query(lay, &info));
    const int h = info.height, w = info.width;
    const size_t obs_bytes = (size_t)E * (size_t)info.obs_bytes_per_agent;
    sizes_t z;
    z.pos = (size_t)E * 16; z.dir = (size_t)E * 4; z.diri = (size_t)E * 8; z.inv = (size_t)E * 8;
    z.goal = (size_t)E * 8 * (size_t)info.n_goal; z.pot = (size_t)E * 8 * (size_t)info.n_pot; z.wall = (size_t)E * h * w;
    z.maze = (size_t)E * (size_t)info.map_bytes_per_env; z.time = (size_t)E * 4; z.term = (size_t)E; z.task = (size_t)E * 4;
    oc_state st;
    st.agent_pos = (uint32_t *)dmalloc(z.pos); st.agent_dir = (int8_t *)dmalloc(z.dir);
    st.agent_dir_idx = (int32_t *)dmalloc(z.diri); st.agent_inv = (int32_t *)dmalloc(z.inv);
    st.goal_pos = (uint32_t *)dmalloc(z.goal); st.pot_pos = (uint32_t *)dmalloc(z.pot);
    st.wall_map = (uint8_t *)dmalloc(z.wall); st.maze_map = (uint8_t *)dmalloc(z.maze);
    st.time = (int32_t *)dmalloc(z.time); st.terminal = (uint8_t *)dmalloc(z.term); st.task_id = (int32_t *)dmalloc(z.task);

    uint32_t *d_rk = (uint32_t *)dmalloc((size_t)E * 8), *d_sk = (uint32_t *)dmalloc((size_t)T * E * 8);
    int32_t *d_a0 = (int32_t *)dmalloc((size_t)T * E * 4), *d_a1 = (int32_t *)dmalloc((size_t)T * E * 4);
    int32_t *a0 = (int32_t *)malloc((size_t)T * E * 4 + 4), *a1 = (int32_t *)malloc((size_t)T * E * 4 + 4);
    for (int64_t i = 0; i < (int64_t)T * E; ++i) { a0[i] = actions[2 * i]; a1[i] = actions[2 * i + 1]; }
    CK(cudaMemcpy(d_rk, reset_keys, (size_t)E * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sk, step_keys, (size_t)T * E * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_a0, a0, (size_t)T * E * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_a1, a1, (size_t)T * E * 4, cudaMemcpyHostToDevice));
    uint8_t *obs = (uint8_t *)dmalloc(2 * obs_bytes * (size_t)(T > 0 ? T : 1));      /* [T, 2, E, h, w, 26] */
    float *rew = (float *)dmalloc((size_t)(T ? T : 1) * E * 4), *sh0 = (float *)dmalloc((size_t)(T ? T : 1) * E * 4),
          *sh1 = (float *)dmalloc((size_t)(T ? T : 1) * E * 4);
    uint8_t *done = (uint8_t *)dmalloc((size_t)(T ? T : 1) * E);
    cudaStream_t stream;
    CK(cudaStreamCreate(&stream));

    /* ---- section A: reset, then T single steps (state in == state out: stepped in place) */
    OC(oc_reset(lay, E, d_rk, st, obs, obs + obs_bytes, stream));
    CK(cudaStreamSynchronize(stream));
    if (dump(out, obs, obs_bytes) || dump(out, obs + obs_bytes, obs_bytes) || dump_state(out, &st, &z)) return 3;
    for (int t = 0; t < T; ++t) {
        OC(oc_step(lay, E, d_sk + (size_t)t * E * 2, st, d_a0 + (size_t)t * E, d_a1 + (size_t)t * E, NULL, st, obs,
                   obs + obs_bytes, rew, sh0, sh1, done, NULL, stream));
        CK(cudaStreamSynchronize(stream));
        if (dump(out, obs, obs_bytes) || dump(out, obs + obs_bytes, obs_bytes) || dump(out, rew, (size_t)E * 4) ||
            dump(out, sh0, (size_t)E * 4) || dump(out, sh1, (size_t)E * 4) || dump(out, done, (size_t)E) ||
            dump_state(out, &st, &z))
            return 3;
    }
    /* ---- section B: reset again, ONE oc_rollout over the same T steps */
    if (T > 0) {
        OC(oc_reset(lay, E, d_rk, st, obs, obs + obs_bytes, stream));
        oc_rollout_io io;
        memset(&io, 0, sizeof(io));
        io.act0 = d_a0; io.act1 = d_a1; io.obs0 = obs; io.obs1 = obs + obs_bytes; io.obs_step_stride = (int64_t)(2 * obs_bytes);
        io.reward = rew; io.shaped0 = sh0; io.shaped1 = sh1; io.done = done;
        OC(oc_rollout(lay, E, T, d_sk, st, &io, NULL, stream));
        CK(cudaStreamSynchronize(stream));
        for (int t = 0; t < T; ++t) {
            if (dump(out, obs + (size_t)t * 2 * obs_bytes, obs_bytes) || dump(out, obs + (size_t)t * 2 * obs_bytes + obs_bytes, obs_bytes) ||
                dump(out, rew + (size_t)t * E, (size_t)E * 4) || dump(out, sh0 + (size_t)t * E, (size_t)E * 4) ||
                dump(out, sh1 + (size_t)t * E, (size_t)E * 4) || dump(out, done + (size_t)t * E, (size_t)E))
                return 3;
        }
        if (dump_state(out, &st, &z)) return 3;
        OC(oc_get_obs(lay, E, st, obs, obs + obs_bytes, stream));
        CK(cudaStreamSynchronize(stream));
        if (dump(out, obs, obs_bytes) || dump(out, obs + obs_bytes, obs_bytes)) return 3;
    }
    fclose(out);
    oc_layout_destroy(lay);
    printf("harness ok: E=%lld T=%d %dx%d epw=%d\n", (long long)E, T, h, w, info.envs_per_warp);
    return 0;
}
