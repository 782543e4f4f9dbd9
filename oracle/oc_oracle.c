/*
 * oc_oracle.c -- TEST INFRASTRUCTURE ONLY.  Plain-C CPU restatement of the reference's batched Overcooked
 * environment step, used as the parity checker for the CUDA path and as the timed CPU baseline.
 *
 * Nothing here is product code.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library (through oracle/oracle.py).
 *
 * Reference (LukavdBoogaard/JAXOvercooked) lines each function follows -- paths relative to
 * jax_marl/environments/:
 *   oco_threefry / split / bits / randint    [JAX-internal, not vendored; jax unpinned, hint jaxlib==0.4.25
 *                                             (requirements.txt:3-4)] restated from jax/_src/prng.py and
 *                                             jax/_src/random.py; see oracle/prng.py for the NumPy twin
 *   make_map()          overcooked_environment/common.py:76-170
 *   reset_one()         overcooked_environment/overcooked.py:136-248
 *   get_obs_one()       overcooked_environment/overcooked.py:250-364
 *   step_agents()       overcooked_environment/overcooked.py:366-514
 *   process_interact()  overcooked_environment/overcooked.py:516-642
 *   step_one()          overcooked_environment/overcooked.py:106-134 + multi_agent_env.py:41-69 (auto-reset)
 *   log_wrapper_one()   ../wrappers/baselines.py:75-107
 *
 * PARITY STATUS: the environment logic is pinned against golden trajectories recorded by executing the
 * reference's own source files (tests/golden/make_golden.py, through a NumPy stand-in for jax because JAX
 * cannot be installed in this image).  The PRNG layer is pinned only by Random123 / JAX-documentation
 * known-answer vectors: random draws are "parity unpinned" against a real JAX install.
 *
 * Batched arrays use the same struct-of-arrays layout as include/oc_b200.h (leading N).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OCO_MAX_CELLS 256
#define OCO_MAX_SPECIAL 16
#define PAD 4

typedef struct {
    int32_t height, width;
    int32_t n_wall, n_goal, n_pot, n_onion_pile, n_plate_pile;
    int32_t wall_idx[OCO_MAX_CELLS * 2];
    int32_t agent_idx[2];
    int32_t goal_idx[OCO_MAX_SPECIAL];
    int32_t pot_idx[OCO_MAX_SPECIAL];
    int32_t onion_pile_idx[OCO_MAX_SPECIAL];
    int32_t plate_pile_idx[OCO_MAX_SPECIAL];
    int32_t random_reset, max_steps, task_id, partitionable;
} oco_layout;

typedef struct {
    uint32_t *agent_pos;     /* [N,2,2] (x,y) */
    int8_t   *agent_dir;     /* [N,2,2] */
    int32_t  *agent_dir_idx; /* [N,2] */
    int32_t  *agent_inv;     /* [N,2] */
    uint32_t *goal_pos;      /* [N,G,2] */
    uint32_t *pot_pos;       /* [N,P,2] */
    uint8_t  *wall_map;      /* [N,h,w] */
    uint8_t  *maze_map;      /* [N,h+8,w+8,3] */
    int32_t  *time;          /* [N] */
    uint8_t  *terminal;      /* [N] */
    int32_t  *task_id;       /* [N] */
} oco_state;

/* one environment's slices */
typedef struct {
    uint32_t *pos; int8_t *dir; int32_t *dir_idx; int32_t *inv;
    uint32_t *goal_pos; uint32_t *pot_pos; uint8_t *wall_map; uint8_t *maze;
    int32_t *time; uint8_t *terminal; int32_t *task_id;
} env_t;

static const uint8_t OBJ_VEC[11][3] = { /* common.py:51-63 */
    {0, 0, 0}, {1, 0, 0}, {2, 5, 0}, {3, 4, 0}, {4, 4, 0}, {5, 6, 0},
    {6, 6, 0}, {7, 1, 0}, {8, 7, 0}, {9, 6, 0}, {10, 0, 0}};
static const int8_t DIR_VEC[4][2] = {{0, -1}, {0, 1}, {1, 0}, {-1, 0}}; /* common.py:67-73 */

/* ------------------------------------------------------------------------------------------- PRNG */

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

void oco_threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t *y0, uint32_t *y1) {
    static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
    uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    x0 += ks[0];
    x1 += ks[1];
    for (int i = 0; i < 5; ++i) {
        for (int j = 0; j < 4; ++j) {
            x0 += x1;
            x1 = rotl32(x1, R[i % 2][j]);
            x1 ^= x0;
        }
        x0 += ks[(i + 1) % 3];
        x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
    }
    *y0 = x0;
    *y1 = x1;
}

/* random_bits(key, n) for 32-bit words */
static void bits32(const uint32_t key[2], int n, int partitionable, uint32_t *out) {
    if (partitionable) {
        for (int j = 0; j < n; ++j) {
            uint32_t a, b;
            oco_threefry2x32(key[0], key[1], 0u, (uint32_t)j, &a, &b);
            out[j] = a ^ b;
        }
        return;
    }
    int m = (n + 1) / 2; /* counters iota(n) (+ one 0 if n is odd) split in two halves */
    for (int j = 0; j < m; ++j) {
        uint32_t c1 = (j + m < n) ? (uint32_t)(j + m) : 0u;
        uint32_t a, b;
        oco_threefry2x32(key[0], key[1], (uint32_t)j, c1, &a, &b);
        out[j] = a;
        if (j + m < n) out[j + m] = b;
    }
}

void oco_split(const uint32_t key[2], int num, int partitionable, uint32_t *out /*[num,2]*/) {
    if (partitionable) {
        for (int i = 0; i < num; ++i)
            oco_threefry2x32(key[0], key[1], 0u, (uint32_t)i, &out[2 * i], &out[2 * i + 1]);
        return;
    }
    bits32(key, 2 * num, 0, out);
}

/* key, subkey = split(key) */
static void split2(uint32_t key[2], uint32_t sub[2], int partitionable) {
    uint32_t o[4];
    oco_split(key, 2, partitionable, o);
    key[0] = o[0]; key[1] = o[1];
    sub[0] = o[2]; sub[1] = o[3];
}

void oco_randint(const uint32_t key[2], int n, int lo, int hi, int partitionable, int32_t *out) {
    uint32_t k[4];
    oco_split(key, 2, partitionable, k);
    uint32_t *hb = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(2 * n + 2));
    uint32_t *lb = hb + n + 1;
    bits32(&k[0], n, partitionable, hb);
    bits32(&k[2], n, partitionable, lb);
    uint32_t span = (uint32_t)(hi - lo);
    uint32_t mult = 65536u % span;
    mult = (mult * mult) % span;
    for (int i = 0; i < n; ++i) {
        uint32_t off = (hb[i] % span) * mult + (lb[i] % span);
        out[i] = lo + (int32_t)(off % span);
    }
    free(hb);
}

/* choice(key, arange(n), (2,), p=free, replace=False) in its integer form:
   largest 23-bit mantissa first, ties to the lower index */
static void choice2_free(const uint32_t key[2], int n, const uint8_t *is_wall, int partitionable, int out[2]) {
    uint32_t bits[OCO_MAX_CELLS];
    bits32(key, n, partitionable, bits);
    int64_t best = -1, second = -1;
    int bi = 0, si = 0;
    for (int i = 0; i < n; ++i) {
        if (is_wall[i]) continue;
        int64_t m = (int64_t)(bits[i] >> 9);
        if (m > best) { second = best; si = bi; best = m; bi = i; }
        else if (m > second) { second = m; si = i; }
    }
    out[0] = bi;
    out[1] = si;
}

/* ------------------------------------------------------------------------------------ map / reset */

static inline uint8_t *cell_at(const oco_layout *L, uint8_t *maze, int x, int y) {
    return maze + ((size_t)(PAD + y) * (L->width + 2 * PAD) + (PAD + x)) * 3;
}

static void set_cell(uint8_t *c, int a, int b, int d) { c[0] = (uint8_t)a; c[1] = (uint8_t)b; c[2] = (uint8_t)d; }

static void make_map(const oco_layout *L, env_t *e, const int *pot_status,
                     const int *onion_idx, const int *plate_idx) {
    const int h = L->height, w = L->width, HP = h + 2 * PAD, WP = w + 2 * PAD;
    for (int i = 0; i < HP * WP; ++i) set_cell(e->maze + 3 * i, 2, 5, 0);              /* common.py:158 */
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            uint8_t *c = cell_at(L, e->maze, x, y);
            if (e->wall_map[y * w + x]) set_cell(c, 2, 5, 0); else set_cell(c, 1, 0, 0); /* :93-96 */
        }
    for (int i = 0; i < 2; ++i)                                                         /* :99-105 */
        set_cell(cell_at(L, e->maze, (int)e->pos[2 * i], (int)e->pos[2 * i + 1]), 10, 2 * i, e->dir_idx[i]);
    for (int g = 0; g < L->n_goal; ++g)                                                 /* :108-111 */
        set_cell(cell_at(L, e->maze, (int)e->goal_pos[2 * g], (int)e->goal_pos[2 * g + 1]), 7, 1, 0);
    for (int k = 0; k < L->n_onion_pile; ++k)                                           /* :114-117 */
        set_cell(cell_at(L, e->maze, onion_idx[k] % w, onion_idx[k] / w), 4, 4, 0);
    for (int k = 0; k < L->n_plate_pile; ++k)                                           /* :120-123 */
        set_cell(cell_at(L, e->maze, plate_idx[k] % w, plate_idx[k] / w), 6, 6, 0);
    for (int p = 0; p < L->n_pot; ++p)                                                  /* :126-132 */
        set_cell(cell_at(L, e->maze, (int)e->pot_pos[2 * p], (int)e->pot_pos[2 * p + 1]), 8, 7, pot_status[p]);
    /* :162-168 "surrounding walls" land on ring cells that are already wall -> no-op */
}

static void reset_one(const oco_layout *L, const uint32_t key_in[2], env_t *e) {
    const int h = L->height, w = L->width, n = h * w;
    const int part = L->partitionable;
    uint32_t key[2] = {key_in[0], key_in[1]}, sub[2];

    memset(e->wall_map, 0, (size_t)n);                                                  /* overcooked.py:159-161 */
    for (int i = 0; i < L->n_wall; ++i) e->wall_map[L->wall_idx[i]] = 1;

    split2(key, sub, part);                                                             /* :164 */
    int agent_idx[2];
    choice2_free(sub, n, e->wall_map, part, agent_idx);                                 /* :165-166 */
    if (!L->random_reset) { agent_idx[0] = L->agent_idx[0]; agent_idx[1] = L->agent_idx[1]; } /* :169 */
    for (int i = 0; i < 2; ++i) {                                                       /* :170 */
        e->pos[2 * i] = (uint32_t)(agent_idx[i] % w);
        e->pos[2 * i + 1] = (uint32_t)(agent_idx[i] / w);
    }

    split2(key, sub, part);                                                             /* :173 */
    oco_randint(sub, 2, 0, 4, part, e->dir_idx);                                        /* :174 choice(arange(4)) */
    for (int i = 0; i < 2; ++i) {                                                       /* :175 */
        e->dir[2 * i] = DIR_VEC[e->dir_idx[i]][0];
        e->dir[2 * i + 1] = DIR_VEC[e->dir_idx[i]][1];
    }
    for (int g = 0; g < L->n_goal; ++g) {                                               /* :181-182 */
        e->goal_pos[2 * g] = (uint32_t)(L->goal_idx[g] % w);
        e->goal_pos[2 * g + 1] = (uint32_t)(L->goal_idx[g] / w);
    }
    for (int p = 0; p < L->n_pot; ++p) {                                                /* :193-194 */
        e->pot_pos[2 * p] = (uint32_t)(L->pot_idx[p] % w);
        e->pot_pos[2 * p + 1] = (uint32_t)(L->pot_idx[p] / w);
    }

    split2(key, sub, part);                                                             /* :197 */
    int32_t pot_status[OCO_MAX_SPECIAL];
    oco_randint(sub, L->n_pot, 0, 24, part, pot_status);                                /* :200 */
    if (!L->random_reset) for (int p = 0; p < L->n_pot; ++p) pot_status[p] = 23;        /* :201 */

    make_map(L, e, pot_status, L->onion_pile_idx, L->plate_pile_idx);                   /* :207-222 */

    split2(key, sub, part);                                                             /* :225 */
    static const int items[4] = {1, 3, 5, 9};                                           /* :226-227 */
    int32_t r[2];
    oco_randint(sub, 2, 0, 4, part, r);                                                 /* :228 */
    for (int i = 0; i < 2; ++i) e->inv[i] = L->random_reset ? items[r[i]] : 1;          /* :229-230 */

    *e->time = 0;                                                                       /* :241-243 */
    *e->terminal = 0;
    *e->task_id = L->task_id;
}

/* ------------------------------------------------------------------------------------------- obs */

static void get_obs_one(const oco_layout *L, const env_t *e, uint8_t *obs0, uint8_t *obs1) {
    const int h = L->height, w = L->width;
    const int urg = (L->max_steps - *e->time) < 40;                                     /* :311 */
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint8_t *c = cell_at(L, e->maze, x, y);
            int o = c[0];                                                               /* :301 */
            int soup_loc = (o == 9);                                                    /* :302 */
            int pot_loc = (o == 8);                                                     /* :304 */
            uint8_t ps = (uint8_t)(c[2] * pot_loc);                                     /* :305 */
            uint8_t m3 = (uint8_t)(23 - ps) < 3 ? (uint8_t)(23 - ps) : 3;
            int onions_in_pot = m3 * (ps >= 20);                                        /* :306 */
            int onions_in_soup = m3 * (ps < 20) * pot_loc + 3 * soup_loc;               /* :307-308 */
            int cook_time = ps * (ps < 20);                                             /* :309 */
            int soup_ready = pot_loc * (ps == 0) + soup_loc;                            /* :310 */
            int a0 = ((int)e->pos[0] == x && (int)e->pos[1] == y);                      /* :313-315 */
            int a1 = ((int)e->pos[2] == x && (int)e->pos[3] == y);
            int inv_items = a0 * e->inv[0] + a1 * e->inv[1];                            /* :318 */
            int on = a0 + a1;
            int oo = on ? inv_items : o;                                                /* :319 */
            soup_ready += (inv_items == 9) * on;                                        /* :320-321 */
            onions_in_soup += (inv_items == 9) * 3 * on;                                /* :322-323 */
            uint8_t v[26];
            memset(v, 0, 26);
            v[0] = (uint8_t)a0; v[1] = (uint8_t)a1;                                     /* :351 */
            if (a0) v[2 + e->dir_idx[0]] = 1;                                           /* :345-347,353 */
            if (a1) v[6 + e->dir_idx[1]] = 1;
            v[10] = (oo == 8); v[11] = (oo == 2); v[12] = (oo == 4); v[14] = (oo == 6); v[15] = (oo == 7);
            v[16] = (uint8_t)onions_in_pot; v[18] = (uint8_t)onions_in_soup; v[20] = (uint8_t)cook_time;
            v[21] = (uint8_t)soup_ready; v[22] = (oo == 5); v[23] = (oo == 3); v[25] = (uint8_t)urg;
            uint8_t *p0 = obs0 + ((size_t)y * w + x) * 26, *p1 = obs1 + ((size_t)y * w + x) * 26;
            memcpy(p0, v, 26);
            memcpy(p1, v, 26);                                                          /* :356-359 */
            p1[0] = v[1]; p1[1] = v[0];
            memcpy(p1 + 2, v + 6, 4);
            memcpy(p1 + 6, v + 2, 4);
        }
}

/* ------------------------------------------------------------------------------------------ step */

static void process_interact(const oco_layout *L, uint8_t *maze, const uint8_t *wall_map, int fx, int fy,
                             int inventory, int *inv_out, float *reward, float *shaped) {
    const int h = L->height, w = L->width;
    uint8_t *cell = cell_at(L, maze, fx, fy);                                           /* :534 */
    const int obj = cell[0];
    const int is_pile = (obj == 6) || (obj == 4);                                       /* :538 */
    const int is_pot = (obj == 8), is_goal = (obj == 7), is_agent = (obj == 10);
    const int pickable = (obj == 5) || (obj == 3) || (obj == 9);                        /* :542-545 */
    const int is_table = wall_map[fy * w + fx] && !is_pot;                              /* :547 */
    const int table_empty = (obj == 2) || (obj == 1);                                   /* :549 */
    const uint8_t status = cell[2];                                                     /* :552 */
    const int inv_empty = (inventory == 1);
    const int case1 = (status > 20) && (inventory == 3) && is_pot;                      /* :562-565 */
    const int case2 = (status == 0) && (inventory == 5) && is_pot;
    const int case3 = (status > 0) && (status <= 20) && is_pot;
    const int else_case = !case1 && !case2 && !case3;
    float sh = 0.f;
    sh += case1 * 3;                                                                    /* :568-569 */
    sh += case2 * 5;
    const uint8_t new_status = (uint8_t)(case1 * (uint8_t)(status - 1) + case2 * 23 + case3 * status +
                                         else_case * status);                           /* :572-576 */
    int new_inv = case1 * 1 + case2 * 9 + case3 * inventory + else_case * inventory;    /* :577-581 */
    const int pickup = is_table && !table_empty && inv_empty && (is_pile || pickable);  /* :585 */
    const int drop = is_table && table_empty && !inv_empty;                             /* :586 */
    const int delivery = is_table && is_goal && (inventory == 9);                       /* :587 */
    const int no_effect = !pickup && !drop && !delivery;
    const int new_obj = no_effect * obj + delivery * obj + pickup * is_pile * obj + pickup * pickable * 2 +
                        drop * inventory;                                               /* :591-596 */
    new_inv = no_effect * new_inv + delivery * 1 + pickup * pickable * obj + pickup * (obj == 6) * 5 +
              pickup * (obj == 4) * 3 + drop * 1;                                       /* :599-605 */
    const int drop_occurred = (inventory != 1) && (new_inv == 1);                       /* :608-613 */
    sh += (drop_occurred && (new_obj == inventory) && drop) * 0;
    const int picked_plate = pickup && (new_inv == 5);                                  /* :616 */
    const int plates_in_inv = (inventory == 5);                                         /* :619 */
    int notempty_pots = 0;                                                              /* :620-622 */
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint8_t *c = cell_at(L, maze, x, y);
            notempty_pots += (c[0] == 8) && (c[2] != 23);
        }
    const int useful = plates_in_inv < notempty_pots;                                   /* :623 */
    long plate_cells = 0;                                                               /* :625-626: the WHOLE padded
                                                      3-channel map is compared with 5, colour bytes included */
    const long nbytes = (long)(h + 2 * PAD) * (w + 2 * PAD) * 3;
    for (long i = 0; i < nbytes; ++i) plate_cells += (maze[i] == 5);
    sh += ((plate_cells == 0) && picked_plate && useful) * 3;                           /* :628 */

    if (is_pot) { set_cell(cell, OBJ_VEC[new_obj][0], OBJ_VEC[new_obj][1], new_status); }  /* :633-638 */
    else if (!is_agent) { set_cell(cell, OBJ_VEC[new_obj][0], OBJ_VEC[new_obj][1], OBJ_VEC[new_obj][2]); }
    *inv_out = new_inv;
    *reward = (float)delivery * 20.f;                                                   /* :641 */
    *shaped = sh;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void step_agents(const oco_layout *L, env_t *e, const int action[2], float *reward, float shaped[2]) {
    const int h = L->height, w = L->width;
    int is_move[2], fwd[2][2], prev[2][2], bounced[2];
    for (int i = 0; i < 2; ++i) {
        prev[i][0] = (int)e->pos[2 * i];
        prev[i][1] = (int)e->pos[2 * i + 1];
        is_move[i] = (action[i] != 4) && (action[i] != 5);                              /* :371 */
        const int a3 = action[i] < 3 ? action[i] : 3;
        for (int k = 0; k < 2; ++k) {                                                   /* :374-378 */
            int v = prev[i][k] + is_move[i] * DIR_VEC[a3][k] + (!is_move[i]) * e->dir[2 * i + k];
            fwd[i][k] = clampi(v, 0, (k == 0 ? w : h) - 1);
        }
        int wall = e->wall_map[fwd[i][1] * w + fwd[i][0]];                              /* :381-391 */
        int goal = 0;
        for (int g = 0; g < L->n_goal; ++g)
            goal |= (fwd[i][0] == (int)e->goal_pos[2 * g]) && (fwd[i][1] == (int)e->goal_pos[2 * g + 1]);
        bounced[i] = wall || goal || !is_move[i];                                       /* :393 */
        if (bounced[i]) { fwd[i][0] = prev[i][0]; fwd[i][1] = prev[i][1]; }             /* :398 */
    }
    const int collision = (fwd[0][0] == fwd[1][0]) && (fwd[0][1] == fwd[1][1]);         /* :399 */
    const int swap = (fwd[0][0] == prev[1][0]) && (fwd[0][1] == prev[1][1]) &&
                     (fwd[1][0] == prev[0][0]) && (fwd[1][1] == prev[0][1]);            /* :414-417 */
    int np_[2][2];
    for (int i = 0; i < 2; ++i)
        for (int k = 0; k < 2; ++k)
            np_[i][k] = (collision || (!collision && swap)) ? prev[i][k] : fwd[i][k];   /* :402-427 */

    int new_dir_idx[2];
    for (int i = 0; i < 2; ++i) new_dir_idx[i] = is_move[i] ? action[i] : e->dir_idx[i]; /* :434 */

    /* interact: OLD position + OLD direction vector, Alice then Bob (:439-465) */
    int inv_new[2] = {e->inv[0], e->inv[1]};
    float rew[2] = {0.f, 0.f};
    shaped[0] = shaped[1] = 0.f;
    for (int i = 0; i < 2; ++i) {
        if (action[i] != 5) continue;  /* select(interact, candidate, unchanged) == do nothing */
        const int fx = prev[i][0] + e->dir[2 * i], fy = prev[i][1] + e->dir[2 * i + 1];
        process_interact(L, e->maze, e->wall_map, fx, fy, e->inv[i], &inv_new[i], &rew[i], &shaped[i]);
    }
    e->inv[0] = inv_new[0];                                                             /* :468 */
    e->inv[1] = inv_new[1];

    for (int i = 0; i < 2; ++i) set_cell(cell_at(L, e->maze, prev[i][0], prev[i][1]), 1, 0, 0); /* :485 */
    for (int i = 0; i < 2; ++i)                                                         /* :486 */
        set_cell(cell_at(L, e->maze, np_[i][0], np_[i][1]), 10, 2 * i, new_dir_idx[i]);

    for (int p = 0; p < L->n_pot; ++p) {                                                /* :488-500 */
        uint8_t *c = cell_at(L, e->maze, (int)e->pot_pos[2 * p], (int)e->pot_pos[2 * p + 1]);
        const uint8_t s = c[2];
        const int is_cooking = s <= 20, not_done = s > 0;
        c[2] = (uint8_t)(is_cooking * not_done * (uint8_t)(s - 1) + (!is_cooking) * s);
    }
    for (int i = 0; i < 2; ++i) {                                                       /* :505-511 */
        e->pos[2 * i] = (uint32_t)np_[i][0];
        e->pos[2 * i + 1] = (uint32_t)np_[i][1];
        e->dir_idx[i] = new_dir_idx[i];
        e->dir[2 * i] = DIR_VEC[new_dir_idx[i]][0];
        e->dir[2 * i + 1] = DIR_VEC[new_dir_idx[i]][1];
    }
    *e->terminal = 0;
    *reward = rew[0] + rew[1];                                                          /* :502 */
}

static void step_one(const oco_layout *L, const uint32_t key_in[2], env_t *e, int a0, int a1,
                     uint8_t *obs0, uint8_t *obs1, float *reward, float *sh0, float *sh1, uint8_t *done) {
    uint32_t key[2] = {key_in[0], key_in[1]}, key_reset[2];
    split2(key, key_reset, L->partitionable);                                           /* multi_agent_env.py:52 */
    int act[2] = {a0, a1};                                                              /* overcooked.py:114 */
    float shaped[2];
    step_agents(L, e, act, reward, shaped);                                             /* :116 */
    *e->time += 1;                                                                      /* :118 */
    const int d = (*e->time >= L->max_steps) || *e->terminal;                           /* :120, :644-647 */
    *e->terminal = (uint8_t)d;                                                          /* :121 */
    *sh0 = shaped[0];
    *sh1 = shaped[1];
    *done = (uint8_t)d;
    /* multi_agent_env.py:55-68: the reference evaluates reset(key_reset) every step and selects by `done`;
       evaluating it only when selected gives the same outputs */
    if (d) reset_one(L, key_reset, e);
    get_obs_one(L, e, obs0, obs1);                                                      /* :123 / :246 */
}

static void log_wrapper_one(float reward, int done, float *ep_ret, float *ep_len, float *ret_ret, float *ret_len) {
    for (int i = 0; i < 2; ++i) {                                                       /* wrappers/baselines.py:90-101 */
        const float nr = ep_ret[i] + reward, nl = ep_len[i] + 1.f;
        const float keep = (float)(1 - done), d = (float)done;
        ep_ret[i] = nr * keep;
        ep_len[i] = nl * keep;
        ret_ret[i] = ret_ret[i] * keep + nr * d;
        ret_len[i] = ret_len[i] * keep + nl * d;
    }
}

/* --------------------------------------------------------------------------------- batched entry */

static env_t view(const oco_layout *L, const oco_state *s, int64_t i) {
    const int h = L->height, w = L->width;
    env_t e;
    e.pos = s->agent_pos + 4 * i;
    e.dir = s->agent_dir + 4 * i;
    e.dir_idx = s->agent_dir_idx + 2 * i;
    e.inv = s->agent_inv + 2 * i;
    e.goal_pos = s->goal_pos + (size_t)2 * L->n_goal * i;
    e.pot_pos = s->pot_pos + (size_t)2 * L->n_pot * i;
    e.wall_map = s->wall_map + (size_t)h * w * i;
    e.maze = s->maze_map + (size_t)(h + 2 * PAD) * (w + 2 * PAD) * 3 * i;
    e.time = s->time + i;
    e.terminal = s->terminal + i;
    e.task_id = s->task_id + i;
    return e;
}

int oco_validate(const oco_layout *L) {
    const int n = L->height * L->width;
    if (L->height < 1 || L->width < 1 || n > OCO_MAX_CELLS) return -1;
    if (L->n_pot < 1 || L->n_pot > OCO_MAX_SPECIAL || L->n_goal > OCO_MAX_SPECIAL ||
        L->n_onion_pile > OCO_MAX_SPECIAL || L->n_plate_pile > OCO_MAX_SPECIAL) return -2;
    if (L->n_wall > 2 * OCO_MAX_CELLS) return -3;
    return 0;
}

void oco_reset(const oco_layout *L, int64_t N, const uint32_t *keys, oco_state s, uint8_t *obs0, uint8_t *obs1) {
    const size_t ob = (size_t)L->height * L->width * 26;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        env_t e = view(L, &s, i);
        reset_one(L, keys + 2 * i, &e);
        get_obs_one(L, &e, obs0 + ob * i, obs1 + ob * i);
    }
}

void oco_get_obs(const oco_layout *L, int64_t N, oco_state s, uint8_t *obs0, uint8_t *obs1) {
    const size_t ob = (size_t)L->height * L->width * 26;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        env_t e = view(L, &s, i);
        get_obs_one(L, &e, obs0 + ob * i, obs1 + ob * i);
    }
}

/* in-place step of N envs; the four LogWrapper arrays (float[N,2]) may be NULL */
void oco_step(const oco_layout *L, int64_t N, const uint32_t *keys, oco_state s, const int32_t *act0,
              const int32_t *act1, uint8_t *obs0, uint8_t *obs1, float *reward, float *shaped0, float *shaped1,
              uint8_t *done, float *ep_ret, float *ep_len, float *ret_ret, float *ret_len) {
    const size_t ob = (size_t)L->height * L->width * 26;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        env_t e = view(L, &s, i);
        step_one(L, keys + 2 * i, &e, act0[i], act1[i], obs0 + ob * i, obs1 + ob * i, reward + i, shaped0 + i,
                 shaped1 + i, done + i);
        if (ep_ret) log_wrapper_one(reward[i], done[i], ep_ret + 2 * i, ep_len + 2 * i, ret_ret + 2 * i, ret_len + 2 * i);
    }
}

int oco_max_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oco_set_threads(int n) {
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
