/*
 * oc_b200.h -- C ABI of the B200-native batched Overcooked environment step.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference (LukavdBoogaard/JAXOvercooked,
 * 100 % Python on JAX) has no FFI of its own; its de-facto interface is the object protocol
 *
 *     obs, state                        = env.reset(key)                  overcooked.py:136
 *     obs, state, rewards, dones, infos = env.step(key, state, actions)   multi_agent_env.py:41-69
 *     obs                               = env.get_obs(state)              overcooked.py:250
 *
 * called under jax.vmap over N environments (baselines/IPPO_CL.py:478, :538).  Each entry point below is
 * what an XLA FFI custom call (or the ctypes host in jaxovercooked_b200/) binds for one of those calls,
 * batched over the leading N; see INTEGRATION.md for the binding stubs.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer owned by the caller, valid until `stream` reaches the call, and
 *     naturally aligned for its element type (agent_pos: 8 bytes -- it is read as (x,y) pairs); observation
 *     buffers that are 16-byte aligned leave the SM through the bulk-copy engine, others through a slower
 *     plain-store path; nothing here allocates per call, synchronises, throws or exits;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - return value: OC_OK or a negative oc_status; oc_status_string() names it;
 *   - an oc_layout is an immutable handle (host struct + a small device-side descriptor on the CUDA device
 *     that was current at creation); calls are re-entrant and share no mutable global state;
 *   - arrays are struct-of-arrays with a leading N; dtypes/shapes per env are exactly the reference's
 *     `State` fields (overcooked.py:44-56), so a State pytree maps onto oc_state field by field.
 */
#ifndef OC_B200_H
#define OC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OC_B200_VERSION 120          /* 0.1.2: + oc_rollout, oc_rollout_host */
#define OC_MAX_DIM 16                /* height, width <= 16 (reference registry max: 10 x 11) */
#define OC_MAX_CELLS 256
#define OC_MAX_SPECIAL 16            /* pots, goals, onion piles, plate piles: at most 16 each */
#define OC_PAD 4                     /* agent_view_size - 1 (overcooked.py:88, common.py:153-154) */
#define OC_OBS_CHANNELS 26           /* overcooked.py:86 */
#define OC_NUM_CELL_CODES 68         /* record codes: 0..10 objects, 11..34 pot with status 0..23, 35..66 agent (which, dir, held) */

typedef enum {
    OC_OK = 0,
    OC_ERR_INVALID_ARG = -1,         /* NULL pointer, N < 0, bad enum */
    OC_ERR_UNSUPPORTED_LAYOUT = -2,  /* see oc_layout_create */
    OC_ERR_CUDA = -3,                /* a CUDA runtime call failed; cudaGetLastError has the detail */
    OC_ERR_ALIGNMENT = -4,           /* a pointer is not aligned for its element type */
    OC_ERR_NO_DEVICE = -5            /* no usable sm_100 device: there is NO CPU fallback */
} oc_status;

typedef enum {
    OC_PRNG_THREEFRY_LEGACY = 0,         /* jax_threefry_partitionable=False (default for jax < 0.5.0) */
    OC_PRNG_THREEFRY_PARTITIONABLE = 1   /* jax_threefry_partitionable=True  (default for jax >= 0.5.0) */
} oc_prng_mode;

/* Host-side description of an environment: the eight layout fields the reference env reads
 * (layouts.py:4-58, 79-127) plus the constructor arguments of `Overcooked` (overcooked.py:71-78).
 * Index arrays are HOST pointers to flat cell indices (row * width + col), copied at creation. */
typedef struct {
    int32_t height, width;
    int32_t n_wall, n_goal, n_pot, n_onion_pile, n_plate_pile;
    const int32_t *wall_idx;
    const int32_t *agent_idx;        /* exactly 2 */
    const int32_t *goal_idx;
    const int32_t *pot_idx;
    const int32_t *onion_pile_idx;
    const int32_t *plate_pile_idx;
    int32_t random_reset;            /* overcooked.py:75  */
    int32_t max_steps;               /* overcooked.py:76  */
    int32_t task_id;                 /* overcooked.py:77  */
    int32_t prng_mode;               /* oc_prng_mode */
} oc_layout_desc;

typedef struct oc_layout oc_layout;  /* opaque */

/* Batched `State` (overcooked.py:44-56), leading N on every array. */
typedef struct {
    uint32_t *agent_pos;     /* [N,2,2] (x, y) per agent        */
    int8_t   *agent_dir;     /* [N,2,2] DIR_TO_VEC[agent_dir_idx] */
    int32_t  *agent_dir_idx; /* [N,2]   */
    int32_t  *agent_inv;     /* [N,2]   */
    uint32_t *goal_pos;      /* [N,G,2] static per layout       */
    uint32_t *pot_pos;       /* [N,P,2] static per layout       */
    uint8_t  *wall_map;      /* [N,h,w] bool, static per layout */
    uint8_t  *maze_map;      /* [N,h+8,w+8,3]                   */
    int32_t  *time;          /* [N]     */
    uint8_t  *terminal;      /* [N] bool */
    int32_t  *task_id;       /* [N]     */
} oc_state;

/* Optional extras of oc_step: the LogWrapper bookkeeping (wrappers/baselines.py:75-107) fused into the
 * step, and exact integer rollout statistics for the multi-GPU reduction.  Any pointer may be NULL. */
typedef struct {
    float   *episode_returns;           /* [N,2] in/out */
    float   *episode_lengths;           /* [N,2] in/out */
    float   *returned_episode_returns;  /* [N,2] in/out */
    float   *returned_episode_lengths;  /* [N,2] in/out */
    uint8_t *returned_episode;          /* [N,2] out (bool) */
    /* int64[8] accumulators, atomically added to: [0] finished episodes, [1] sum of sparse reward,
     * [2] sum shaped agent_0, [3] sum shaped agent_1, [4] sum returned_episode_returns[:,0] over finished
     * episodes, [5] sum returned_episode_lengths[:,0] over finished episodes, [6] env-steps, [7] reserved.
     * All addends are integers, so the totals are order-independent and bit-reproducible. */
    long long *stats;
    /* Fused key derivation: when oc_step is called with keys == NULL, env i uses
     * jax.random.split(*split_key, split_num)[split_first + i] (split_key: DEVICE u32[2]), evaluated in the
     * kernel and only for environments that reset -- the callers' `split(_rng, num_envs)`
     * (baselines/IPPO_CL.py:535) without materialising or reading the [N,2] key array. */
    const uint32_t *split_key;
    int64_t split_num;
    int64_t split_first;
    /* Compact form of the observations this step writes: uint8[N, cell_stride] (oc_layout_info), or NULL.  Per env:
     *   [0, n_scan)            for each "scan cell" (pots and the counters an agent can face, oc_layout_cell_tables):
     *                          the code of the record the cell shows, or 0xFF when it shows the layout's static image
     *   [cell_agents + 2a, +2) agent a: its cell index (y*w + x), then its record code as agent_0 sees it
     *   [cell_agents + 4]      1 when the urgency plane (channel 25) is set, else 0
     * Every observation is the static image with those records written over it (agent_1's view: the two agent codes
     * with bit 4 of (code - 35) flipped), plus channel 25 everywhere when urgent -- oc_policy_forward_cells consumes
     * this instead of re-reading the 26-channel images.  The buffer must be 16-byte aligned (OC_ERR_ALIGNMENT); padding
     * bytes of a row are written as 0xFF. */
    uint8_t *cells;
    /* Observation layout.  0: obs0 / obs1 are two uint8[N,h,w,26] arrays (default).  1 (OC_OBS_INTERLEAVED): obs0 points at
     * ONE uint8[N,2,h,w,26] array -- agent_0's and agent_1's view of an env side by side, which flattened per env is the
     * CTRolloutManager's global state obs["__all__"] (wrappers/baselines.py:196: concatenate([obs[a].flatten() for a in
     * agents])) while [:,0] / [:,1] are the per-agent observations; obs1 must be NULL or obs0 + h*w*26.  The array must be
     * 16-byte aligned for the bulk-copy path; not combinable with `cells`. */
    int32_t obs_interleaved;
} oc_step_extras;
#define OC_OBS_INTERLEAVED 1

typedef struct {
    int32_t height, width, n_goal, n_pot;
    int64_t map_bytes_per_env;       /* (h+8)*(w+8)*3 */
    int64_t obs_bytes_per_agent;     /* h*w*26        */
    int64_t algorithmic_bytes;       /* 58*h*w + 117 per env-step */
    int32_t envs_per_warp, warps_per_cta, smem_bytes_per_cta;   /* the layout's default tile shape (large one-step launches); oc_rollout
                                                                    and small batches pick 2 / 4 / 8 / 16 envs per warp per launch */
    int32_t n_scan, cell_stride, cell_agents;   /* layout of oc_step_extras.cells */
} oc_layout_info;

/* Build a layout handle on the current CUDA device.
 * OC_ERR_UNSUPPORTED_LAYOUT unless: 1 <= h,w <= OC_MAX_DIM; exactly two agent cells, distinct, not walls;
 * 1 <= n_pot <= OC_MAX_SPECIAL (the reference itself needs >= 1 pot, common.py:128-131), other lists
 * <= OC_MAX_SPECIAL; every goal / pile / pot cell is listed in wall_idx (layouts.py:116-118 guarantees it for
 * grid layouts); every border cell of the grid is a wall, so an agent never faces outside the grid
 * (true for all 38 registered layouts and for anything produced by pad_observation_space); max_steps >= 1. */
int  oc_layout_create(const oc_layout_desc *desc, oc_layout **out);
void oc_layout_destroy(oc_layout *layout);
int  oc_layout_query(const oc_layout *layout, oc_layout_info *info);

/* env.reset under vmap: overcooked.py:136-248.  keys u32[N,2].  Writes every field of `out` (the whole
 * padded maze_map included) and both observations u8[N,h,w,26]. */
int oc_reset(const oc_layout *layout, int64_t N, const uint32_t *keys, oc_state out,
             uint8_t *obs0, uint8_t *obs1, void *stream);

/* env.step under vmap: multi_agent_env.py:41-69 -> overcooked.py:106-134 (step_env), :366-514 (step_agents),
 * :516-642 (process_interact), :250-364 (get_obs), with auto-reset from `split(key)[1]`.
 *   keys          u32[N,2] per-env keys, or NULL with extras->split_key set (fused split, see oc_step_extras)
 *   act0/act1     i32[N], values 0..5 (overcooked.py:93-100)
 *   reset_state   NULL, or the caller-supplied `reset_state` of multi_agent_env.py:47,55-59
 *   out           may alias `in` field by field (the fast path: XLA input_output_aliases / donated scan
 *                 carries).  Where a field of `out` differs from `in`, it is first copied device-to-device.
 *   reward f32[N] (same value for both agents, overcooked.py:124), shaped0/shaped1 f32[N]
 *   (info["shaped_reward"]), done u8[N] (dones["__all__"] == dones[agent_i]).
 * State invariants of SURVEY.md App. A.8 are assumed on input (every state produced by oc_reset / oc_step
 * satisfies them); the static fields goal_pos / pot_pos / wall_map are taken from the layout. */
int oc_step(const oc_layout *layout, int64_t N, const uint32_t *keys, oc_state in,
            const int32_t *act0, const int32_t *act1, const oc_state *reset_state, oc_state out,
            uint8_t *obs0, uint8_t *obs1, float *reward, float *shaped0, float *shaped1, uint8_t *done,
            const oc_step_extras *extras, void *stream);

/* oc_step for a caller whose policy lives on the HOST: actions come from host memory and the step's results go
 * back to host memory, all enqueued on `stream` (H2D copy, step kernel, D2H copy; nothing synchronises).
 *   h_actions  HOST  i32[2,N]  (row 0 = agent_0, row 1 = agent_1); pinned memory makes the copies asynchronous
 *   h_results  HOST  13*N bytes: reward f32[N] | shaped0 f32[N] | shaped1 f32[N] | done u8[N]
 *   d_scratch  DEVICE 21*N bytes, 16-byte aligned (device copies of the two blocks above); N % 4 == 0
 *   state      stepped in place; observations stay on the device (obs0/obs1), where the learner consumes them.
 * Different env batches on different streams overlap their copies with each other's kernels. */
int oc_step_host(const oc_layout *layout, int64_t N, const uint32_t *keys, oc_state state,
                 const int32_t *h_actions, void *h_results, void *d_scratch, uint8_t *obs0, uint8_t *obs1,
                 const oc_step_extras *extras, void *stream);

/* oc_step_host with every argument but the step's subkey and actions bound once (a closure over one env batch
 * and one stream), so a host-side rollout loop pays one three-argument call per step:
 *   create: `extras` must carry split_num / split_first (and optionally the LogWrapper / stats pointers);
 *   step:   env i uses jax.random.split(*split_key, split_num)[split_first + i]; h_actions HOST i32[2,N];
 *   wait:   blocks until the batch's stream has delivered the last step's results to h_results. */
typedef struct oc_host_stepper oc_host_stepper;
int  oc_host_stepper_create(const oc_layout *layout, int64_t N, oc_state state, void *h_results, void *d_scratch,
                            uint8_t *obs0, uint8_t *obs1, const oc_step_extras *extras, void *stream,
                            oc_host_stepper **out);
int  oc_host_stepper_step(oc_host_stepper *stepper, const uint32_t *split_key, const int32_t *h_actions);
int  oc_host_stepper_wait(oc_host_stepper *stepper);
void oc_host_stepper_destroy(oc_host_stepper *stepper);

/* T consecutive env.step calls on the same batch in ONE launch: the callers' `jax.lax.scan(_env_step, ..., length=T)` over
 * `jax.vmap(env.step)` (baselines/unused/random_actions.py:197-225; baselines/IPPO_CL.py:520-560 with the actions given).
 * Each warp loads its environments once, keeps them in shared memory and registers for the T steps and writes them back
 * once, so a step costs no state traffic and no launch: results are bit-identical to T oc_step calls (tests/test_gpu_rollout.py).
 * Per-step arrays carry a leading T:
 *   keys            u32[T,N,2], or NULL with extras->split_key = DEVICE u32[T,2]: env i at step t uses
 *                   jax.random.split(split_key[t], split_num)[split_first + i]
 *   io->act0/act1   i32[T,N]            (or io->actions8: u8[T,N] = agent_0's action | agent_1's action << 4)
 *   io->obs0/obs1   step t's observations u8[N,h,w,26] start obs_step_stride BYTES after step t-1's
 *                   (>= N*h*w*26; e.g. 2*N*h*w*26 with obs1 = obs0 + N*h*w*26 writes the learners' [T,2,N,h,w,26] buffer)
 *   io->reward/shaped0/shaped1 f32[T,N], io->done u8[T,N]
 *                   (or io->results8: u8[T,N] = reward/20 | class(shaped0) << 2 | class(shaped1) << 4 | done << 6 with
 *                    class(0) = 0, class(3) = 1, class(5) = 2 -- the only values the reference produces, overcooked.py:30-35,502)
 *   extras          LogWrapper trackers [N,2] are read once, updated T times and written once; returned_episode is
 *                   u8[T,N,2]; cells is u8[T,N,cell_stride]; stats accumulate the T steps.
 * `state` is stepped in place.  reset_state is not supported here (use oc_step). */
typedef struct {
    const int32_t *act0, *act1;
    const uint8_t *actions8;
    uint8_t *obs0, *obs1;
    int64_t obs_step_stride;
    float *reward, *shaped0, *shaped1;
    uint8_t *done;
    uint8_t *results8;
    int32_t obs_interleaved;   /* as oc_step_extras.obs_interleaved; step t's uint8[N,2,h,w,26] starts at obs0 + t*obs_step_stride */
    /* Optional DEVICE scratch u8[T,N] for int32 actions: when given (and the batch is large enough for it to pay), a small
     * pre-pass packs act0 / act1 into it and the rollout reads one byte per env and step instead of eight -- the int32 rows
     * are DRAM reads scattered through the observation write stream, and 8x fewer of them is worth ~8 % at 65,536 envs. */
    uint8_t *actions8_scratch;
} oc_rollout_io;

int oc_rollout(const oc_layout *layout, int64_t N, int32_t T, const uint32_t *keys, oc_state state,
               const oc_rollout_io *io, const oc_step_extras *extras, void *stream);

/* oc_rollout for a caller whose policy lives on the HOST -- the T-step form of oc_step_host: ONE host-to-device copy of
 * the T steps' actions, one launch, ONE device-to-host copy of the T steps' results, all enqueued on `stream`.
 *   wire = OC_WIRE_REFERENCE  h_actions HOST i32[2,T,N] (agent-major); h_results HOST 13*T*N bytes:
 *                             reward f32[T,N] | shaped0 f32[T,N] | shaped1 f32[T,N] | done u8[T,N]; d_scratch 21*T*N bytes
 *   wire = OC_WIRE_PACKED     h_actions HOST u8[T,N] (actions8); h_results HOST u8[T,N] (results8); d_scratch 2*T*N bytes
 * d_scratch is DEVICE memory, 16-byte aligned; T*N % 4 == 0.  extras->split_key (DEVICE u32[T,2]) is required. */
typedef enum { OC_WIRE_REFERENCE = 0, OC_WIRE_PACKED = 1 } oc_wire_format;
int oc_rollout_host(const oc_layout *layout, int64_t N, int32_t T, oc_state state, int wire, const void *h_actions,
                    void *h_results, void *d_scratch, uint8_t *obs0, uint8_t *obs1, int64_t obs_step_stride,
                    const oc_step_extras *extras, void *stream);

/* env.get_obs under vmap: overcooked.py:250-364. */
int oc_get_obs(const oc_layout *layout, int64_t N, oc_state in, uint8_t *obs0, uint8_t *obs1, void *stream);
/* ... and their compact form as well (see oc_step_extras.cells): what a rollout needs right after oc_reset. */
int oc_get_obs_cells(const oc_layout *layout, int64_t N, oc_state in, uint8_t *obs0, uint8_t *obs1, uint8_t *cells,
                     void *stream);
/* ... as ONE uint8[N,2,h,w,26] array (oc_step_extras.obs_interleaved): what CTRolloutManager.batch_reset returns as
 * obs["__all__"] (wrappers/baselines.py:217) with the per-agent observations as its [:,0] / [:,1] slices. */
int oc_get_obs_all(const oc_layout *layout, int64_t N, oc_state in, uint8_t *obs_all, void *stream);
/* Decoding tables of the compact form, into HOST memory (either may be NULL): the 26-byte record of every cell code
 * (uint8[OC_NUM_CELL_CODES][26], as agent_0 sees it) and the cell index of every scan cell (int32[n_scan]). */
int oc_layout_cell_tables(const oc_layout *layout, uint8_t *records, int32_t *scan_cells);

/* jax.random.split(key, num)[first : first+count] on the device (callers: IPPO_CL.py:477,535;
 * wrappers/baselines.py:202,207).  `key` is a DEVICE pointer to u32[2]; out u32[count,2] must not alias it.
 * Row i depends only on (key, i, num), so a GPU holding envs [first, first+count) of a global batch of
 * `num` derives exactly the keys the single-device program would. */
int oc_prng_split(int prng_mode, const uint32_t *key, int64_t num, int64_t first, int64_t count,
                  uint32_t *out, void *stream);

/* n sequential `key, sub = jax.random.split(key)` on the device (the callers' per-step
 * `rng, _rng = jax.random.split(rng)`, IPPO_CL.py:533): subkeys u32[n,2] receives the n `sub`s and *key (DEVICE
 * u32[2]) is advanced in place. */
int oc_prng_chain(int prng_mode, uint32_t *key, int n, uint32_t *subkeys, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Actor-critic policy forward ("next" row f3 of the scope table): the consumer of the observations on the rollout
 * path.  Replaces, for the flattened-observation MLP of architectures/mlp.py:11-59,
 *     pi, value = network.apply(train_state.params, obs_batch)      baselines/IPPO_MLP.py:435
 *     action    = pi.sample(seed=_rng)                               baselines/IPPO_MLP.py:438
 *     log_prob  = pi.log_prob(action)                                baselines/IPPO_MLP.py:439
 * with obs_batch = batchify(last_obs, env.agents, num_actors) (baselines/utils.py:48-61): uint8 rows, agent-major.
 * FLOATING POINT: layer 1 is accumulated in fp32 from fp32 weights, layer 2 runs on the tensor cores with fp16
 * operands and fp32 accumulation, tanh is the hardware approximation -- results are within a tolerance of the
 * fp32 reference (tests/test_gpu_policy.py states it), never bit-exact. */

#define OC_POLICY_HIDDEN 128         /* architectures/mlp.py:27,35,49,51 */
#define OC_POLICY_MAX_ACTIONS 8

typedef enum { OC_ACT_TANH = 0, OC_ACT_RELU = 1 } oc_activation;   /* architectures/mlp.py:20-23 */

/* The six Dense layers in flax's creation order (Dense_0..2 actor, Dense_3..5 critic); kernels are [in, out]
 * row-major float32, biases [out]; DEVICE pointers, read by oc_policy_create / oc_policy_update only. */
typedef struct {
    int32_t obs_dim;                 /* K = h * w * 26 (any even K >= 16 works) */
    int32_t action_dim;              /* 1..OC_POLICY_MAX_ACTIONS (6 for Overcooked) */
    int32_t activation;              /* oc_activation */
    int32_t prng_mode;               /* oc_prng_mode, for the categorical draw */
    const float *kernel[6];
    const float *bias[6];
    const uint8_t *obs_template;     /* HOST pointer, uint8[obs_dim], or NULL: an observation most rows are close to
                                        (oc_layout_obs_template); layer 1 is evaluated as a correction to it.  Any
                                        input is handled correctly with any template -- only the speed depends on it */
    const oc_layout *layout;         /* or NULL.  When given, obs_dim must be h*w*26, the layout's static image is the template
                                        and oc_policy_forward_cells becomes available.  Everything the policy needs from the layout
                                        is copied by oc_policy_create (the layout may be destroyed first); it must live on
                                        the current device (OC_ERR_INVALID_ARG otherwise) */
} oc_policy_desc;

typedef struct oc_policy oc_policy;  /* opaque: the parameters repacked for the kernels, in device memory */

int oc_policy_create(const oc_policy_desc *desc, oc_policy **out);
/* Re-read the parameters after an optimiser step (enqueue-only, same pointers or new ones). */
int oc_policy_update(oc_policy *p, const oc_policy_desc *desc, void *stream);
void oc_policy_destroy(oc_policy *p);

/* Sharded draws.  After oc_policy_set_shard(p, first, global_count) the categorical draws of the forward calls use the noise
 * the single-device program would use: for oc_policy_forward_cells, env i of a call is env first + i of global_count
 * environments (rows first + i and global_count + first + i of the 2 * global_count actor rows); for oc_policy_forward,
 * row m is row first + m of global_count rows.  A GPU that owns one shard of the env axis then samples exactly the
 * actions a single device would sample for those environments.  global_count = 0 switches it off (the default). */
int oc_policy_set_shard(oc_policy *p, int64_t first, int64_t global_count);

/* logits[M, action_dim], value[M] for obs[M, obs_dim] (rows contiguous).  If `key` (DEVICE uint32[2]) is given,
 * action[m] = argmax_a(logits[m, a] + gumbel(key, (1, M, A))[0, m, a]) and log_prob[m] = log_softmax(logits[m])[action[m]]
 * are written too.  logits / value / action / log_prob may each be NULL. */
int oc_policy_forward(const oc_policy *p, int64_t M, const uint8_t *obs, float *logits, float *value,
                      const uint32_t *key, int32_t *action, float *log_prob, void *stream);

/* The same for BOTH agents of N environments, read from the compact observations an oc_step / oc_get_obs_cells call
 * produced (oc_step_extras.cells) instead of the 26-channel images: row m < N is agent_0 of env m, row N + m agent_1
 * (batchify order), so logits are [2N, action_dim].  Layer 1 becomes base + one precomputed fp16 embedding per
 * deviating cell; the observations are not read at all.  OC_ERR_UNSUPPORTED_LAYOUT if the policy was created without
 * a layout or the layout has more than 59 scan cells. */
int oc_policy_forward_cells(const oc_policy *p, int64_t N, const uint8_t *cells, float *logits, float *value,
                            const uint32_t *key, int32_t *action, float *log_prob, void *stream);

/* The draw alone, for callers that keep `pi` around: action / log_prob (either may be NULL) from logits[M, A]. */
int oc_policy_sample(int prng_mode, int64_t M, int32_t A, const float *logits, const uint32_t *key, int32_t *action,
                     float *log_prob, void *stream);

/* The layout's static observation (walls, piles, goals, pot locations: the part of get_obs that never changes),
 * uint8[h * w * 26] into HOST memory: the natural `obs_template` for oc_policy_desc. */
int oc_layout_obs_template(const oc_layout *layout, uint8_t *out);

const char *oc_status_string(int status);
int oc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OC_B200_H */
