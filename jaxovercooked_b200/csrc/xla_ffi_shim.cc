// xla_ffi_shim.cc -- XLA FFI (typed custom call) handlers that forward XLA's device buffers to the C ABI of
// liboc_b200.so (include/oc_b200.h).  This is the binding that lets the reference's own `Overcooked.reset / step /
// get_obs` keep their signatures under jit / vmap / scan (see INTEGRATION.md section 1 and
// jaxovercooked_b200/jax_ffi.py).
//
// STATUS: EXPERIMENTAL, NEVER RUN -- JAX (and with it xla/ffi/api/ffi.h) cannot be installed in this image.  What is checked
// here: the file compiles against a stand-in of the FFI header (tests/c_abi/xla_ffi_stub) that static_asserts every handler's
// parameter list against its binding, position by position (tests/test_abi_and_host.py).  Run-time behaviour under XLA
// (buffer aliasing, command-buffer capture, PRED layout) is unverified.
// __graft_entry__.build() compiles this file against the real header only when `import jax` succeeds:
//   g++ -O2 -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") -Iinclude
//       -I/usr/local/cuda/include jaxovercooked_b200/csrc/xla_ffi_shim.cc -Ljaxovercooked_b200 -loc_b200
//       -o jaxovercooked_b200/liboc_b200_ffi.so
//
// Conventions: the layout handle (an `oc_layout*` created once per env object with oc_layout_create) travels as an
// int64 attribute -- the env is a static jit argument in the reference (baselines/IPPO_CL.py:454), so there is one
// handle per compiled program.  N is the product of the leading (vmapped) dimensions of `keys`; the un-vmapped call
// has N = 1.  The eleven State results are meant to be aliased onto the eleven State operands with
// `input_output_aliases`, which makes oc_step run in place.
#include <cstdint>
#include <string>

#include <cuda_runtime_api.h>

#include "oc_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

template <typename T>
int64_t leading_count(const T &keys) {   // keys: u32[..., 2]
    int64_t n = 1;
    const auto dims = keys.dimensions();
    for (size_t i = 0; i + 1 < dims.size(); ++i) n *= dims[i];
    return n;
}

ffi::Error to_error(int rc, const char *what) {
    if (rc == OC_OK) return ffi::Error::Success();
    return ffi::Error::Internal(std::string(what) + ": " + oc_status_string(rc));
}

#define OC_STATE_ARGS                                                                                             \
    ffi::Buffer<ffi::U32> agent_pos, ffi::Buffer<ffi::S8> agent_dir, ffi::Buffer<ffi::S32> agent_dir_idx,         \
        ffi::Buffer<ffi::S32> agent_inv, ffi::Buffer<ffi::U32> goal_pos, ffi::Buffer<ffi::U32> pot_pos,           \
        ffi::Buffer<ffi::PRED> wall_map, ffi::Buffer<ffi::U8> maze_map, ffi::Buffer<ffi::S32> time,               \
        ffi::Buffer<ffi::PRED> terminal, ffi::Buffer<ffi::S32> task_id
#define OC_STATE_RETS                                                                                             \
    ffi::ResultBuffer<ffi::U32> o_agent_pos, ffi::ResultBuffer<ffi::S8> o_agent_dir,                              \
        ffi::ResultBuffer<ffi::S32> o_agent_dir_idx, ffi::ResultBuffer<ffi::S32> o_agent_inv,                     \
        ffi::ResultBuffer<ffi::U32> o_goal_pos, ffi::ResultBuffer<ffi::U32> o_pot_pos,                            \
        ffi::ResultBuffer<ffi::PRED> o_wall_map, ffi::ResultBuffer<ffi::U8> o_maze_map,                           \
        ffi::ResultBuffer<ffi::S32> o_time, ffi::ResultBuffer<ffi::PRED> o_terminal,                              \
        ffi::ResultBuffer<ffi::S32> o_task_id

#define OC_IN_STATE                                                                                               \
    oc_state {                                                                                                    \
        agent_pos.typed_data(), agent_dir.typed_data(), agent_dir_idx.typed_data(), agent_inv.typed_data(),       \
            goal_pos.typed_data(), pot_pos.typed_data(), reinterpret_cast<uint8_t *>(wall_map.typed_data()),      \
            maze_map.typed_data(), time.typed_data(), reinterpret_cast<uint8_t *>(terminal.typed_data()),         \
            task_id.typed_data()                                                                                  \
    }
#define OC_OUT_STATE                                                                                              \
    oc_state {                                                                                                    \
        o_agent_pos->typed_data(), o_agent_dir->typed_data(), o_agent_dir_idx->typed_data(),                      \
            o_agent_inv->typed_data(), o_goal_pos->typed_data(), o_pot_pos->typed_data(),                         \
            reinterpret_cast<uint8_t *>(o_wall_map->typed_data()), o_maze_map->typed_data(),                      \
            o_time->typed_data(), reinterpret_cast<uint8_t *>(o_terminal->typed_data()), o_task_id->typed_data()  \
    }

// env.step under vmap  (jax_marl/environments/multi_agent_env.py:41-69)
ffi::Error StepImpl(cudaStream_t stream, int64_t layout, ffi::Buffer<ffi::U32> keys, OC_STATE_ARGS,
                    ffi::Buffer<ffi::S32> act0, ffi::Buffer<ffi::S32> act1, OC_STATE_RETS,
                    ffi::ResultBuffer<ffi::U8> obs0, ffi::ResultBuffer<ffi::U8> obs1,
                    ffi::ResultBuffer<ffi::F32> reward, ffi::ResultBuffer<ffi::F32> shaped0,
                    ffi::ResultBuffer<ffi::F32> shaped1, ffi::ResultBuffer<ffi::PRED> done) {
    const int64_t n = leading_count(keys);
    const int rc = oc_step(reinterpret_cast<const oc_layout *>(layout), n, keys.typed_data(), OC_IN_STATE,
                           act0.typed_data(), act1.typed_data(), /*reset_state=*/nullptr, OC_OUT_STATE,
                           obs0->typed_data(), obs1->typed_data(), reward->typed_data(), shaped0->typed_data(),
                           shaped1->typed_data(), reinterpret_cast<uint8_t *>(done->typed_data()),
                           /*extras=*/nullptr, stream);
    return to_error(rc, "oc_step");
}

// env.step(key, state, actions, reset_state): the caller-supplied state an environment restarts from
// (multi_agent_env.py:47,55-59) instead of reset(key_reset).  The second State travels as eleven more operands.
#define OC_RS_ARGS                                                                                                \
    ffi::Buffer<ffi::U32> r_agent_pos, ffi::Buffer<ffi::S8> r_agent_dir, ffi::Buffer<ffi::S32> r_agent_dir_idx,   \
        ffi::Buffer<ffi::S32> r_agent_inv, ffi::Buffer<ffi::U32> r_goal_pos, ffi::Buffer<ffi::U32> r_pot_pos,     \
        ffi::Buffer<ffi::PRED> r_wall_map, ffi::Buffer<ffi::U8> r_maze_map, ffi::Buffer<ffi::S32> r_time,         \
        ffi::Buffer<ffi::PRED> r_terminal, ffi::Buffer<ffi::S32> r_task_id
ffi::Error StepResetStateImpl(cudaStream_t stream, int64_t layout, ffi::Buffer<ffi::U32> keys, OC_STATE_ARGS,
                              ffi::Buffer<ffi::S32> act0, ffi::Buffer<ffi::S32> act1, OC_RS_ARGS, OC_STATE_RETS,
                              ffi::ResultBuffer<ffi::U8> obs0, ffi::ResultBuffer<ffi::U8> obs1,
                              ffi::ResultBuffer<ffi::F32> reward, ffi::ResultBuffer<ffi::F32> shaped0,
                              ffi::ResultBuffer<ffi::F32> shaped1, ffi::ResultBuffer<ffi::PRED> done) {
    const int64_t n = leading_count(keys);
    const oc_state rs = {r_agent_pos.typed_data(), r_agent_dir.typed_data(), r_agent_dir_idx.typed_data(),
                         r_agent_inv.typed_data(), r_goal_pos.typed_data(), r_pot_pos.typed_data(),
                         reinterpret_cast<uint8_t *>(r_wall_map.typed_data()), r_maze_map.typed_data(), r_time.typed_data(),
                         reinterpret_cast<uint8_t *>(r_terminal.typed_data()), r_task_id.typed_data()};
    const int rc = oc_step(reinterpret_cast<const oc_layout *>(layout), n, keys.typed_data(), OC_IN_STATE,
                           act0.typed_data(), act1.typed_data(), &rs, OC_OUT_STATE, obs0->typed_data(), obs1->typed_data(),
                           reward->typed_data(), shaped0->typed_data(), shaped1->typed_data(),
                           reinterpret_cast<uint8_t *>(done->typed_data()), /*extras=*/nullptr, stream);
    return to_error(rc, "oc_step");
}

// T steps of the same batch in one custom call -- what `jax.lax.scan(_env_step, ..., length=T)` over `jax.vmap(env.step)` does
// when the actions are given (baselines/unused/random_actions.py:197-225): keys u32[T,N,2], actions i32[T,N]; observations
// u8[T,N,h,w,26] per agent, reward / shaped f32[T,N], done bool[T,N].  The State results are aliased onto the State operands.
ffi::Error RolloutImpl(cudaStream_t stream, int64_t layout, ffi::Buffer<ffi::U32> keys, OC_STATE_ARGS,
                       ffi::Buffer<ffi::S32> act0, ffi::Buffer<ffi::S32> act1, OC_STATE_RETS,
                       ffi::ResultBuffer<ffi::U8> obs0, ffi::ResultBuffer<ffi::U8> obs1,
                       ffi::ResultBuffer<ffi::F32> reward, ffi::ResultBuffer<ffi::F32> shaped0,
                       ffi::ResultBuffer<ffi::F32> shaped1, ffi::ResultBuffer<ffi::PRED> done) {
    const auto kd = keys.dimensions();                          // [T, N, 2]
    if (kd.size() != 3) return ffi::Error::Internal("oc_rollout: keys must be u32[T, N, 2]");
    const int64_t T = kd[0], n = kd[1];
    const oc_state in = OC_IN_STATE, out = OC_OUT_STATE;
    if (in.maze_map != out.maze_map)      // oc_rollout steps in place: the caller must alias the State (input_output_aliases)
        return ffi::Error::Internal("oc_rollout: the State results must alias the State operands");
    oc_rollout_io io = {};
    io.act0 = act0.typed_data(); io.act1 = act1.typed_data();
    io.obs0 = obs0->typed_data(); io.obs1 = obs1->typed_data();
    io.obs_step_stride = (int64_t)(obs0->element_count() / (size_t)(T > 0 ? T : 1));
    io.reward = reward->typed_data(); io.shaped0 = shaped0->typed_data(); io.shaped1 = shaped1->typed_data();
    io.done = reinterpret_cast<uint8_t *>(done->typed_data());
    const int rc = oc_rollout(reinterpret_cast<const oc_layout *>(layout), n, (int32_t)T, keys.typed_data(), out, &io,
                              /*extras=*/nullptr, stream);
    return to_error(rc, "oc_rollout");
}

// the same step that additionally returns the compact observations uint8[N, cell_stride] (oc_step_extras.cells), the input
// of OcPolicyActCells
ffi::Error StepCellsImpl(cudaStream_t stream, int64_t layout, ffi::Buffer<ffi::U32> keys, OC_STATE_ARGS,
                         ffi::Buffer<ffi::S32> act0, ffi::Buffer<ffi::S32> act1, OC_STATE_RETS,
                         ffi::ResultBuffer<ffi::U8> obs0, ffi::ResultBuffer<ffi::U8> obs1,
                         ffi::ResultBuffer<ffi::F32> reward, ffi::ResultBuffer<ffi::F32> shaped0,
                         ffi::ResultBuffer<ffi::F32> shaped1, ffi::ResultBuffer<ffi::PRED> done,
                         ffi::ResultBuffer<ffi::U8> cells) {
    const int64_t n = leading_count(keys);
    oc_step_extras ex = {};
    ex.cells = cells->typed_data();
    const int rc = oc_step(reinterpret_cast<const oc_layout *>(layout), n, keys.typed_data(), OC_IN_STATE,
                           act0.typed_data(), act1.typed_data(), /*reset_state=*/nullptr, OC_OUT_STATE,
                           obs0->typed_data(), obs1->typed_data(), reward->typed_data(), shaped0->typed_data(),
                           shaped1->typed_data(), reinterpret_cast<uint8_t *>(done->typed_data()), &ex, stream);
    return to_error(rc, "oc_step");
}

// env.reset under vmap  (overcooked_environment/overcooked.py:136-248)
ffi::Error ResetImpl(cudaStream_t stream, int64_t layout, ffi::Buffer<ffi::U32> keys, OC_STATE_RETS,
                     ffi::ResultBuffer<ffi::U8> obs0, ffi::ResultBuffer<ffi::U8> obs1) {
    const int64_t n = leading_count(keys);
    const int rc = oc_reset(reinterpret_cast<const oc_layout *>(layout), n, keys.typed_data(), OC_OUT_STATE,
                            obs0->typed_data(), obs1->typed_data(), stream);
    return to_error(rc, "oc_reset");
}

// env.get_obs under vmap  (overcooked_environment/overcooked.py:250-364)
ffi::Error GetObsImpl(cudaStream_t stream, int64_t layout, OC_STATE_ARGS, ffi::ResultBuffer<ffi::U8> obs0,
                      ffi::ResultBuffer<ffi::U8> obs1) {
    const int64_t n = time.element_count();
    const int rc = oc_get_obs(reinterpret_cast<const oc_layout *>(layout), n, OC_IN_STATE, obs0->typed_data(),
                              obs1->typed_data(), stream);
    return to_error(rc, "oc_get_obs");
}

// pi, value = network.apply(params, obs_batch); action = pi.sample(seed=key); log_prob = pi.log_prob(action)
// (baselines/IPPO_MLP.py:435-439) from the compact observations an OcStep call produced; `policy` is an oc_policy* created
// with oc_policy_create (layout given) and refreshed with oc_policy_update after each optimiser step.
ffi::Error PolicyActCellsImpl(cudaStream_t stream, int64_t policy, ffi::Buffer<ffi::U8> cells, ffi::Buffer<ffi::U32> key,
                              ffi::ResultBuffer<ffi::F32> logits, ffi::ResultBuffer<ffi::F32> value,
                              ffi::ResultBuffer<ffi::S32> action, ffi::ResultBuffer<ffi::F32> log_prob) {
    const auto dims = cells.dimensions();                       // [N, cell_stride]
    const int64_t n = dims.size() ? dims[0] : 0;
    const int rc = oc_policy_forward_cells(reinterpret_cast<const oc_policy *>(policy), n, cells.typed_data(),
                                           logits->typed_data(), value->typed_data(), key.typed_data(),
                                           action->typed_data(), log_prob->typed_data(), stream);
    return to_error(rc, "oc_policy_forward_cells");
}

#define OC_BIND_STATE_ARGS()                                                                                      \
    .Arg<ffi::Buffer<ffi::U32>>().Arg<ffi::Buffer<ffi::S8>>().Arg<ffi::Buffer<ffi::S32>>()                        \
        .Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::U32>>().Arg<ffi::Buffer<ffi::U32>>()                   \
        .Arg<ffi::Buffer<ffi::PRED>>().Arg<ffi::Buffer<ffi::U8>>().Arg<ffi::Buffer<ffi::S32>>()                   \
        .Arg<ffi::Buffer<ffi::PRED>>().Arg<ffi::Buffer<ffi::S32>>()
#define OC_BIND_STATE_RETS()                                                                                      \
    .Ret<ffi::Buffer<ffi::U32>>().Ret<ffi::Buffer<ffi::S8>>().Ret<ffi::Buffer<ffi::S32>>()                        \
        .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::U32>>().Ret<ffi::Buffer<ffi::U32>>()                   \
        .Ret<ffi::Buffer<ffi::PRED>>().Ret<ffi::Buffer<ffi::U8>>().Ret<ffi::Buffer<ffi::S32>>()                   \
        .Ret<ffi::Buffer<ffi::PRED>>().Ret<ffi::Buffer<ffi::S32>>()

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(OcStep, StepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("layout")
                                  .Arg<ffi::Buffer<ffi::U32>>()   // keys
                                  OC_BIND_STATE_ARGS()
                                  .Arg<ffi::Buffer<ffi::S32>>()   // agent_0 actions
                                  .Arg<ffi::Buffer<ffi::S32>>()   // agent_1 actions
                                  OC_BIND_STATE_RETS()
                                  .Ret<ffi::Buffer<ffi::U8>>()    // obs agent_0
                                  .Ret<ffi::Buffer<ffi::U8>>()    // obs agent_1
                                  .Ret<ffi::Buffer<ffi::F32>>()   // reward
                                  .Ret<ffi::Buffer<ffi::F32>>()   // shaped agent_0
                                  .Ret<ffi::Buffer<ffi::F32>>()   // shaped agent_1
                                  .Ret<ffi::Buffer<ffi::PRED>>()  // done
);

#define OC_BIND_STEP_RETS()                                                                                       \
    .Ret<ffi::Buffer<ffi::U8>>().Ret<ffi::Buffer<ffi::U8>>().Ret<ffi::Buffer<ffi::F32>>()                         \
        .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::PRED>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(OcStepResetState, StepResetStateImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("layout")
                                  .Arg<ffi::Buffer<ffi::U32>>()   // keys
                                  OC_BIND_STATE_ARGS()
                                  .Arg<ffi::Buffer<ffi::S32>>()   // agent_0 actions
                                  .Arg<ffi::Buffer<ffi::S32>>()   // agent_1 actions
                                  OC_BIND_STATE_ARGS()            // reset_state
                                  OC_BIND_STATE_RETS()
                                  OC_BIND_STEP_RETS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(OcRollout, RolloutImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("layout")
                                  .Arg<ffi::Buffer<ffi::U32>>()   // keys [T, N, 2]
                                  OC_BIND_STATE_ARGS()
                                  .Arg<ffi::Buffer<ffi::S32>>()   // agent_0 actions [T, N]
                                  .Arg<ffi::Buffer<ffi::S32>>()   // agent_1 actions [T, N]
                                  OC_BIND_STATE_RETS()
                                  OC_BIND_STEP_RETS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(OcStepCells, StepCellsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("layout")
                                  .Arg<ffi::Buffer<ffi::U32>>()   // keys
                                  OC_BIND_STATE_ARGS()
                                  .Arg<ffi::Buffer<ffi::S32>>()   // agent_0 actions
                                  .Arg<ffi::Buffer<ffi::S32>>()   // agent_1 actions
                                  OC_BIND_STATE_RETS()
                                  .Ret<ffi::Buffer<ffi::U8>>()    // obs agent_0
                                  .Ret<ffi::Buffer<ffi::U8>>()    // obs agent_1
                                  .Ret<ffi::Buffer<ffi::F32>>()   // reward
                                  .Ret<ffi::Buffer<ffi::F32>>()   // shaped agent_0
                                  .Ret<ffi::Buffer<ffi::F32>>()   // shaped agent_1
                                  .Ret<ffi::Buffer<ffi::PRED>>()  // done
                                  .Ret<ffi::Buffer<ffi::U8>>()    // compact observations
);

XLA_FFI_DEFINE_HANDLER_SYMBOL(OcReset, ResetImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("layout")
                                  .Arg<ffi::Buffer<ffi::U32>>()   // keys
                                  OC_BIND_STATE_RETS()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(OcGetObs, GetObsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("layout")
                                  OC_BIND_STATE_ARGS()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Ret<ffi::Buffer<ffi::U8>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(OcPolicyActCells, PolicyActCellsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("policy")
                                  .Arg<ffi::Buffer<ffi::U8>>()    // cells [N, cell_stride]
                                  .Arg<ffi::Buffer<ffi::U32>>()   // key [2]
                                  .Ret<ffi::Buffer<ffi::F32>>()   // logits [2N, A]
                                  .Ret<ffi::Buffer<ffi::F32>>()   // value [2N]
                                  .Ret<ffi::Buffer<ffi::S32>>()   // action [2N]
                                  .Ret<ffi::Buffer<ffi::F32>>()); // log_prob [2N]
