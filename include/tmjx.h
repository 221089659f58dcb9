/*
 * tmjx.h — C ABI of the B200-native rodent-tracking environment step.
 *
 * This is the drop-in boundary for the one hot path this repository replaces: the batched
 * `reset` / `step` of track-mjx's motion-tracking environment, i.e. the work behind
 *
 *   SingleClipTracking.reset_from_clip   reference track_mjx/environment/task/single_clip_tracking.py:121-205
 *   SingleClipTracking.step              reference track_mjx/environment/task/single_clip_tracking.py:207-320
 *   MultiClipTracking.reset / _get_reference_clip   reference .../task/multi_clip_tracking.py:74-109
 *   compute_tracking_rewards             reference track_mjx/environment/task/reward.py:359-485
 *   BaseWalker.compute_local_*           reference track_mjx/environment/walker/base.py:170-258
 *   PipelineEnv.pipeline_init/step -> brax.mjx.pipeline.init/step -> mujoco.mjx.forward/step
 *                                        (upstream brax 0.12.3 / mujoco-mjx 3.3.2, called at
 *                                         single_clip_tracking.py:163 and :219)
 *   auto-reset + episode wrapper         reference track_mjx/environment/wrappers.py:104-144, 288-310
 *
 * The reference has no FFI of its own: it is pure Python/JAX and the XLA program is the
 * "operator".  A maintainer binds these entry points either through an XLA-FFI handler
 * (csrc/tmjx_xla_ffi.cc, built only when jaxlib's headers exist) or through ctypes
 * (track-mjx_b200/_lib.py); INTEGRATION.md shows both.
 *
 * Conventions
 *   - plain C types only; all array pointers in TmjxState / TmjxOut / `action` are DEVICE pointers,
 *     row-major `[n_env, dim]` fp32 (int32 for indices) — the layout `jax.vmap` gives the same leaves;
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as void*): no allocation, no
 *     synchronisation, re-entrant; errors are returned as negative codes, text via tmjx_last_error();
 *   - simulation failure is never an error: NaNs raise `done` exactly as the reference does
 *     (single_clip_tracking.py:286-293).
 */
#ifndef TMJX_H_
#define TMJX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMJX_ABI_VERSION 1
#define TMJX_MAX_IDX 80      /* upper bound on the walker index tables */
#define TMJX_N_METRICS 20    /* single_clip_tracking.py:176-197 */

enum { TMJX_OK = 0, TMJX_E_ARG = -1, TMJX_E_BLOB = -2, TMJX_E_CUDA = -3, TMJX_E_UNSUPPORTED = -4 };
enum { TMJX_SOLVER_CG = 0, TMJX_SOLVER_NEWTON = 1 };

/* order of TmjxOut.metrics columns == insertion order of the reference's metrics dict */
enum {
  TMJX_M_POS_REWARD = 0, TMJX_M_QUAT_REWARD, TMJX_M_JOINT_REWARD, TMJX_M_ANGVEL_REWARD, TMJX_M_BODYPOS_REWARD,
  TMJX_M_ENDEFF_REWARD, TMJX_M_CTRL_COST, TMJX_M_CTRL_DIFF_COST, TMJX_M_ENERGY_COST, TMJX_M_DONE, TMJX_M_TOO_FAR,
  TMJX_M_BAD_POSE, TMJX_M_BAD_QUAT, TMJX_M_FALL, TMJX_M_NAN, TMJX_M_JOINT_DISTANCE, TMJX_M_SUMMED_POS_DISTANCE,
  TMJX_M_QUAT_DISTANCE, TMJX_M_VAR_COST, TMJX_M_JERK_COST
};

/* Task constants: the kwargs of MultiClipTracking.__init__ (multi_clip_tracking.py:16-31), RewardConfig
 * (reward.py:15-54) and the walker index tables (walker/rodent.py:89-114). */
typedef struct TmjxTaskConfig {
  int32_t abi_version;                 /* TMJX_ABI_VERSION */
  /* env_args */
  int32_t physics_steps_per_control_step;
  int32_t solver;                      /* TMJX_SOLVER_* */
  int32_t iterations;
  int32_t ls_iterations;
  float mj_model_timestep;
  float mocap_hz;
  int32_t clip_length;
  int32_t traj_length;
  int32_t episode_length;              /* brax EpisodeWrapper limit (train.py:221-225); used only with TMJX_F_AUTORESET */
  /* reward_weights */
  float too_far_dist, bad_pose_dist, bad_quat_dist;
  float ctrl_cost_weight, ctrl_diff_cost_weight, energy_cost_weight;
  float pos_reward_weight, quat_reward_weight, joint_reward_weight;
  float angvel_reward_weight, bodypos_reward_weight, endeff_reward_weight;
  float healthy_z_min, healthy_z_max;
  float pos_reward_exp_scale, quat_reward_exp_scale, joint_reward_exp_scale;
  float angvel_reward_exp_scale, bodypos_reward_exp_scale, endeff_reward_exp_scale;
  float penalty_pos_distance_scale[3];
  int32_t var_window_size;
  float var_coeff, jerk_coeff;
  /* walker index tables: MODEL ids exactly as mj_name2id returns them (the off-by-one between
   * `data.xpos[1:]` and these ids, and the clamp of id 67, are reproduced inside the step) */
  int32_t n_joint_idxs, n_body_idxs, n_endeff_idxs;
  int32_t joint_idxs[TMJX_MAX_IDX];
  int32_t body_idxs[TMJX_MAX_IDX];
  int32_t endeff_idxs[TMJX_MAX_IDX];
  int32_t torso_idx;
  /* named bindings used by _get_proprioception / _get_appendages_pos (single_clip_tracking.py:322-354) */
  int32_t torso_body_id;
  int32_t n_appendages;
  int32_t appendage_body_ids[8];
} TmjxTaskConfig;

/* Per-env state carried between calls (caller-owned device buffers, all `[n_env, dim]`).
 * pipeline_state subset of mjx.Data that is persistent: qpos, qvel, act, time, qacc_warmstart, plus the
 * derived quantities the reference reads one call later (they are "stale by one substep" in the reference
 * too because mjx.step integrates after forward): xpos, xquat, qfrc_actuator. */
typedef struct TmjxState {
  float* qpos;            /* [n, nq]  */
  float* qvel;            /* [n, nv]  */
  float* act;             /* [n, na]  */
  float* time;            /* [n]      */
  float* qacc_warmstart;  /* [n, nv]  */
  float* xpos;            /* [n, nbody, 3]  derived, written by forward/step */
  float* xquat;           /* [n, nbody, 4]  derived, written by forward/step */
  float* qfrc_actuator;   /* [n, nv]        derived, written by forward/step */
  /* info */
  int32_t* clip_idx;      /* [n] */
  int32_t* start_frame;   /* [n] */
  int32_t* buffer_index;  /* [n] */
  float* prev_ctrl;       /* [n, nu] */
  float* action_buffer;   /* [n, var_window_size, nu] */
  /* wrapper state (EpisodeWrapper + AutoResetWrapperTracking); may be NULL when flags do not ask for it */
  float* steps;           /* [n] */
  float* truncation;      /* [n] */
  float* first_qpos;      /* [n, nq] snapshot taken by tmjx_forward with TMJX_F_SNAPSHOT */
  float* first_qvel;      /* [n, nv] */
  float* first_act;       /* [n, na] */
  float* first_time;      /* [n] */
  float* first_qacc_warmstart; /* [n, nv] */
  float* first_xpos;      /* [n, nbody, 3] */
  float* first_xquat;     /* [n, nbody, 4] */
  float* first_qfrc_actuator; /* [n, nv] */
  float* first_obs;       /* [n, obs] */
  float* first_prev_ctrl; /* [n, nu] */
} TmjxState;

typedef struct TmjxOut {
  float* obs;             /* [n, obs_size]  reference_obs ‖ proprioceptive_obs */
  float* reward;          /* [n] */
  float* done;            /* [n] fp32 0/1 like the reference; read as the PREVIOUS done with TMJX_F_AUTORESET */
  float* metrics;         /* [n, TMJX_N_METRICS] */
  int32_t* cur_frame;     /* [n] */
  /* optional debug taps for parity tests (NULL = skip) */
  float* dbg_qacc;        /* [n, nv]   solver output of the last substep */
  float* dbg_qacc_smooth; /* [n, nv] */
  float* dbg_qfrc_bias;   /* [n, nv] */
  float* dbg_qfrc_constraint; /* [n, nv] */
  float* dbg_contact_dist;/* [n, ncon] */
  float* dbg_efc_force;   /* [n, nefc] */
  float* dbg_qM;          /* [n, nv, nv] dense symmetric */
  float* dbg_subtree_com; /* [n, 3] of the walker tree root */
} TmjxOut;

/* step flags */
#define TMJX_F_AUTORESET   1u  /* fuse EpisodeWrapper + AutoResetWrapperTracking (wrappers.py:288-310) */
#define TMJX_F_SNAPSHOT    2u  /* tmjx_forward: also store first_* (wrappers.py:281-286) */
#define TMJX_F_EPILOGUE_ONLY 4u /* tmjx_step test hook: skip the physics substeps; qpos / qvel / time / xpos / xquat / qfrc_actuator of TmjxState
                                   are taken as the POST-physics state and only the task layer (single_clip_tracking.py:221-320: frame lookup,
                                   info updates, rewards, obs, done, metrics) runs on them.  Lets the golden vectors computed by the reference's
                                   own task code (tests/golden/task_layer.npz) check the CUDA epilogue without the physics in between. */

typedef struct TmjxModel TmjxModel;
typedef struct TmjxClips TmjxClips;

/* sizes derived from the model + task (for buffer allocation by the caller) */
typedef struct TmjxDims {
  int32_t nq, nv, nu, na, nbody, njnt, ncon, nefc;
  int32_t obs_size, reference_obs_size, proprioceptive_obs_size;
  int32_t var_window_size, n_metrics;
  int32_t smem_bytes_per_env, envs_per_block, threads_per_env;
} TmjxDims;

int tmjx_abi_version(void);
const char* tmjx_last_error(void);

/* `blob` = model-constant table produced by track-mjx_b200/model_blob.py (host memory). */
int tmjx_model_create(const void* blob, size_t nbytes, const TmjxTaskConfig* cfg, int device, TmjxModel** out);
void tmjx_model_destroy(TmjxModel* m);
int tmjx_model_dims(const TmjxModel* m, TmjxDims* out);
/* brax EpisodeWrapper limit of the fused wrappers (`wrappers.wrap(env, episode_length=...)`, wrappers.py:18-56): synchronous, takes
 * effect for every later tmjx_step with TMJX_F_AUTORESET. */
int tmjx_model_set_episode_length(TmjxModel* m, int episode_length);

/* Reference clips: HOST pointers to the eight ReferenceClip fields (io/load.py:16-38), each
 * `[n_clips, clip_len, d]` fp32 row-major. The hot subset is packed into one device table. */
int tmjx_clips_create(const TmjxModel* m, const float* position, const float* quaternion, const float* joints,
                      const float* body_positions, const float* velocity, const float* angular_velocity,
                      const float* joints_velocity, const float* body_quaternions, int n_clips, int clip_len,
                      int n_ref_bodies, TmjxClips** out);
void tmjx_clips_destroy(TmjxClips* c);
size_t tmjx_clips_device_bytes(const TmjxClips* c);

/* reset path: qpos/qvel/(act,time,warmstart are zeroed) given -> mjx.forward -> obs; zero reward/done/metrics,
 * zero action_buffer/buffer_index (single_clip_tracking.py:163-205). */
int tmjx_forward(const TmjxModel* m, const TmjxClips* c, TmjxState* s, TmjxOut* o, int n_env, unsigned flags,
                 void* stream);

/* one control step (single_clip_tracking.py:207-320), optionally with the wrappers fused. */
int tmjx_step(const TmjxModel* m, const TmjxClips* c, const float* action, TmjxState* s, TmjxOut* o, int n_env,
              unsigned flags, void* stream);

/* FP32 FMA-throughput microbenchmark used as the roofline denominator (returns TFLOP/s, <0 on error). */
double tmjx_fp32_peak_tflops(int device, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * In-loop policy inference (the acting half of PPO; BASELINE configs[2]).  Replaces, for inference only,
 *   IntentionNetwork.__call__ / Encoder / Decoder   reference track_mjx/agent/mlp_ppo/intention_network.py:14-142
 *   make_inference_fn().policy                      reference track_mjx/agent/mlp_ppo/ppo_networks.py:34-100
 *   running_statistics.normalize                    reference track_mjx/agent/masked_running_statistics.py:217-236
 * Dense layers run on the tcgen05 tensor cores (kind::tf32, fp32 accumulate in TMEM).
 *
 * `params` is ONE host fp32 vector, in this order (W = flax Dense kernel [in, out] row-major):
 *   normaliser mean[obs], std[obs];
 *   encoder hidden_i: W, b, LayerNorm scale, LayerNorm bias   (i = 0 .. n_encoder_layers-1);
 *   fc2_mean W, b; fc2_logvar W, b;
 *   decoder hidden_i: W, b, LayerNorm scale, LayerNorm bias   (i = 0 .. n_decoder_layers-1);
 *   decoder output W [., 2 action_size], b.
 * Noise is supplied by the caller (eps ~ N(0,1); the reference draws it with jax.random inside the policy), so the
 * call is a pure function of its inputs.  All array arguments of tmjx_policy_act are DEVICE pointers, row-major.
 * tmjx_policy_act / tmjx_value_apply are ONE kernel launch each (csrc/tmjx_chain.cuh: 128 environments per CTA walk the whole
 * network; tmjx_policy_launches_per_act == 1).  Environment knobs read at create time: TMJX_POLICY_FUSED=0 (per-layer launches, the
 * A/B reference), TMJX_CHAIN_CLUSTER=2|4 (weight slices multicast over thread-block clusters). */
#define TMJX_POLICY_MAX_LAYERS 8
typedef struct TmjxPolicyDesc {
  int32_t obs_size, reference_obs_size, latent_size, action_size;
  int32_t n_encoder_layers, encoder_layers[TMJX_POLICY_MAX_LAYERS];
  int32_t n_decoder_layers, decoder_layers[TMJX_POLICY_MAX_LAYERS];   /* hidden sizes; the 2*action_size output is implied */
} TmjxPolicyDesc;
typedef struct TmjxPolicy TmjxPolicy;

size_t tmjx_policy_param_count(const TmjxPolicyDesc* d);
int tmjx_policy_create(const TmjxPolicyDesc* d, const float* params, size_t n_params, int device, int max_env, TmjxPolicy** out);
void tmjx_policy_destroy(TmjxPolicy* p);
const char* tmjx_policy_last_error(void);
/* obs [n_env, obs_size] -> action [n_env, action_size] (tanh-squashed).  Optional outputs may be NULL:
 * raw_action [n_env, action_size], log_prob [n_env], logits [n_env, 2 action_size], latent_mean / latent_logvar
 * [n_env, latent_size].  deterministic != 0: z = latent mean, action = tanh(loc) (eps may then be NULL). */
int tmjx_policy_act(const TmjxPolicy* p, const float* obs, const float* eps_latent, const float* eps_action, int deterministic,
                    float* action, float* raw_action, float* log_prob, float* logits, float* latent_mean, float* latent_logvar,
                    int n_env, void* stream);
/* one Dense (+ SiLU + LayerNorm for hidden layers) through the same tensor-core kernel; layers are numbered encoder
 * hidden.., (mean|logvar), decoder hidden.., logits.  x [n_env, ldx], y [n_env, ldy]; pitches >= the padded widths. */
int tmjx_policy_linear(const TmjxPolicy* p, int which, const float* x, int ldx, float* y, int ldy, int n_env, void* stream);
int tmjx_policy_launches_per_act(const TmjxPolicy* p);

/* Value network forward (the baseline / bootstrap value of the PPO loss).  Replaces
 *   networks.make_value_network(...).apply   as built at reference track_mjx/agent/mlp_ppo/ppo_networks.py:180-185 and called at
 *                                            losses.py:151-155 (brax 0.12.3 training/networks.py: normalise, MLP with swish on every
 *                                            layer but the last, Dense to 1, squeeze)
 * on the same tcgen05 TF32 GEMM as the policy.  `params`: ONE host fp32 vector: normaliser mean[obs], std[obs]; per hidden layer W
 * [in, out] row-major, b; output W [., 1], b[1].  obs [n_env, obs_size] and value [n_env] are DEVICE pointers.  The object is
 * released with tmjx_policy_destroy. */
typedef struct TmjxValueDesc {
  int32_t obs_size, n_hidden_layers, hidden_layers[TMJX_POLICY_MAX_LAYERS];
} TmjxValueDesc;
size_t tmjx_value_param_count(const TmjxValueDesc* d);
int tmjx_value_create(const TmjxValueDesc* d, const float* params, size_t n_params, int device, int max_env, TmjxPolicy** out);
int tmjx_value_apply(const TmjxPolicy* v, const float* obs, float* value, int n_env, void* stream);

/* Generalised Advantage Estimation over a rollout (first piece of the learner side).  Replaces
 *   compute_gae   reference track_mjx/agent/mlp_ppo/losses.py:39-101
 * All arrays are DEVICE pointers, time-major [T, B] fp32 (bootstrap_value [B]); outputs vs and advantages [T, B].  Same
 * float32 operation order as the reference, no fused multiply-add: bit-identical to its numpy execution. */
int tmjx_gae(const float* truncation, const float* termination, const float* rewards, const float* values, const float* bootstrap_value,
             float lambda, float discount, float* vs, float* advantages, int T, int B, void* stream);

/* Observation-normaliser statistics (the state `running_statistics.normalize` reads).  Replaces
 *   running_statistics.update   reference track_mjx/agent/masked_running_statistics.py:80-214 (as called at ppo.py:357-361)
 * in three steps so that a multi-GPU caller can all-reduce between them (the reference psums twice inside `update`); the batch
 * is read ONCE (step 1), steps 2 and 3 touch [D] vectors only:
 *   1 tmjx_running_stats_sums : over the LOCAL batch [N, D]: sums[0..D) = sum_rows (x - mean), sums[D..2D) = batch mean,
 *                               sums[2D..3D) = sum_rows (x - batch mean)^2                  -> all-reduce sums[0..D) and N
 *   2 tmjx_running_stats_mean : u = sums[0..D) / (count + increment); mean += u (in place);
 *                               sums[0..D) = this GPU's share of sum_rows (x - old mean)(x - new mean)   -> all-reduce sums[0..D)
 *   3 tmjx_running_stats_apply: count += increment; summed_variance += var;
 *                               std = clip(sqrt(max(summed_variance, 0) / count), std_min, std_max)       (in place)
 * All pointers are DEVICE pointers (count, increment: one float each); sums: 3 D floats; scratch >=
 * tmjx_running_stats_scratch_floats(D) floats, the same buffer for the three calls of one update. */
size_t tmjx_running_stats_scratch_floats(int D);
int tmjx_running_stats_sums(const float* batch, int N, int D, const float* mean, float* sums, float* scratch, void* stream);
int tmjx_running_stats_mean(float* sums, const float* increment, int n_local, int D, const float* count, float* mean, float* scratch,
                            void* stream);
int tmjx_running_stats_apply(const float* var, int D, float std_min, float std_max, float* count, float* summed_variance, float* std,
                             const float* scratch, void* stream);

/* PPO loss head: everything of `compute_ppo_loss` after the network applications.  Replaces
 *   compute_ppo_loss   reference track_mjx/agent/mlp_ppo/losses.py:154-245 (+ compute_gae :39-101 inside it)
 * and additionally returns the gradients a backward pass through the networks starts from.  All arrays are DEVICE pointers,
 * time-major fp32: logits [T, B, 2A] (loc | raw scale of brax's NormalTanhDistribution), latent_mean / latent_logvar [T, B, L],
 * baseline / reward / discount / truncation / behaviour_log_prob [T, B], bootstrap_value [B], raw_action [T, B, A], eps_entropy
 * [T, B, A] (the standard-normal draw the reference takes from `entropy_key`).  Outputs: losses[8] = total, policy, value,
 * latent KL, entropy loss, advantage mean, advantage std, 1 / (std + 1e-8); vs [T, B]; advantages [T, B] (normalised when
 * hyper->normalize_advantage); d_logits [T, B, 2A], d_latent_mean / d_latent_logvar [T, B, L], d_baseline [T, B] = d total /
 * d input (vs and advantages are stop_gradient in the reference, so the bootstrap value has no gradient).  scratch >=
 * tmjx_ppo_loss_scratch_floats(T, B) floats.  Deterministic (fixed summation order). */
typedef struct TmjxPpoHyper {
  float entropy_cost, kl_weight, discounting, reward_scaling, gae_lambda, clipping_epsilon;
  int32_t normalize_advantage;
} TmjxPpoHyper;
size_t tmjx_ppo_loss_scratch_floats(int T, int B);
int tmjx_ppo_loss_head(const float* logits, const float* latent_mean, const float* latent_logvar, const float* baseline,
                       const float* bootstrap_value, const float* reward, const float* discount, const float* truncation,
                       const float* raw_action, const float* behaviour_log_prob, const float* eps_entropy, int T, int B, int A, int L,
                       const TmjxPpoHyper* hyper, float* losses, float* vs, float* advantages, float* d_logits, float* d_latent_mean,
                       float* d_latent_logvar, float* d_baseline, float* scratch, void* stream);

/* Optimiser step on flat fp32 DEVICE buffers of n elements (parameters, gradients, Adam moments mu / nu).  Replaces the update of
 *   optax.chain(optax.clip_by_global_norm(10.0), optax.adam(learning_rate))   reference track_mjx/agent/mlp_ppo/ppo.py:517-520
 * (optax 0.2.5, not vendored): g = grads * grad_scale (1 / world_size after a SUM all-reduce = the reference's pmean);
 * if max_grad_norm > 0 and not ||g|| < max_grad_norm: g = (g / ||g||) * max_grad_norm; mu = b1 mu + (1 - b1) g;
 * nu = b2 nu + (1 - b2) g^2; params += -lr (mu / (1 - b1^count)) / (sqrt(nu / (1 - b2^count)) + eps).  count = the 1-based
 * step number (optax's count after its increment).  grad_norm_out (device, one float, may be NULL) receives ||g|| before
 * clipping.  scratch >= tmjx_adam_scratch_floats() floats, 8-byte aligned.  Deterministic. */
size_t tmjx_adam_scratch_floats(void);
int tmjx_adam_step(float* params, const float* grads, float* mu, float* nu, size_t n, float learning_rate, float b1, float b2, float eps,
                   float max_grad_norm, float grad_scale, int count, float* grad_norm_out, float* scratch, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * XLA-FFI custom-call handlers (csrc/tmjx_xla_ffi.cc): what `jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(lib.tmjx_step_ffi),
 * platform="CUDA")` binds so that the step stays inside `jit` / `lax.scan` / `pmap` like the reference's env (ppo.py:333-340, 409).
 * They decode XLA's call frame (operands: [action,] the 25 TmjxState leaves, donated; results: the same 25 leaves, then obs, reward,
 * done, metrics, cur_frame; attributes: model, clips, flags as int64) and forward to tmjx_step / tmjx_forward on XLA's stream.
 * tmjx_xla_ffi_available() is 1 when the library was built against jaxlib's own xla/ffi/api/c_api.h, 0 when against the recalled
 * subset shipped next to the adapter; tmjx_ffi_selftest builds a call frame by hand and runs a handler through it (tests). */
struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;
struct XLA_FFI_Error* tmjx_step_ffi(struct XLA_FFI_CallFrame* call_frame);
struct XLA_FFI_Error* tmjx_forward_ffi(struct XLA_FFI_CallFrame* call_frame);
int tmjx_xla_ffi_available(void);
int tmjx_ffi_selftest(int is_step, const void* model, const void* clips, const float* action, int nu, const TmjxState* s, const TmjxOut* o,
                      const int* state_dims, const int* state_is_int, const int* out_dims, int n_env, unsigned flags, void* stream, int alias);
const char* tmjx_ffi_selftest_error(void);

/* Refresh an acting policy (tmjx_policy_create) or value network (tmjx_value_create) from a flat DEVICE parameter vector in the layout
 * of its create call (normaliser mean, std first): the hand-off from the learner to the actor after every optimiser / normaliser
 * update.  The reference has no such step because its policy closure reads `training_state.params` directly (ppo.py:326-328). */
int tmjx_policy_set_params(TmjxPolicy* p, const float* params_device, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Backward pass of the intention network and the value network: what `jax.value_and_grad(compute_ppo_loss)` differentiates through
 * the networks for one minibatch.  Replaces
 *   gradient_update_fn = gradients.gradient_update_fn(loss_fn, optimizer, pmap_axis_name, has_aux=True)   reference ppo.py:621-623
 *   (the value_and_grad half; the optimiser half is tmjx_adam_step, the pmean is the caller's NCCL all-reduce of `grads`)
 * The trainer owns ONE flat fp32 DEVICE parameter buffer = [policy vector | value vector], each in the layout of tmjx_policy_create /
 * tmjx_value_create (normaliser mean, std at the head of each; they receive zero gradient), and a gradient buffer of the same
 * layout.  One minibatch:
 *   tmjx_trainer_policy_forward   obs [rows, obs], eps_latent [rows, L] -> logits [rows, 2 A], latent mean / logvar [rows, L]
 *   tmjx_trainer_value_forward    obs [rows, obs] -> value [rows]
 *   tmjx_ppo_loss_head            -> d_logits, d_latent_mean, d_latent_logvar, d_baseline
 *   tmjx_trainer_value_backward / tmjx_trainer_policy_backward   -> grads (each parameter written once per call)
 *   all-reduce(grads); tmjx_adam_step(params, grads, ...); tmjx_trainer_sync (parameters -> GEMM operand layouts)
 * All array arguments are DEVICE pointers, row-major; rows <= max_rows.  Dense layers, dgrad and wgrad run on the tcgen05 TF32 GEMM. */
typedef struct TmjxTrainer TmjxTrainer;
int tmjx_trainer_create(const TmjxPolicyDesc* pd, const TmjxValueDesc* vd, const float* policy_params /* host */, const float* value_params /* host */,
                        int device, int max_rows, TmjxTrainer** out);
void tmjx_trainer_destroy(TmjxTrainer* t);
size_t tmjx_trainer_param_count(const TmjxTrainer* t);          /* policy vector + value vector */
size_t tmjx_trainer_policy_param_count(const TmjxTrainer* t);   /* offset of the value vector */
int tmjx_trainer_buffers(TmjxTrainer* t, float** params, float** grads);
int tmjx_trainer_sync(TmjxTrainer* t, void* stream);
int tmjx_trainer_policy_forward(TmjxTrainer* t, const float* obs, const float* eps_latent, int rows, float* logits, float* latent_mean,
                                float* latent_logvar, void* stream);
int tmjx_trainer_policy_backward(TmjxTrainer* t, const float* d_logits, const float* d_latent_mean, const float* d_latent_logvar, int rows,
                                 void* stream);
int tmjx_trainer_value_forward(TmjxTrainer* t, const float* obs, int rows, float* value, void* stream);
int tmjx_trainer_value_backward(TmjxTrainer* t, const float* d_value, int rows, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TMJX_H_ */
