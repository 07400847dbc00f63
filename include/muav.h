/* muav.h -- C ABI of libmuav_b200.so: batched Multi-UAV-TA "Windowed Pop-up Strike"
 * step path for NVIDIA B200 (sm_100a).
 *
 * Every entry point takes plain pointers and sizes; pointers prefixed d_ are
 * DEVICE pointers owned by the caller (e.g. torch tensors), h_ are HOST
 * pointers.  All calls are stream-ordered on `stream` (a cudaStream_t passed as
 * void*; NULL = default stream) and return 0 on success or a negative errno /
 * -1000-cudaError on failure.  No exceptions cross this boundary.
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   muav_step / muav_rollout  -> MultiUAVEnv.step             mUAV_TA/DroneEnv.py:774-1206
 *                                (+ UAV/Task bookkeeping       mUAV_TA/DroneEnvComponents.py:55-179,280-326)
 *   muav_allocate (also fused in muav_step via muav_alloc_opts)
 *                             -> HungarianAllocator.allocate_tasks
 *                                                              TaskAllocation/OptimizationBased/HungarianAllocator.py:72-208
 *                                and the driver glue           experiments/paper_eval.py:85-101, experiments/wps_eval.py:55-73
 *   muav_lsap                 -> scipy.optimize.linear_sum_assignment call at HungarianAllocator.py:181
 *   muav_avoid_obstacles      -> core_sim.SimCore.avoid_obstacles  core_sim/src/sim_core.rs:24-59 (PyO3 export lib.rs:10-18)
 *   muav_tokens_pair          -> build_pair_tokens             TaskAllocation/Hybrid/PairCostHybrid.py:31-65
 *                                (build_att_tokens             TaskAllocation/Hybrid/AttentionRAH.py:50-173)
 *   muav_tokens_context       -> build_context_pair_tokens     TaskAllocation/Hybrid/ContextPairHybrid.py:33-78
 *   muav_tokens_commit        -> enrich_commit_tokens          TaskAllocation/Hybrid/AttentionCommit.py:49-62
 *   muav_tokens_escort        -> build_escort_tokens           TaskAllocation/Hybrid/AttentionEscort.py:76-241
 *   muav_observe              -> _generate_observations/get_task_info  mUAV_TA/DroneEnv.py:365-492
 *   muav_metrics              -> calculate_metrics / compute_s_wps / compute_s_esc  DroneEnv.py:1231-1337,2002-2011
 *   muav_pair_mask            -> _expert_mask / _selected_mask experiments/train_pair_cost.py:53-70, PairCostHybrid.py:293-306
 *   muav_att_pair_scores, muav_att_context_pair_scores
 *                             -> AttPairNet / AttContextPairNet forward + act(explore=False)
 *                                                              PairCostHybrid.py:89-151,266-278, ContextPairHybrid.py:81-151,235-246
 *   muav_alloc_opts.planner   -> UrgencyPair / UrgencyCommit / UrgencyCoalition .plan, AttentionCommit / AttentionEscort
 *                                ._plan_from_scores            PairCostHybrid.py:520-550, AttentionCommit.py:266-357,
 *                                                              AttentionEscort.py:500-517,720-767
 *   muav_alloc_opts.planner = 6 -> PerformanceImpact.allocate_tasks(max_tasks_per_agent=1)
 *                                                              TaskAllocation/MarketBased/PerformanceImpact.py:59-224
 *                                (expand_slot_keys / agent_eligible: TaskAllocation/MarketBased/CBBA.py:27-65)
 *   muav_reset_upload / muav_state_bytes / muav_tape_bytes / muav_snapshot / muav_field_info / muav_record_bytes:
 *                                device-resident state that replaces the UAV / Task / Threat objects (host packing at
 *                                reset, snapshots for the object proxies); reset itself (DroneEnv.py:522-762) stays in Python.
 */
#ifndef MUAV_H_
#define MUAV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MUAV_N_UAV_TYPES 7  /* R1 R2 E1 F1 F2 T1 T2   (MultiDroneEnvData.py:15) */
#define MUAV_N_TASK_TYPES 6 /* Hold Rec Att Def Int Det (MultiDroneEnvData.py:18) */
#define MUAV_MAX_GROUPS 8
#define MUAV_MAX_AGENTS 64
#define MUAV_MAX_TASK_CAP 640 /* task slots (open or still referenced tasks) */
#define MUAV_MAX_ID_CAP 2048  /* tasks ever created in one episode */
#define MUAV_MAX_QUEUE 32

typedef struct muav_config {
  /* capacities of one environment record */
  int32_t n_agents, task_cap, n_threats, queue_cap, event_cap, n_obstacles, n_groups;
  /* scenario constants (DroneEnv.py:145-147,101) */
  int32_t n_tasks_cfg, max_tasks, max_time_steps;
  /* flags (agentEnvOptions, MultiDroneEnvUtils.py:5-105) */
  int32_t multiple_tasks_per_agent, early_terminate, capability_mask, saturate_mask;
  int32_t hard_windows, burst_mode, dual_region_bursts, share_knowledge, escort_enabled;
  int32_t threat_delay, window_length, burst_size, commit_horizon;
  int32_t escort_required_agents, escort_type_mask;
  int32_t tape_words[3];    /* words per env in each RNG tape: agent, tgt, mission */
  int32_t group_start[MUAV_MAX_GROUPS + 1]; /* threat ids of group g: [group_start[g], group_start[g+1]) */
  int32_t duration[MUAV_N_TASK_TYPES];
  int32_t id_cap;           /* task ids per episode (>= task_cap); 0 = task_cap */
  double arrival_rate, sense_radius, miss_penalty, on_time_bonus, dynamic_idle_penalty, reassign_penalty;
  double escort_radius, escort_requirement, escort_intercept_radius, mutual_support_radius;
  double threat_gen_prob, threat_wide, max_coord, area_w, area_h, base_x, base_y, contact_line;
  double rw[8]; /* action distance quality s_quality time alloc time_penaulty step */
  double speed[MUAV_N_UAV_TYPES];  /* per step, already scaled by frame rate (DroneEnv.py:611,725) */
  double engage[MUAV_N_UAV_TYPES];
  double cap_table[MUAV_N_UAV_TYPES][MUAV_N_TASK_TYPES];
} muav_config;

/* Allocator fused in front of each step (NULL opts = actions come from d_actions). */
typedef struct muav_alloc_opts {
  int32_t mode;            /* 0 none (actions come from d_actions);
                              1 HungarianAllocator.should_replan rule: t - last_plan_step >= interval or a listed event tag
                                (HungarianAllocator.py:27-41,88), force=False;
                              2 hybrid rule (wps_eval.py:64-73, escort_eval.py:52-58): t == 0 or t % interval == 0 or a
                                listed event tag, then allocate_tasks(force=True);
                              3 allocate_tasks(force=True) unconditionally (the caller evaluated its own rule) */
  int32_t replan_interval; /* 20 Local-Hungarian, 12 Coalition-Hungarian, 15 hybrids */
  int32_t event_mask;      /* bit i = tag i triggers (0 Reset_Allocation 1 Agent_Fail 2 New_Threat 3 Escort_Created 4 Escort_Retired) */
  int32_t use_visibility;  /* agent_known_ids=env.agent_visibility_map() */
  int32_t pair_tokens;     /* 1: task list = build_pair_tokens' open list (PairCostHybrid.py:34-36, AttentionRAH.py:67-71)
                                 and d_edge_scores is indexed in token space [live agent row, token task column];
                              2: d_edge_scores is indexed [live agent row, position in d_task_order] (the layout of
                                 muav_tokens_escort; AttentionEscort.edge_score_dict, AttentionEscort.py:478-489) */
  int32_t score_rows, score_cols; /* edge score tensor shape per env (max_agents, max_tasks) */
  int32_t score_f64;       /* 1: d_edge_scores points to double, else float */
  int32_t planner;         /* 0 none; 1 UrgencyCommit.plan (AttentionCommit.py:310-357): priorities 0.6 urg + 0.4 scarcity,
                                committed agents reserved, lock ranking, commit_until writes;
                              2 UrgencyCoalition.plan (AttentionEscort.py:720-767): engineered edge scores, committed agents
                                reserved, every assigned agent locked for commit_horizon;
                              3 AttentionCommit._plan_from_scores (AttentionCommit.py:266-300): priorities
                                0.35 urg + 0.40 d_plan_pri[column] + 0.25 scarcity over build_att_tokens' open list, committed
                                agents reserved, free assigned agents whose d_plan_commit[row] >= commit_threshold are locked;
                              4 AttentionEscort._plan_from_scores (AttentionEscort.py:500-517): d_edge_scores in
                                muav_tokens_escort layout over d_task_order, committed agents reserved, every assigned
                                agent locked for commit_horizon;
                              5 UrgencyPair.plan (PairCostHybrid.py:520-550): urgency_edge_scores (:68-86) on the valid
                                edges of build_pair_tokens(score_cols = max_tasks, score_rows = max_agents), no locks;
                              6 PerformanceImpact.allocate_tasks (max_tasks_per_agent below) (MarketBased/PerformanceImpact.py:59-224,
                                slots and eligibility from MarketBased/CBBA.py:27-65) instead of the Hungarian allocator:
                                mode / replan_interval / event_mask / use_visibility / d_reserved as for planner 0;
                              7 CBBAReplan.allocate_tasks (max_tasks_per_agent below) (MarketBased/CBBA_Replan.py:15-69 around
                                MarketBased/CBBA.py:68-324; a fresh CBBA(seed + n_replans) per replan, seed = d_cbba_seed[env]):
                                same options as planner 6.  Bit-exact with the reference run under PYTHONHASHSEED=0 (its
                                auction order starts from a set of strings, see csrc/muav_cbba.cuh) */
  int32_t order_hint_mode; /* mode == 0 only (actions come from the caller): the replan rule (1 / 2 / 3, with replan_interval,
                              event_mask and planner as above) that the caller's allocator follows, used solely to fill
                              muav_step_out.d_env_order_next; 0 = no hint */
  double commit_fraction;  /* UrgencyCommit(commit_fraction=0.35) */
  double max_coord;        /* HungarianAllocator(max_coord=...) */
  const void* d_edge_scores;   /* [E, score_rows, score_cols] float (or double when score_f64) or NULL;
                                  pair_tokens=0: indexed [agent id, task index] */
  const double* d_priorities;  /* [E, id_cap] by task index = id - 1 (task_priorities) or NULL */
  const uint8_t* d_reserved;   /* [E, n_agents] 1 = excluded (reserved_agent_names) or NULL */
  const int32_t* d_task_order; /* [E, id_cap] the `tasks` argument: task indices (id-1) in the caller's order, -1 terminated;
                                  NULL = every open task in id order (_open_tasks, paper_eval.py:96-101) */
  const float* d_plan_pri;     /* planner 3: [E, score_cols] AttCommitNet priorities by token task column */
  const float* d_plan_commit;  /* planner 3: [E, score_rows] AttCommitNet commit gates by live-agent row */
  double commit_threshold;     /* planner 3: AttentionCommit(commit_threshold=0.5) */
  const int32_t* d_cbba_seed;  /* planner 7: [E] the `seed` argument of CBBAReplan (the drivers pass the episode seed) or NULL = 0 */
  int32_t max_tasks_per_agent; /* planners 6 and 7: 0 / 1 = one task per agent (what the reference drivers pass); 2..4 = bundles
                                  (planner 6: the full inclusion phase of PerformanceImpact.py:106-165 over agent paths with insertion
                                  points, consensus clean-up :168-205; planner 7: CBBA.py:128-190 with bundles of that
                                  length, listed in the order the slots were won).  The step takes the FIRST task of every path
                                  (_apply_assign, wps_eval.py:55-61); the whole plan goes to d_bundle_pairs */
  int32_t reserved0;
  int32_t* d_bundle_pairs;     /* planners 6 / 7 with bundles, optional: [E, n_agents * max_tasks_per_agent] (agent id << 16) | task
                                  id of every path / bundle entry in the order allocate_tasks returns them (planner 6: agents
                                  ascending, path order; planner 7: live-agent order, bundle order, CBBA.py:192-204) */
  int32_t* d_n_bundle_pairs;   /* [E] number of entries written to d_bundle_pairs (0 when the call did not replan) */
} muav_alloc_opts;

/* Per-step outputs (any pointer may be NULL). */
typedef struct muav_step_out {
  double* d_reward;     /* [E]  the scalar every agent receives (DroneEnv.py:1164-1178; F_Reward on the last step :1202) */
  uint8_t* d_terminated; /* [E] */
  uint8_t* d_truncated;  /* [E] */
  int32_t* d_n_events;   /* [E] */
  int32_t* d_events;     /* [E, event_cap] drained events (infos['events']): (arg+1)<<8 | tag */
  int32_t* d_n_pairs;    /* [E] allocator output this step */
  int32_t* d_pairs;      /* [E, n_agents] (agent_id<<16) | task_id, in reference order */
  int32_t* d_n_open;     /* [E] len(env.last_tasks_info) after the step */
  /* Scheduling hint, no effect on results.  The warps of a CTA walk the step phase by phase together, so a CTA is as
   * slow as its slowest environment; environments that will run the allocator at the next step are therefore grouped
   * into the same CTAs.  d_env_order (NULL = identity): permutation of [0, E) giving the environment of each launch
   * slot.  d_env_order_next (NULL = none): int32 [E + 2], receives the permutation for the NEXT launch (replanning
   * environments first); its two trailing counters must be zero on entry -- the kernel zeroes those of d_env_order,
   * so two zero-initialised buffers used alternately need no further care. */
  const int32_t* d_env_order;
  int32_t* d_env_order_next;
  /* Optional workspace int32 [E, n_agents, 2].  When given, a single-step call with a fused allocator (n_steps == 1,
   * opts->mode != 0) runs as two kernels: the allocator for the environments whose replan rule fires (the others leave
   * after reading their record header), then the step for all environments without the allocator's scratch, i.e.
   * half as many environments again per SM.  Same results as the one-kernel form; on B200 the one-kernel form is
   * faster (the allocator of a few environments hides among the others' steps), so callers normally leave this NULL. */
  int32_t* d_actions_ws;
} muav_step_out;

/* Optional pair-token emission fused at the end of each step (saves the separate muav_tokens_pair pass):
 * for every env whose hybrid replan rule will fire before the NEXT step -- (T % interval == 0) or a listed
 * event tag among the events this step drained -- d_need[e] = 1 and the env's token rows are written. */
typedef struct muav_token_out {
  float* d_task_feats;    /* [E, max_tasks, 13] */
  uint8_t* d_task_mask;   /* [E, max_tasks] 1 = padding */
  float* d_agent_feats;   /* [E, max_agents, 12] */
  uint8_t* d_agent_mask;  /* [E, max_agents] */
  float* d_edge_valid;    /* [E, max_agents, max_tasks] */
  int32_t* d_task_ids;    /* [E, max_tasks] */
  uint8_t* d_need;        /* [E] */
  int32_t max_tasks, max_agents, interval, event_mask;
  float* d_context;       /* optional [E, 8]: build_context_summary (ContextPairHybrid.py:33-70) of the same tokens */
  int32_t agent_feat_dim; /* 0 / 12: pair tokens; 13: commit tokens (enrich_commit_tokens, AttentionCommit.py:49-62):
                             d_agent_feats is then [E, max_agents, 13] and d_edge_valid may be NULL; 16: escort tokens
                             (build_escort_tokens, AttentionEscort.py:76-241; escort_enabled configurations): d_task_feats
                             [E, max_tasks, 22], d_agent_feats [E, max_agents, 16], d_edge_valid required */
  int32_t* d_task_order;  /* escort tokens only, optional [E, id_cap]: the muav_tokens_escort d_task_order output (rank of
                             each task id in the token order, the d_task_order input of planner 4) */
} muav_token_out;

const char* muav_version(void);
size_t muav_config_size(void);
size_t muav_record_bytes(const muav_config* cfg);
size_t muav_scratch_bytes(const muav_config* cfg);
/* bytes at the start of a record that the step kernel stages on chip (the rest is touched in place, see csrc/muav_layout.h) */
size_t muav_hot_bytes(const muav_config* cfg);
int muav_num_fields(void);
/* name/offset/count/element size of field `idx` of the record for this config */
int muav_field_info(const muav_config* cfg, int idx, const char** name, int64_t* offset, int64_t* count, int32_t* elem_size);
int muav_header_index(const char* name); /* index into the hi (int32) or hf (double) header arrays, -1 if unknown */

/* K env steps for E environments.  d_records: [E, record_bytes]; d_tapes: [E, sum(tape_words)] uint32;
 * d_actions: [E, n_agents, 2] int32 ordered (agent_id, index into last_tasks_info), agent_id = -1 ends the list
 * (ignored when opts->mode != 0).  */
int muav_step(const muav_config* cfg, void* d_records, const uint32_t* d_tapes, const int32_t* d_actions,
              const muav_alloc_opts* opts, const muav_step_out* out, const muav_token_out* tok, int n_envs, int n_steps,
              void* stream);
/* Allocator only (HungarianAllocator.allocate_tasks + _apply_assign): no env step.  Writes out->d_pairs / d_n_pairs and,
 * when d_actions_out != NULL, the ordered action list [E, n_agents, 2] (agent_id, index into last_tasks_info; agent_id -1
 * terminates) that muav_step accepts.  Per-env allocator state (last_plan_step, n_replans, n_calls) advances. */
int muav_allocate(const muav_config* cfg, void* d_records, const muav_alloc_opts* opts, const muav_step_out* out,
                  int32_t* d_actions_out, int n_envs, void* stream);
/* Same call with HOST action / output buffers: copies in, runs, copies out, synchronises.  The staging block is a
 * stream-ordered allocation of the call itself (thread-safe, any device); muav_ctx_step_host below avoids it. */
int muav_step_host(const muav_config* cfg, void* d_records, const uint32_t* d_tapes, const int32_t* h_actions,
                   const muav_alloc_opts* opts, const muav_token_out* tok, double* h_reward, uint8_t* h_terminated,
                   uint8_t* h_truncated, int n_envs, int n_steps, void* stream, const int32_t* d_env_order,
                   int32_t* d_env_order_next);

/* Host-buffer context.  The reference's env object owns its buffers (MultiUAVEnv instance state, mUAV_TA/DroneEnv.py:
 * 73-323); here a caller-owned handle owns the device staging block and the pinned host image of the per-step outputs
 * for n_envs environments of ONE device.  The library keeps no mutable process-wide state: calls on different handles
 * are independent (one handle per thread / per device); a handle must not be used from two threads at once. */
typedef struct muav_ctx muav_ctx;
int muav_ctx_create(const muav_config* cfg, int n_envs, int device, muav_ctx** out);
void muav_ctx_destroy(muav_ctx* ctx);
/* muav_step_host on the handle's buffers: one H2D copy (actions), the step, ONE D2H copy (reward | terminated |
 * truncated, packed), one synchronisation.  MultiUAVEnv.step with host actions in / host rewards out (DroneEnv.py:774). */
int muav_ctx_step_host(muav_ctx* ctx, void* d_records, const uint32_t* d_tapes, const int32_t* h_actions,
                       const muav_alloc_opts* opts, const muav_token_out* tok, double* h_reward, uint8_t* h_terminated,
                       uint8_t* h_truncated, int n_steps, void* stream, const int32_t* d_env_order,
                       int32_t* d_env_order_next);
/* muav_allocate with the ordered action list delivered to HOST memory [E, n_agents, 2] (HungarianAllocator.allocate_tasks
 * returning its dict to the caller, HungarianAllocator.py:72-208): allocator kernel, one D2H copy, one synchronisation. */
int muav_ctx_allocate_host(muav_ctx* ctx, void* d_records, const muav_alloc_opts* opts, const muav_step_out* out,
                           int32_t* h_actions_out, void* stream);

/* K fused (plan -> allocate -> step) iterations with the state resident on chip: muav_step with a fused allocator and
 * no external actions (the episode loops of experiments/wps_eval.py:96-140 and escort_eval.py:85-200 without the
 * Python round trip). */
int muav_rollout(const muav_config* cfg, void* d_records, const uint32_t* d_tapes, const muav_alloc_opts* opts,
                 const muav_step_out* out, const muav_token_out* tok, int n_envs, int n_steps, void* stream);
/* Bytes of device memory the caller allocates for n_envs records / RNG tapes. */
size_t muav_state_bytes(const muav_config* cfg, int n_envs);
size_t muav_tape_bytes(const muav_config* cfg, int n_envs);
/* Upload the host-packed reset state (MultiUAVEnv.reset runs on the host, DroneEnv.py:522-762): h_records
 * [n_envs, record_bytes], h_tapes [n_envs, sum(tape_words)] uint32 -> device, stream-ordered. */
int muav_reset_upload(const muav_config* cfg, void* d_records, uint32_t* d_tapes, const void* h_records,
                      const uint32_t* h_tapes, int n_envs, void* stream);
/* Copy the record of environment `env_index` to the host (record_bytes; decoded with muav_field_info) and wait for it:
 * what the object proxies and the parity snapshots read. */
int muav_snapshot(const muav_config* cfg, const void* d_records, int env_index, void* h_record, void* stream);

/* Batched rectangular LSAP: B problems, cost [B, nr_max, nc_max] row-major with per-problem sizes.
 * out_col4row [B, nr_max]: column assigned to each row or -1 (SciPy tie-breaking reproduced). */
int muav_lsap(const double* d_cost, const int32_t* d_nr, const int32_t* d_nc, int nr_max, int nc_max,
              int32_t* d_col4row, int n_problems, void* stream);

/* core_sim.SimCore.avoid_obstacles over a batch: pos [N,2], movement [N,2], obstacles [M,3] -> out [N,2] */
int muav_avoid_obstacles(const double* d_pos, const double* d_move, const double* d_obstacles, int n_obstacles,
                         double* d_out, int n, void* stream);

/* Episode metrics: out [E, MUAV_N_METRICS] doubles, order given by muav_metric_name(). */
#define MUAV_N_METRICS 30
const char* muav_metric_name(int idx);
int muav_metrics(const muav_config* cfg, const void* d_records, double* d_out, int n_envs, void* stream);

/* Pair tokens: task_feats [E,max_tasks,13] f32, task_mask [E,max_tasks] u8 (1 = padding), agent_feats [E,max_agents,12] f32,
 * agent_mask [E,max_agents] u8, edge_valid [E,max_agents,max_tasks] f32, task_ids [E,max_tasks] i32 (0 = padding). */
int muav_tokens_pair(const muav_config* cfg, const void* d_records, int max_tasks, int max_agents,
                     float* d_task_feats, uint8_t* d_task_mask, float* d_agent_feats, uint8_t* d_agent_mask,
                     float* d_edge_valid, int32_t* d_task_ids, int n_envs, void* stream);

/* Context-pair tokens = build_context_pair_tokens(env, raw) (ContextPairHybrid.py:33-78): the pair tokens plus the team /
 * situation vector context [E, 8] f32.  raw != 0 selects the per-entity feature variant of build_att_tokens(raw=True)
 * (AttentionRAH.py:86-97,140-146): task_feats [E,max_tasks,9], agent_feats [E,max_agents,11], context [E,1]. */
int muav_tokens_context(const muav_config* cfg, const void* d_records, int max_tasks, int max_agents, int raw,
                        float* d_task_feats, uint8_t* d_task_mask, float* d_agent_feats, uint8_t* d_agent_mask,
                        float* d_edge_valid, int32_t* d_task_ids, float* d_context, int n_envs, void* stream);

/* Commit tokens = enrich_commit_tokens(build_att_tokens(env)) (AttentionCommit.py:49-62, AttentionRAH.py:50-173):
 * as muav_tokens_pair but agent_feats is [E, max_agents, 13] (last column: remaining commit-lock fraction) and there is
 * no edge_valid. */
int muav_tokens_commit(const muav_config* cfg, const void* d_records, int max_tasks, int max_agents, float* d_task_feats,
                       uint8_t* d_task_mask, float* d_agent_feats13, uint8_t* d_agent_mask, int32_t* d_task_ids, int n_envs,
                       void* stream);

/* Escort tokens = build_escort_tokens(env, max_tasks=48, max_agents=16) (AttentionEscort.py:76-241, with
 * _open_tasks_residual :31-43, _threat_stats :46-65, _task_priority_key :68-73): tasks known to at least one live agent
 * (all open tasks if none), sorted by the priority key (stable), first max_tasks kept.
 * task_feats [E,max_tasks,22] f32, task_mask [E,max_tasks] u8, agent_feats [E,max_agents,16] f32, agent_mask [E,max_agents] u8,
 * edge_valid [E,max_agents,max_tasks] f32, task_ids [E,max_tasks] i32 in column order (0 = padding); d_task_order
 * (optional) [E, id_cap] i32: the kept task indices (id-1) in column order, -1 terminated -- the `tasks` argument and the
 * score layout of planner 4 (muav_alloc_opts). */
int muav_tokens_escort(const muav_config* cfg, const void* d_records, int max_tasks, int max_agents, float* d_task_feats22,
                       uint8_t* d_task_mask, float* d_agent_feats16, uint8_t* d_agent_mask, float* d_edge_valid,
                       int32_t* d_task_ids, int32_t* d_task_order, int n_envs, void* stream);

/* Training masks over the pair-token grid: mask [E, max_agents, max_tasks] f32, 1 where the allocator output of this step
 * (muav_step_out.d_pairs / d_n_pairs) pairs token row i (i-th live agent) with token column j (d_task_ids).
 * require_valid != 0: only through valid edges = _expert_mask (experiments/train_pair_cost.py:53-70, imitation of the
 * Global-Hungarian teacher); require_valid == 0: PairCostHybrid._selected_mask (PairCostHybrid.py:293-306). */
int muav_pair_mask(const muav_config* cfg, const void* d_records, const int32_t* d_pairs, const int32_t* d_n_pairs,
                   const int32_t* d_task_ids, const float* d_edge_valid, int max_tasks, int max_agents, int require_valid,
                   float* d_mask, int n_envs, void* stream);

/* Observation tensors (DroneEnv.py:365-492).  tasks_info [E, max_rows, 21] f64 per open task:
 * id, x/max_coord, y/max_coord, status, current_reqs[6], alloc_reqs[6], init_time, end_time, type_idx, unmet, age
 * (status = -1 marks padding rows); pad_mask [E,max_rows] u8 ("mask"); legal_mask [E, n_agents, max_rows] u8;
 * agent_obs [E, n_agents, 9] f64: agent_position[2], agent_caps[6], alloc_task; event_flags [E,5] f32;
 * n_rows [E] i32 = number of rows the reference would emit before padding (> max_rows means truncated). */
#define MUAV_OBS_TASK_DIM 21
#define MUAV_OBS_AGENT_DIM 9
int muav_observe(const muav_config* cfg, const void* d_records, int max_rows, double* d_tasks_info, uint8_t* d_pad_mask,
                 uint8_t* d_legal_mask, double* d_agent_obs, float* d_event_flags, int32_t* d_n_rows, int n_envs,
                 void* stream);

/* Fused Att-Pair scorer forward (AttPairNet, PairCostHybrid.py:89-151; scores = tanh(logits) * clamp * edge_valid,
 * :266-278) for n environments in one launch.  d_params: the module's float32 parameters packed in one buffer,
 * offsets (in floats) below; d_env_idx: optional list of environment indices (NULL = 0..n-1) into the token tensors
 * and d_scores [E, max_agents, max_tasks]; d_need: optional [E] flags (muav_token_out.d_need) -- environments whose flag
 * is 0 are skipped on the device, so no host synchronisation is needed to select the replanning subset.
 * d_model 64, 4 heads, 1 encoder layer, feed-forward 128 are fixed. */
typedef struct muav_attpair_offsets {
  int32_t agent_proj_w, agent_proj_b, task_proj_w, task_proj_b, type_embed;
  int32_t enc_in_w, enc_in_b, enc_out_w, enc_out_b, enc_l1_w, enc_l1_b, enc_l2_w, enc_l2_b;
  int32_t enc_n1_w, enc_n1_b, enc_n2_w, enc_n2_b;
  int32_t a2t_in_w, a2t_in_b, a2t_out_w, a2t_out_b, t2a_in_w, t2a_in_b, t2a_out_w, t2a_out_b;
  int32_t head1_w, head1_b, head2_w, head2_b, head3_w, head3_b;
  /* AttContextPairNet (ContextPairHybrid.py:81-151) only, else 0: ctx_proj [8 -> 64]; head1_w then has 256 input rows
   * (agent, task, product, context) instead of 192 */
  int32_t ctx_proj_w, ctx_proj_b, has_context;
} muav_attpair_offsets;
int muav_att_pair_scores(const float* d_params, const muav_attpair_offsets* offsets, const float* d_task_feats,
                         const uint8_t* d_task_mask, const float* d_agent_feats, const uint8_t* d_agent_mask,
                         const float* d_edge_valid, const int32_t* d_env_idx, const uint8_t* d_need, int n, int max_tasks,
                         int max_agents, float score_clamp, float* d_scores, void* stream);
/* Same kernel for AttContextPairNet: d_context [E, 8] (offsets->has_context must be set). */
int muav_att_context_pair_scores(const float* d_params, const muav_attpair_offsets* offsets, const float* d_task_feats,
                                 const uint8_t* d_task_mask, const float* d_agent_feats, const uint8_t* d_agent_mask,
                                 const float* d_edge_valid, const float* d_context, const int32_t* d_env_idx,
                                 const uint8_t* d_need, int n, int max_tasks, int max_agents, float score_clamp,
                                 float* d_scores, void* stream);

/* Fused AttCoalitionNet forward (TaskAllocation/Hybrid/AttentionEscort.py:244-330: the Att-Pair architecture at d_model
 * 128, 4 heads, 2 encoder layers, feed-forward 512) on escort tokens (muav_tokens_escort layout: task features [., 22],
 * agent features [., 16]): scores = sigmoid(clip(logits, -20, 20)) * edge_valid (AttentionEscort.act without exploration,
 * :449-466), 0 on padded rows / columns; the d_edge_scores input of planner 4.  Weights packed like muav_attpair_offsets
 * (transposed, offsets in floats).  max_agents <= 16, max_agents + max_tasks <= 64. */
typedef struct muav_attcoal_offsets {
  int32_t agent_proj_w, agent_proj_b, task_proj_w, task_proj_b, type_embed;
  int32_t enc_in_w[2], enc_in_b[2], enc_out_w[2], enc_out_b[2], enc_l1_w[2], enc_l1_b[2], enc_l2_w[2], enc_l2_b[2];
  int32_t enc_n1_w[2], enc_n1_b[2], enc_n2_w[2], enc_n2_b[2];
  int32_t a2t_in_w, a2t_in_b, a2t_out_w, a2t_out_b, t2a_in_w, t2a_in_b, t2a_out_w, t2a_out_b;
  int32_t head1_w, head1_b, head2_w, head2_b, head3_w, head3_b;
} muav_attcoal_offsets;
int muav_att_coalition_scores(const float* d_params, const muav_attcoal_offsets* offsets, const float* d_task_feats,
                              const uint8_t* d_task_mask, const float* d_agent_feats, const uint8_t* d_agent_mask,
                              const float* d_edge_valid, const int32_t* d_env_idx, const uint8_t* d_need, int n,
                              int max_tasks, int max_agents, float* d_scores, void* stream);

/* The same Att-Pair / Att-ContextPair forward on the tensor cores (tcgen05.mma kind::tf32 in 3xTF32, accumulators and
 * activations in TMEM; csrc/muav_scorer_tc.cu).  Same function, arguments and tolerance as muav_att_context_pair_scores
 * (d_context NULL for AttPairNet); d_tc_weights: muav_att_pair_tc_floats() floats written once per parameter set by
 * muav_att_pair_tc_pack (the linear layers' weights split into TF32 hi / lo planes in the MMA's shared-memory layout). */
int64_t muav_att_pair_tc_floats(void);
int muav_att_pair_tc_pack(const float* d_params, const muav_attpair_offsets* offsets, float* d_tc_weights, void* stream);
int muav_att_pair_scores_tc(const float* d_params, const muav_attpair_offsets* offsets, const float* d_tc_weights,
                            const float* d_task_feats, const uint8_t* d_task_mask, const float* d_agent_feats,
                            const uint8_t* d_agent_mask, const float* d_edge_valid, const float* d_context,
                            const int32_t* d_env_idx, const uint8_t* d_need, int n, int max_tasks, int max_agents,
                            float score_clamp, float* d_scores, void* stream);

/* Fused AttCommitNet forward (TaskAllocation/Hybrid/AttentionCommit.py:68-100; AttentionCommit.act without exploration,
 * :167-175): commit tokens (muav_tokens_commit layout, agent features [.,13]) -> priorities f32 [E, max_tasks] and commit
 * gates f32 [E, max_agents] (0 on padded columns / rows), the d_plan_pri / d_plan_commit inputs of planner 3.  d_model 64,
 * 4 heads, 2 encoder layers, feed-forward 128 are fixed; weights packed like muav_attpair_offsets (transposed, offsets in
 * floats).  d_env_idx / d_need as in muav_att_pair_scores. */
typedef struct muav_attcommit_offsets {
  int32_t agent_proj_w, agent_proj_b, task_proj_w, task_proj_b, type_embed;
  int32_t enc_in_w[2], enc_in_b[2], enc_out_w[2], enc_out_b[2], enc_l1_w[2], enc_l1_b[2], enc_l2_w[2], enc_l2_b[2];
  int32_t enc_n1_w[2], enc_n1_b[2], enc_n2_w[2], enc_n2_b[2];
  int32_t priority_w, priority_b, commit_w, commit_b;
} muav_attcommit_offsets;
int muav_att_commit_vectors(const float* d_params, const muav_attcommit_offsets* offsets, const float* d_task_feats,
                            const uint8_t* d_task_mask, const float* d_agent_feats13, const uint8_t* d_agent_mask,
                            const int32_t* d_env_idx, const uint8_t* d_need, int n, int max_tasks, int max_agents,
                            float* d_priorities, float* d_commits, void* stream);
/* AttCommitNet forward on the tensor cores (same machinery as muav_att_pair_scores_tc: two encoder layers with
 * activations in TMEM, then the two sigmoid heads); same function, arguments and tolerance as muav_att_commit_vectors. */
int64_t muav_att_commit_tc_floats(void);
int muav_att_commit_tc_pack(const float* d_params, const muav_attcommit_offsets* offsets, float* d_tc_weights, void* stream);
int muav_att_commit_vectors_tc(const float* d_params, const muav_attcommit_offsets* offsets, const float* d_tc_weights,
                               const float* d_task_feats, const uint8_t* d_task_mask, const float* d_agent_feats13,
                               const uint8_t* d_agent_mask, const int32_t* d_env_idx, const uint8_t* d_need, int n,
                               int max_tasks, int max_agents, float* d_priorities, float* d_commits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MUAV_H_ */
