// TEST INFRASTRUCTURE -- CPU compile of the CUDA simulation core (muav_core.cuh / muav_alloc.cuh are
// __host__ __device__) so that the kernel's logic can be diffed against the oracle and the golden
// fixtures in the GPU-less authoring container.  Built by tests/hostcheck/build.py into
// tests/hostcheck/_build/libmuav_hostcheck.so and loaded ONLY by tests (never by the package,
// bench.py or smoke()): the product path has no CPU fallback.
// g++ -O2 -ffp-contract=off: no FMA contraction, same rounding as nvcc -fmad=false.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../multi_uav_ta_gym_env_b200/csrc/muav_alloc.cuh"
#include "../../multi_uav_ta_gym_env_b200/csrc/muav_views.cuh"

using namespace muav;

extern "C" {

#include "../../multi_uav_ta_gym_env_b200/csrc/muav_abi_common.inl"

// Same contract as muav_step but every pointer is a HOST pointer.
int hostcheck_step(const muav_config* cfg, void* records, const uint32_t* tapes, const int32_t* actions,
                   const muav_alloc_opts* opts, const muav_step_out* out, int n_envs, int n_steps) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  Layout L = make_layout(*cfg);
  muav_alloc_opts O;
  memset(&O, 0, sizeof(O));
  if (opts) O = *opts;
  muav_step_out Z;
  memset(&Z, 0, sizeof(Z));
  if (out) Z = *out;
  char* scratch = (char*)malloc((size_t)L.scratch_bytes + (size_t)cbba_scratch_bytes(L.D.A) + 64);
  int tape_stride = cfg->tape_words[0] + cfg->tape_words[1] + cfg->tape_words[2];
  const int A = L.D.A;
  int16_t act_agent[MUAV_MAX_AGENTS], act_tid[MUAV_MAX_AGENTS];
  for (int e = 0; e < n_envs; ++e) {
    Sim S;
    S.V.at((char*)records + (size_t)e * L.record_bytes);
    S.V.set_layout(&L);
    S.Cp = cfg;
    S.tape = tapes + (size_t)e * tape_stride;
    S.scratch = scratch;
    S.out_events = Z.d_events ? Z.d_events + (size_t)e * L.D.EVC : nullptr;
    S.n_out_events = 0;
    S.step_reward = 0.0;
    S.phase_cycles = nullptr;
    View& V = S.V;
    for (int s = 0; s < n_steps; ++s) {
      if (HIv(DONE)) break;
      int n_act = 0;
      if (O.mode != 0) {
        int np = plan_and_allocate(S, O, e, act_agent, act_tid, 0, 1);
        if (Z.d_n_pairs) Z.d_n_pairs[e] = np;
        if (Z.d_pairs)
          for (int i = 0; i < np; ++i) Z.d_pairs[(size_t)e * A + i] = ((int)act_agent[i] << 16) | (int)act_tid[i];
        for (int i = 0; i < np; ++i) {
          if (HIv(N_OPEN) > 0 && S.in_last_open(act_tid[i])) {
            act_agent[n_act] = act_agent[i];
            act_tid[n_act] = act_tid[i];
            ++n_act;
          }
        }
      } else if (actions) {
        const int32_t* act = actions + (size_t)e * A * 2;
        for (int i = 0; i < A; ++i) {
          int a = act[2 * i];
          if (a < 0) break;
          act_agent[n_act] = (int16_t)a;
          act_tid[n_act] = (int16_t)S.open_task_at(act[2 * i + 1]);
          ++n_act;
        }
      }
      StepResult r = S.step(act_agent, act_tid, n_act, 0, 1, true, 0);
      if (Z.d_reward) Z.d_reward[e] = r.reward;
      if (Z.d_terminated) Z.d_terminated[e] = (uint8_t)r.terminated;
      if (Z.d_truncated) Z.d_truncated[e] = (uint8_t)r.truncated;
      if (Z.d_n_events) Z.d_n_events[e] = S.n_out_events;
      if (Z.d_n_open) Z.d_n_open[e] = HIv(N_OPEN);
    }
  }
  free(scratch);
  return 0;
}

// muav_allocate with host pointers
int hostcheck_allocate(const muav_config* cfg, void* records, const muav_alloc_opts* opts, const muav_step_out* out,
                       int32_t* actions_out, int n_envs) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  Layout L = make_layout(*cfg);
  muav_step_out Z;
  memset(&Z, 0, sizeof(Z));
  if (out) Z = *out;
  char* scratch = (char*)malloc((size_t)L.scratch_bytes + (size_t)cbba_scratch_bytes(L.D.A) + 64);
  const int A = L.D.A;
  int16_t act_agent[MUAV_MAX_AGENTS], act_tid[MUAV_MAX_AGENTS];
  for (int e = 0; e < n_envs; ++e) {
    Sim S;
    S.V.at((char*)records + (size_t)e * L.record_bytes);
    S.V.set_layout(&L);
    S.Cp = cfg;
    S.tape = nullptr;
    S.scratch = scratch;
    S.out_events = nullptr;
    S.phase_cycles = nullptr;
    View& V = S.V;
    int np = HIv(DONE) ? 0 : plan_and_allocate(S, *opts, e, act_agent, act_tid, 0, 1);
    if (Z.d_n_pairs) Z.d_n_pairs[e] = np;
    int n_act = 0;
    for (int i = 0; i < np; ++i) {
      if (Z.d_pairs) Z.d_pairs[(size_t)e * A + i] = ((int)act_agent[i] << 16) | (int)act_tid[i];
      if (actions_out && HIv(N_OPEN) > 0 && S.in_last_open(act_tid[i])) {
        int k = act_tid[i] - 1, idx = 0;
        for (int kk = 0; kk < k; ++kk) idx += (V.open_mask()[kk >> 5] >> (kk & 31)) & 1u;
        actions_out[((size_t)e * A + n_act) * 2] = act_agent[i];
        actions_out[((size_t)e * A + n_act) * 2 + 1] = idx;
        ++n_act;
      }
    }
    if (actions_out && n_act < A) actions_out[((size_t)e * A + n_act) * 2] = -1;
  }
  free(scratch);
  return 0;
}

int hostcheck_tokens_pair(const muav_config* cfg, const void* records, int max_tasks, int max_agents, float* tf,
                          uint8_t* tm, float* af, uint8_t* am, float* ev, int32_t* ids, int n_envs) {
  Layout L = make_layout(*cfg);
  int16_t* cols = (int16_t*)malloc(sizeof(int16_t) * (max_tasks + 2));
  for (int e = 0; e < n_envs; ++e) {
    View V;
    V.at((char*)records + (size_t)e * L.record_bytes);
    V.set_layout(&L);
    tokens_pair_env(V, *cfg, max_tasks, max_agents, tf + (size_t)e * max_tasks * 13, tm + (size_t)e * max_tasks,
                    af + (size_t)e * max_agents * 12, am + (size_t)e * max_agents, ev + (size_t)e * max_agents * max_tasks,
                    ids + (size_t)e * max_tasks, cols, 0, 1);
  }
  free(cols);
  return 0;
}

int hostcheck_tokens_context(const muav_config* cfg, const void* records, int max_tasks, int max_agents, int raw, float* tf,
                             uint8_t* tm, float* af, uint8_t* am, float* ev, int32_t* ids, float* ctx, int n_envs) {
  Layout L = make_layout(*cfg);
  int16_t* cols = (int16_t*)malloc(sizeof(int16_t) * (max_tasks + 2));
  const int TD = raw ? 9 : 13, AD = raw ? 11 : 12, CD = raw ? 1 : 8;
  for (int e = 0; e < n_envs; ++e) {
    View V;
    V.at((char*)records + (size_t)e * L.record_bytes);
    V.set_layout(&L);
    tokens_pair_env(V, *cfg, max_tasks, max_agents, tf + (size_t)e * max_tasks * TD, tm + (size_t)e * max_tasks,
                    af + (size_t)e * max_agents * AD, am + (size_t)e * max_agents, ev + (size_t)e * max_agents * max_tasks,
                    ids + (size_t)e * max_tasks, cols, 0, 1, 12, raw, ctx + (size_t)e * CD);
  }
  free(cols);
  return 0;
}

int hostcheck_tokens_escort(const muav_config* cfg, const void* records, int max_tasks, int max_agents, float* tf,
                            uint8_t* tm, float* af, uint8_t* am, float* ev, int32_t* ids, int32_t* order, int n_envs) {
  Layout L = make_layout(*cfg);
  char* scratch = (char*)malloc(escort_tok_scratch_bytes(L.D.TC, max_tasks));
  EscortTokScratch W = carve_escort_tok(scratch, L.D.TC, max_tasks);
  for (int e = 0; e < n_envs; ++e) {
    View V;
    V.at((char*)records + (size_t)e * L.record_bytes);
    V.set_layout(&L);
    tokens_escort_env(V, *cfg, max_tasks, max_agents, tf + (size_t)e * max_tasks * 22, tm + (size_t)e * max_tasks,
                      af + (size_t)e * max_agents * 16, am + (size_t)e * max_agents, ev + (size_t)e * max_agents * max_tasks,
                      ids + (size_t)e * max_tasks, order ? order + (size_t)e * L.D.IC : nullptr, W, 0, 1);
  }
  free(scratch);
  return 0;
}

int hostcheck_observe(const muav_config* cfg, const void* records, int max_rows, double* ti, uint8_t* pad, uint8_t* legal,
                      double* ao, float* ef, int32_t* n_rows, int n_envs) {
  Layout L = make_layout(*cfg);
  for (int e = 0; e < n_envs; ++e) {
    View V;
    V.at((char*)records + (size_t)e * L.record_bytes);
    V.set_layout(&L);
    int32_t nr = 0;
    observe_env(V, *cfg, max_rows, ti + (size_t)e * max_rows * 21, pad + (size_t)e * max_rows,
                legal + (size_t)e * L.D.A * max_rows, ao + (size_t)e * L.D.A * 9, ef + (size_t)e * 5, &nr);
    if (n_rows) n_rows[e] = nr;
  }
  return 0;
}

int hostcheck_metrics(const muav_config* cfg, const void* records, double* out, int n_envs) {
  Layout L = make_layout(*cfg);
  for (int e = 0; e < n_envs; ++e) {
    View V;
    V.at((char*)records + (size_t)e * L.record_bytes);
    V.set_layout(&L);
    metrics_env(V, *cfg, out + (size_t)e * 30);
  }
  return 0;
}

int hostcheck_lsap(const double* cost, const int32_t* nr_arr, const int32_t* nc_arr, int nr_max, int nc_max,
                   int32_t* col4row, int n) {
  char* scratch = (char*)malloc((size_t)alloc_scratch_bytes(nr_max, nc_max) + 64);
  for (int b = 0; b < n; ++b) {
    int nr = nr_arr[b], nc = nc_arr[b];
    AllocScratch W = carve_scratch(scratch, nr_max, nc_max);
    const double* src = cost + (size_t)b * nr_max * nc_max;
    for (int i = 0; i < nr; ++i)
      for (int j = 0; j < nc; ++j) W.cost[i * nc + j] = src[i * nc_max + j];
    bool ok = (nr > 0 && nc > 0) ? lsap_solve(W.cost, nr, nc, W, W.col_of_row, 0, 1) : true;
    for (int i = 0; i < nr_max; ++i) col4row[(size_t)b * nr_max + i] = (ok && i < nr && nc > 0) ? W.col_of_row[i] : -1;
  }
  free(scratch);
  return 0;
}

// "<x>#..." < "<y>#..." as the reference compares its slot keys (strings): muav_alloc.cuh slot_id_less
int hostcheck_slot_id_less(int x, int y) { return slot_id_less(x, y) ? 1 : 0; }

}  // extern "C"
