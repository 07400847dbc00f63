// Read-only views of an environment record: pair tokens for the learned scorers and the
// observation tensors.
//   tokens_pair_env  -> build_pair_tokens (TaskAllocation/Hybrid/PairCostHybrid.py:31-65) over
//                       build_att_tokens  (TaskAllocation/Hybrid/AttentionRAH.py:50-173, raw=False)
//   observe_env      -> _generate_observations / get_task_info / _event_flag_vector
//                       (mUAV_TA/DroneEnv.py:365-492)
// Values are computed in float64 exactly as the reference's Python floats and narrowed to
// float32 only where the reference stores float32 (np.float32 token matrices, event flags).
#pragma once
#include "muav_core.cuh"

namespace muav {

MUAV_HD inline double urgency_of(const View& V, int k, int t) {
  int dl = V.k_deadline()[k];
  if (dl < 0) return 0.0;
  int rem = dl - t;
  if (rem < 0) rem = 0;
  return 1.0 - dmin((double)rem / 40.0, 1.0);
}

MUAV_HD inline bool view_known(const View& V, int a, int k) {
  return (V.known()[(k >> 5) * V.lay().D.A + a] >> (k & 31)) & 1u;
}

// task_feats [max_tasks,13] f32, task_mask [max_tasks] u8 (1 = padding), agent_feats [max_agents,12] f32,
// agent_mask [max_agents] u8, edge_valid [max_agents,max_tasks] f32, task_ids [max_tasks] i32.
// Work is spread over (lane, nlanes): lane 0 builds the ordered token-task list in `cols`
// (int16 scratch, >= max_tasks entries + 2), then columns and agent rows are independent.
// af_dim = 12 (Att-Pair / Att-RAH agent features) or 13: enrich_commit_tokens (AttentionCommit.py:49-62) appends the
// remaining commit-lock fraction min(max(commit_until - t, 0) / max(commit_horizon or 25, 1), 1); ev may be null.
// raw != 0: per-entity attributes only (build_att_tokens(raw=True), AttentionRAH.py:86-97,140-146): task_feats [.,9],
// agent_feats [.,11].  ctx != NULL: build_context_summary (ContextPairHybrid.py:33-70), 8 floats (1 when raw).
MUAV_HD inline void tokens_pair_env(const View& V, const muav_config& C, int max_tasks, int max_agents, float* tf,
                                    uint8_t* tm, float* af, uint8_t* am, float* ev, int32_t* ids, int16_t* cols,
                                    int lane, int nlanes, int af_dim = 12, int raw = 0, float* ctx = nullptr) {
  const int TD = raw ? 9 : 13;
  if (raw) af_dim = 11;
  const int A = V.lay().D.A, TC = V.lay().D.TC;
  const int n = V.hi()[HI_N_TASKS];
  const int t = V.hi()[HI_T];
  const double mc = C.max_coord;
  const int horizon = C.max_time_steps > 1 ? C.max_time_steps : 1;
  const double mid_x = C.area_w * 0.5;
  const bool vis_none = !(C.sense_radius != 0.0) && !(C.threat_delay != 0);
  const double urgent_thr = 1.0 - 12.0 / 40.0;
#if defined(__CUDA_ARCH__)
  // ordered token-task list: lane <-> task id, compacted with ballots (the order is the id order either way)
  {
    int n_open_res = 0;
    for (int base = 0; base < n; base += 32) {
      const int k = base + lane;
      bool res = k < n && V.k_status()[k] != 2;
      if (res) {
        const int ti = V.k_type()[k];
        res = V.k_alloc_ti()[k] < V.k_cur_ti()[k];  // AttentionRAH.py:67-71
      }
      const unsigned m = __ballot_sync(0xffffffffu, res);
      const int pos = n_open_res + __popc(m & ((1u << lane) - 1u));
      if (res && pos < max_tasks) cols[pos] = (int16_t)k;
      n_open_res += __popc(m);
    }
    if (lane == 0) {
      cols[max_tasks] = (int16_t)(n_open_res < max_tasks ? n_open_res : max_tasks);
      cols[max_tasks + 1] = (int16_t)n_open_res;
    }
  }
#else
  if (lane == 0) {
    int n_open_all = 0, col = 0;
    for (int k = 0; k < n; ++k) {
      if (V.k_status()[k] == 2) continue;
      int ti = V.k_type()[k];
      if (!(V.k_alloc_ti()[k] < V.k_cur_ti()[k])) continue;  // AttentionRAH.py:67-71
      ++n_open_all;
      if (col < max_tasks) cols[col++] = (int16_t)k;
    }
    cols[max_tasks] = (int16_t)col;
    cols[max_tasks + 1] = (int16_t)n_open_all;
  }
#endif
  MUAV_WARP_SYNC();
  const int ncol = cols[max_tasks];
  const int n_open_all = cols[max_tasks + 1];
  int n_live = 0;
  for (int a = 0; a < A; ++a) n_live += V.a_state()[a] != -1;
  const int n_agents = n_live > 1 ? n_live : 1;
  // ---- context summary (ContextPairHybrid.py:33-70)
  if (ctx && lane == 0) {
    const double tfrac = (double)t / (double)horizon;
    if (raw) {
      ctx[0] = (float)tfrac;
    } else {
      int n_urgent = 0, left = 0, right = 0, free_a = 0, fighters = 0;
      for (int j = 0; j < ncol; ++j) {
        const int k = cols[j];
        if (urgency_of(V, k, t) >= urgent_thr && V.k_deadline()[k] >= 0) ++n_urgent;
        if (V.k_posx()[k] < mid_x) ++left;
        else ++right;
      }
      for (int a = 0; a < A; ++a) {
        if (V.a_state()[a] == -1) continue;
        if (V.a_qlen()[a] == 0) ++free_a;
        if (is_fighter(V.a_type()[a])) ++fighters;
      }
      const double nt = (double)(ncol > 1 ? ncol : 1), na = (double)n_agents;
      const int diff = left > right ? left - right : right - left;
      ctx[0] = (float)((double)n_urgent / nt);
      ctx[1] = (float)(dmin((double)ncol / na, 4.0) / 4.0);
      ctx[2] = (float)((double)free_a / na);
      ctx[3] = (float)((double)fighters / na);
      ctx[4] = (float)((double)left / nt);
      ctx[5] = (float)((double)right / nt);
      ctx[6] = (float)((double)diff / nt);
      ctx[7] = (float)tfrac;
    }
  }
  // ---- task columns
  for (int j = lane; j < max_tasks; j += nlanes) {
    float* f = tf + j * TD;
    if (j >= ncol) {
      tm[j] = 1;
      ids[j] = 0;
      for (int c = 0; c < TD; ++c) f[c] = 0.0f;
      continue;
    }
    const int k = cols[j];
    const int ti = V.k_type()[k];
    const double cur = V.k_cur_ti()[k], al = V.k_alloc_ti()[k];
    double urg = urgency_of(V, k, t);
    int n_know_i = 0;
    for (int a = 0; a < A; ++a) n_know_i += view_known(V, a, k) ? 1 : 0;
    double scar = 0.0;
    if (!vis_none) scar = 1.0 - dmin((double)n_know_i / (double)n_agents, 1.0);
    double rem = dmax(cur - al, 0.0);
    double is_dyn = V.k_deadline()[k] >= 0 ? 1.0 : 0.0;
    double n_know = vis_none ? 1.0 : (double)n_know_i;
    double d_spec = mc;
    bool any_spec = false;
    for (int a = 0; a < A; ++a) {
      if (V.a_state()[a] == -1 || V.a_type()[a] != UT_F2) continue;
      double d = norm2(V.a_posx()[a] - V.k_posx()[k], V.a_posy()[a] - V.k_posy()[k]);
      if (!any_spec || d < d_spec) d_spec = d;
      any_spec = true;
    }
    f[0] = (float)(V.k_posx()[k] / mc);
    f[1] = (float)(V.k_posy()[k] / mc);
    f[2] = (float)((double)ti / 8.0);
    f[3] = ti == TT_ATT ? 1.0f : 0.0f;
    f[4] = ti == TT_REC ? 1.0f : 0.0f;
    f[5] = ti == TT_INT ? 1.0f : 0.0f;
    if (raw) {
      const int dl = V.k_deadline()[k];
      int remt = dl - t;
      if (remt < 0) remt = 0;
      f[6] = dl < 0 ? 1.0f : (float)dmin((double)remt / (double)horizon, 1.0);
      f[7] = (float)dmin(rem / 4.0, 1.0);
      f[8] = (float)is_dyn;
    } else {
      f[6] = (float)urg;
      f[7] = (float)scar;
      f[8] = (float)dmin(rem / 4.0, 1.0);
      f[9] = (float)is_dyn;
      f[10] = (float)dmin(n_know / (double)n_agents, 1.0);
      f[11] = (float)dmin(d_spec / mc, 1.0);
      f[12] = V.k_posx()[k] < mid_x ? 0.0f : 1.0f;
    }
    tm[j] = 0;
    ids[j] = k + 1;
  }
  // ---- agent rows (row i = i-th live agent)
#if defined(__CUDA_ARCH__)
  const bool warp_rows = max_agents <= 32;   // the warp-cooperative forms below put row i on lane i
#else
  const bool warp_rows = false;
#endif
  for (int i = lane; i < (warp_rows ? 32 : max_agents); i += nlanes) {
    float* f = af + i * af_dim;
    float* evr = (ev && !warp_rows) ? ev + i * max_tasks : nullptr;
    int a = -1, seen = 0;
    for (int b = 0; b < A; ++b) {
      if (V.a_state()[b] == -1) continue;
      if (seen == i) { a = b; break; }
      ++seen;
    }
    int n_known_urgent = 0;
#if defined(__CUDA_ARCH__)
    if (warp_rows) {
      // known urgent tasks per agent: the per-task predicate is evaluated once (lane <-> task), every row then counts
      // the bits its agent knows; edge_valid is filled afterwards over (row, column) pairs with coalesced stores
      const int words = (n + 31) >> 5;
      for (int w = 0; w < words; ++w) {
        const int k = (w << 5) + lane;
        bool u = k < n && V.k_status()[k] != 2 && V.k_deadline()[k] >= 0;
        if (u) {
          const int ti = V.k_type()[k];
          u = V.k_alloc_ti()[k] < V.k_cur_ti()[k] && urgency_of(V, k, t) >= urgent_thr;
        }
        const unsigned m = __ballot_sync(0xffffffffu, u);
        if (a >= 0) n_known_urgent += __popc(vis_none ? m : (m & V.known()[w * A + a]));
      }
      const int at_l = a >= 0 ? (int)V.a_type()[a] : 0;
      const int n_edges = ev ? max_agents * max_tasks : 0;
      int er = lane / max_tasks, ej = lane - er * max_tasks;   // (row, column) of this lane's edge, advanced without dividing
      for (int base = 0; base < n_edges; base += 32) {   // warp-uniform trip count: the shuffles need every lane
        const int idx = base + lane;
        const bool in = idx < n_edges;
        const int r = in ? er : 0, j = in ? ej : 0;
        ej += 32;
        while (ej >= max_tasks) { ej -= max_tasks; ++er; }
        const int ar = __shfl_sync(0xffffffffu, a, r);
        const int atr = __shfl_sync(0xffffffffu, at_l, r);
        float v = 0.0f;
        if (in && ar >= 0 && j < ncol) {
          const int k = cols[j];
          const int el = V.k_elig()[k];
          const bool ok = (vis_none || view_known(V, ar, k)) && (el == 0 || ((el >> atr) & 1)) &&
                          V.a_caps()[V.k_type()[k] * A + ar] > 0;
          v = ok ? 1.0f : 0.0f;
        }
        if (in) ev[idx] = v;
      }
      if (i >= max_agents) continue;
    }
#endif
    if (a < 0) {
      am[i] = 1;
      for (int c = 0; c < af_dim; ++c) f[c] = 0.0f;
      if (evr)
        for (int j = 0; j < max_tasks; ++j) evr[j] = 0.0f;
      continue;
    }
    int at = V.a_type()[a];
    double cap_rec = V.a_caps()[1 * A + a], cap_att = V.a_caps()[2 * A + a], cap_def = V.a_caps()[3 * A + a];
    if (!warp_rows) {
      for (int k = 0; k < n; ++k) {
        if (V.k_status()[k] == 2) continue;
        if (V.k_deadline()[k] < 0) continue;
        int ti = V.k_type()[k];
        if (!(V.k_alloc_ti()[k] < V.k_cur_ti()[k])) continue;
        if (!vis_none && !view_known(V, a, k)) continue;
        if (urgency_of(V, k, t) >= urgent_thr) ++n_known_urgent;
      }
    }
    f[0] = (float)(V.a_posx()[a] / mc);
    f[1] = (float)(V.a_posy()[a] / mc);
    f[2] = is_fighter(at) ? 1.0f : 0.0f;
    f[3] = is_recon(at) ? 1.0f : 0.0f;
    f[4] = V.a_qlen()[a] == 0 ? 1.0f : 0.0f;
    f[5] = (float)dmin(cap_att / 2.0, 1.0);
    f[6] = (float)dmin(cap_def / 2.0, 1.0);
    f[7] = (float)dmin(cap_rec / 2.0, 1.0);
    f[8] = (float)((double)V.a_state()[a] / 5.0);
    f[9] = (float)((double)t / (double)horizon);
    if (raw) {
      f[10] = at == UT_F2 ? 1.0f : 0.0f;
    } else {
      f[10] = (float)dmin((double)n_known_urgent / (double)(n_open_all > 1 ? n_open_all : 1), 1.0);
      f[11] = at == UT_F2 ? 1.0f : 0.0f;
    }
    if (!raw && af_dim > 12) {
      const int hz = C.commit_horizon != 0 ? C.commit_horizon : 25;
      const double rem = dmax((double)V.a_commit()[a] - (double)t, 0.0);
      f[12] = (float)dmin(rem / (double)(hz > 1 ? hz : 1), 1.0);
    }
    am[i] = 0;
    // edge_valid (PairCostHybrid.py:41-60)
    for (int j = 0; evr && j < max_tasks; ++j) {
      float v = 0.0f;
      if (j < ncol) {
        int k = cols[j];
        int el = V.k_elig()[k];
        bool ok = (vis_none || view_known(V, a, k)) && (el == 0 || ((el >> at) & 1)) &&
                  V.a_caps()[V.k_type()[k] * A + a] > 0;
        v = ok ? 1.0f : 0.0f;
      }
      evr[j] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// build_escort_tokens (TaskAllocation/Hybrid/AttentionEscort.py:76-241).
struct EscortTokScratch {
  double* keys;    // [TC] priority key of each candidate
  int16_t* cand;   // [TC] non-closed task indices in id order
  int16_t* sel;    // [TC] candidates that enter the sort (indices into cand)
  uint8_t* flag;   // [TC] bit 0: residual-open (_open_tasks_residual), bit 1: known to a live agent
  int16_t* cols;   // [max_tasks + 4] kept task index per column; [max_tasks] = #columns, [max_tasks+1] = #candidates,
                   //                 [max_tasks+2] = #selected
};

MUAV_HD inline size_t escort_tok_scratch_bytes(int TC, int max_tasks) {
  size_t b = (size_t)TC * 13 + (size_t)(max_tasks + 4) * 2;
  return (b + 15) & ~(size_t)15;
}
MUAV_HD inline EscortTokScratch carve_escort_tok(char* p, int TC, int max_tasks) {
  EscortTokScratch W;
  W.keys = (double*)p;
  W.cand = (int16_t*)(p + (size_t)TC * 8);
  W.sel = W.cand + TC;
  W.cols = W.sel + TC;
  W.flag = (uint8_t*)(W.cols + max_tasks + 4);
  return W;
}

MUAV_HD inline bool view_in_queue(const View& V, int a, int tid) {
  const int A = V.lay().D.A;
  const int ql = V.a_qlen()[a];
  for (int q = 0; q < ql; ++q)
    if (V.a_queue()[q * A + a] == tid) return true;
  return false;
}

// _threat_stats (AttentionEscort.py:46-65)
MUAV_HD inline void view_threat_stats(const View& V, const muav_config& C, int k, double* pressure, double* dist_n,
                                      double* fighter_pressure) {
  const double mc = C.max_coord;
  double ax = V.k_posx()[k], ay = V.k_posy()[k];
  const int prot = V.k_prot_agent()[k];
  if (prot >= 0) { ax = V.a_posx()[prot]; ay = V.a_posy()[prot]; }
  double best = mc;
  int n_near = 0;
  const int na = V.hi()[HI_N_ACTIVE];
  for (int i = 0; i < na; ++i) {
    const int hid = V.h_order()[i];
    if (V.h_status()[hid] == 2) continue;
    const double d = norm2(V.h_posx()[hid] - ax, V.h_posy()[hid] - ay);
    best = dmin(best, d);
    if (d < 150.0) ++n_near;
  }
  *pressure = 1.0 - dmin(best / mc, 1.0);
  *dist_n = dmin(best / mc, 1.0);
  *fighter_pressure = dmin((double)n_near / 4.0, 1.0);
}

MUAV_HD inline void tokens_escort_env(const View& V, const muav_config& C, int max_tasks, int max_agents, float* tf,
                                      uint8_t* tm, float* af, uint8_t* am, float* ev, int32_t* ids, int32_t* order,
                                      EscortTokScratch W, int lane, int nlanes) {
  const int A = V.lay().D.A, TC = V.lay().D.TC, IC = V.lay().D.IC;
  const int n = V.hi()[HI_N_TASKS];
  const int t = V.hi()[HI_T];
  const double mc = C.max_coord;
  const double mid_x = C.area_w * 0.5;
  const bool vis_none = !(C.sense_radius != 0.0) && !(C.threat_delay != 0);
  const double urgent_thr = 1.0 - 12.0 / 40.0;
  int16_t* cols = W.cols;
  if (lane == 0) {
    int nc = 0;
    for (int k = 0; k < n && nc < TC; ++k)
      if (V.k_status()[k] != 2) W.cand[nc++] = (int16_t)k;
    cols[max_tasks + 1] = (int16_t)nc;
  }
  MUAV_WARP_SYNC();
  const int ncand = cols[max_tasks + 1];
  // ---- per candidate: residual-open, known to a live agent, priority key
  for (int c = lane; c < ncand; c += nlanes) {
    const int k = W.cand[c];
    const int ti = V.k_type()[k];
    uint8_t fl = 0;
    if (V.k_kind()[k] == 1 || V.k_req_agents()[k] > 0) {
      const int ra = V.k_req_agents()[k];
      int cnt = 0;
      for (int a = 0; a < A; ++a) cnt += view_in_queue(V, a, k + 1) ? 1 : 0;
      if ((double)(ra != 0 ? ra : 1) - (double)cnt > 0.0) fl = 1;
    } else if (V.k_alloc_ti()[k] < V.k_cur_ti()[k]) {
      fl = 1;
    }
    double key = 0.0;
    if (fl) {
      for (int a = 0; a < A; ++a)
        if (V.a_state()[a] != -1 && view_known(V, a, k)) { fl |= 2; break; }
      double pressure, dn, fp;
      view_threat_stats(V, C, k, &pressure, &dn, &fp);
      const double urg = urgency_of(V, k, t);
      const double is_escort = V.k_kind()[k] == 1 ? 1.0 : 0.0;
      const double is_int = ti == TT_INT ? 1.0 : 0.0;
      key = -(1.5 * urg + 1.2 * pressure + 0.8 * is_escort + 0.5 * is_int);
    }
    W.flag[c] = fl;
    W.keys[c] = key;
  }
  MUAV_WARP_SYNC();
  if (lane == 0) {
    bool any_known = false;
    for (int c = 0; c < ncand; ++c)
      if ((W.flag[c] & 3) == 3) { any_known = true; break; }
    const bool filter = !vis_none && any_known;
    int m = 0;
    for (int c = 0; c < ncand; ++c) {
      if (!(W.flag[c] & 1)) continue;
      if (filter && !(W.flag[c] & 2)) continue;
      W.sel[m++] = (int16_t)c;
    }
    cols[max_tasks + 2] = (int16_t)m;
    cols[max_tasks] = (int16_t)(m < max_tasks ? m : max_tasks);
  }
  MUAV_WARP_SYNC();
  const int m = cols[max_tasks + 2];
  const int ncol = cols[max_tasks];
  // ---- stable sort by key: rank = number of entries that come first
  for (int i = lane; i < m; i += nlanes) {
    const double ki = W.keys[W.sel[i]];
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const double kj = W.keys[W.sel[j]];
      rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
    }
    if (rank < max_tasks) cols[rank] = W.cand[W.sel[i]];
  }
  MUAV_WARP_SYNC();
  int n_live = 0;
  for (int a = 0; a < A; ++a) n_live += V.a_state()[a] != -1;
  const int n_agents = n_live > 1 ? n_live : 1;
  if (order)
    for (int j = lane; j < IC; j += nlanes) order[j] = j < ncol ? (int32_t)cols[j] : -1;
  // ---- task columns
  for (int j = lane; j < max_tasks; j += nlanes) {
    float* f = tf + j * 22;
    if (j >= ncol) {
      tm[j] = 1;
      ids[j] = 0;
      for (int c = 0; c < 22; ++c) f[c] = 0.0f;
      continue;
    }
    const int k = cols[j];
    const int ti = V.k_type()[k];
    const double urg = urgency_of(V, k, t);
    int n_know_i = 0;
    for (int a = 0; a < A; ++a) n_know_i += view_known(V, a, k) ? 1 : 0;
    double scar = 0.0, n_know = 0.0;
    if (!vis_none) {
      scar = 1.0 - dmin((double)n_know_i / (double)n_agents, 1.0);
      n_know = (double)n_know_i;
    }
    double rem, req_agents;
    if (V.k_kind()[k] == 1 || V.k_req_agents()[k] > 0) {
      const int ra = V.k_req_agents()[k];
      int cnt = 0;
      for (int a = 0; a < A; ++a) cnt += view_in_queue(V, a, k + 1) ? 1 : 0;
      req_agents = (double)(ra != 0 ? ra : 1);
      rem = dmax(req_agents - (double)cnt, 0.0);
    } else {
      rem = dmax(V.k_cur_ti()[k] - V.k_alloc_ti()[k], 0.0);
      req_agents = 1.0;
    }
    double d_spec = mc;
    bool any_spec = false;
    for (int a = 0; a < A; ++a) {
      if (V.a_state()[a] == -1 || V.a_type()[a] != UT_F2) continue;
      const double d = norm2(V.a_posx()[a] - V.k_posx()[k], V.a_posy()[a] - V.k_posy()[k]);
      if (!any_spec || d < d_spec) d_spec = d;
      any_spec = true;
    }
    const double deficit = dmin(rem / 4.0, 1.0);
    double pressure, threat_dist, fighter_pressure;
    view_threat_stats(V, C, k, &pressure, &threat_dist, &fighter_pressure);
    const int prot = V.k_prot_agent()[k];
    double prot_x = V.k_posx()[k] / mc, prot_y = V.k_posy()[k] / mc, prot_alive = 0.0;
    if (prot >= 0) {
      prot_x = V.a_posx()[prot] / mc;
      prot_y = V.a_posy()[prot] / mc;
      prot_alive = V.a_state()[prot] == -1 ? 0.0 : 1.0;
    }
    f[0] = (float)(V.k_posx()[k] / mc);
    f[1] = (float)(V.k_posy()[k] / mc);
    f[2] = (float)((double)ti / 8.0);
    f[3] = ti == TT_ATT ? 1.0f : 0.0f;
    f[4] = ti == TT_REC ? 1.0f : 0.0f;
    f[5] = ti == TT_INT ? 1.0f : 0.0f;
    f[6] = (float)urg;
    f[7] = (float)scar;
    f[8] = (float)deficit;
    f[9] = V.k_deadline()[k] >= 0 ? 1.0f : 0.0f;
    f[10] = (float)dmin(n_know / (double)n_agents, 1.0);
    f[11] = (float)dmin(d_spec / mc, 1.0);
    f[12] = V.k_posx()[k] < mid_x ? 0.0f : 1.0f;
    f[13] = V.k_kind()[k] == 1 ? 1.0f : 0.0f;
    f[14] = (float)deficit;
    f[15] = (float)pressure;
    f[16] = (float)prot_x;
    f[17] = (float)prot_y;
    f[18] = (float)dmin(req_agents / 4.0, 1.0);
    f[19] = (float)threat_dist;
    f[20] = (float)prot_alive;
    f[21] = (float)fighter_pressure;
    tm[j] = 0;
    ids[j] = k + 1;
  }
  // ---- agent rows (row i = i-th live agent)
  const int hz0 = C.commit_horizon != 0 ? C.commit_horizon : 20;
  const double horizon = (double)(hz0 > 1 ? hz0 : 1);
  const int mt_steps = C.max_time_steps > 1 ? C.max_time_steps : 1;
  for (int i = lane; i < max_agents; i += nlanes) {
    float* f = af + i * 16;
    float* evr = ev + i * max_tasks;
    int a = -1, seen = 0;
    for (int b = 0; b < A; ++b) {
      if (V.a_state()[b] == -1) continue;
      if (seen == i) { a = b; break; }
      ++seen;
    }
    if (a < 0) {
      am[i] = 1;
      for (int c = 0; c < 16; ++c) f[c] = 0.0f;
      for (int j = 0; j < max_tasks; ++j) evr[j] = 0.0f;
      continue;
    }
    const int at = V.a_type()[a];
    const double cap_rec = V.a_caps()[1 * A + a], cap_att = V.a_caps()[2 * A + a], cap_def = V.a_caps()[3 * A + a];
    int n_known_urgent = 0;
    for (int c = 0; c < ncand; ++c) {
      if (!(W.flag[c] & 1)) continue;
      const int k = W.cand[c];
      if (V.k_deadline()[k] < 0) continue;
      if (!vis_none && !view_known(V, a, k)) continue;
      if (urgency_of(V, k, t) >= urgent_thr) ++n_known_urgent;
    }
    int n_known_tasks = 0;
    if (!vis_none) {
      const int KWn = (n + 31) >> 5;
      for (int w = 0; w < KWn; ++w) {
        uint32_t bits = V.known()[w * A + a];
        if (w == KWn - 1 && (n & 31)) bits &= (1u << (n & 31)) - 1u;
        while (bits) { bits &= bits - 1; ++n_known_tasks; }
      }
    }
    double is_escorting = 0.0, dist_prot = 1.0, near_escort = 0.0;
    if (V.a_qlen()[a] > 0) {
      const int hk = V.a_queue()[a] - 1;
      if (hk >= 0 && V.k_kind()[hk] == 1) {
        is_escorting = 1.0;
        const int prot = V.k_prot_agent()[hk];
        if (prot >= 0) {
          dist_prot = dmin(norm2(V.a_posx()[a] - V.a_posx()[prot], V.a_posy()[a] - V.a_posy()[prot]) / mc, 1.0);
          near_escort = 1.0 - dist_prot;
        }
      }
    }
    const double rem_commit = dmax((double)V.a_commit()[a] - (double)t, 0.0);
    f[0] = (float)(V.a_posx()[a] / mc);
    f[1] = (float)(V.a_posy()[a] / mc);
    f[2] = is_fighter(at) ? 1.0f : 0.0f;
    f[3] = is_recon(at) ? 1.0f : 0.0f;
    f[4] = V.a_qlen()[a] == 0 ? 1.0f : 0.0f;
    f[5] = (float)dmin(cap_att / 2.0, 1.0);
    f[6] = (float)dmin(cap_def / 2.0, 1.0);
    f[7] = (float)dmin(cap_rec / 2.0, 1.0);
    f[8] = (float)((double)V.a_state()[a] / 5.0);
    f[9] = (float)((double)t / (double)mt_steps);
    f[10] = (float)dmin((double)n_known_urgent / 8.0, 1.0);
    f[11] = at == UT_F2 ? 1.0f : 0.0f;
    f[12] = (float)is_escorting;
    f[13] = (float)dist_prot;
    f[14] = (float)dmin(rem_commit / horizon, 1.0);
    f[15] = (float)dmin(near_escort + (double)n_known_tasks / 16.0, 1.0);
    am[i] = 0;
    for (int j = 0; j < max_tasks; ++j) {
      float v = 0.0f;
      if (j < ncol) {
        const int k = cols[j];
        const int el = V.k_elig()[k];
        v = ((vis_none || view_known(V, a, k)) && (el == 0 || ((el >> at) & 1))) ? 1.0f : 0.0f;
      }
      evr[j] = v;
    }
  }
}

#define MUAV_OBS_TASK_DIM_ 21
// tasks_info [max_rows, 21] f64: id, x, y, status, cur[6], alloc[6], init_time, end_time, type_idx, unmet, age
// (status = -1 marks padding); pad_mask [max_rows]; legal_mask [A, max_rows]; agent_obs [A, 9]; event_flags [5] f32.
MUAV_HD inline void observe_env(const View& V, const muav_config& C, int max_rows, double* ti_out, uint8_t* pad,
                                uint8_t* legal, double* ao, float* ef, int32_t* n_rows_out) {
  const int A = V.lay().D.A, TC = V.lay().D.TC;
  const int n = V.hi()[HI_N_TASKS];
  const int t = V.hi()[HI_T];
  const double mc = C.max_coord;
  const double mt = (double)(C.max_time_steps > 1 ? C.max_time_steps : 1);
  for (int r = 0; r < max_rows; ++r) {
    pad[r] = 0;
    for (int c = 0; c < MUAV_OBS_TASK_DIM_; ++c) ti_out[r * MUAV_OBS_TASK_DIM_ + c] = 0.0;
    ti_out[r * MUAV_OBS_TASK_DIM_ + 3] = -1.0;
    for (int a = 0; a < A; ++a) legal[a * max_rows + r] = 0;
  }
  int rows = 0, n_open = 0;
  for (int k = 0; k < n; ++k) {
    if (V.k_status()[k] == 2) continue;
    ++n_open;
    if (rows >= max_rows) continue;
    double* o = ti_out + rows * MUAV_OBS_TASK_DIM_;
    int tt = V.k_type()[k];
    o[0] = (double)(k + 1);
    o[1] = V.k_posx()[k] / mc;
    o[2] = V.k_posy()[k] / mc;
    o[3] = (double)V.k_status()[k];
    for (int c = 0; c < 6; ++c) {
      o[4 + c] = V.k_cur2(c, k);
      o[10 + c] = V.k_alloc2(c, k);
    }
    o[16] = (V.k_init()[k] - (double)t) / mt;
    o[17] = (V.k_dtime()[k] - (double)t) / mt;
    o[18] = (double)tt / 6.0;
    double unmet = dmax(V.k_cur2(tt, k) - V.k_alloc2(tt, k), 0.0);
    o[19] = unmet / dmax(V.k_org_ti()[k], 1e-6);
    o[20] = dmin(((double)t - (double)V.k_created()[k]) / mt, 1.0);
    pad[rows] = 1;
    ++rows;
  }
  if (n_open == 0) {
    // single idle row (DroneEnv.py:387-396)
    ti_out[3] = 0.0;
    pad[0] = 1;
    for (int a = 0; a < A; ++a) {
      int head = V.a_qlen()[a] > 0 ? V.a_queue()[a] : 0;
      legal[a * max_rows] = (V.a_state()[a] == 2) ? (head == 0 ? 1 : 0) : 1;
    }
    rows = 1;
  } else {
    Sim S;
    S.V = V;
    S.Cp = &C;
    for (int a = 0; a < A; ++a) {
      uint8_t* lm = legal + a * max_rows;
      int head = V.a_qlen()[a] > 0 ? V.a_queue()[a] : 0;
      bool any = false;
      for (int r = 0; r < rows; ++r) {
        int tid = (int)ti_out[r * MUAV_OBS_TASK_DIM_];
        lm[r] = S.is_valid(a, tid) ? 1 : 0;
        any = any || lm[r];
      }
      if (!any) {
        int hit = -1;
        for (int r = 0; r < rows; ++r)
          if ((int)ti_out[r * MUAV_OBS_TASK_DIM_] == head) { hit = r; break; }
        lm[hit >= 0 ? hit : 0] = 1;
      }
      if (V.a_state()[a] == 2) {
        for (int r = 0; r < rows; ++r) lm[r] = ((int)ti_out[r * MUAV_OBS_TASK_DIM_] == head) ? 1 : 0;
      }
    }
  }
  *n_rows_out = n_open == 0 ? 1 : n_open;
  for (int a = 0; a < A; ++a) {
    double* o = ao + a * 9;
    o[0] = V.a_posx()[a] / mc;
    o[1] = V.a_posy()[a] / mc;
    for (int c = 0; c < 6; ++c) o[2 + c] = V.a_caps()[c * A + a];
    o[8] = (double)(V.a_qlen()[a] > 0 ? V.a_queue()[a] : 0);
  }
  float fail = 0.0f, thr = 0.0f, rst = 0.0f;
  int nev = V.hi()[HI_N_EVENTS];
  for (int i = 0; i < nev; ++i) {
    int tag = V.events()[i] & 0xff;
    if (tag == EV_FAIL) fail = 1.0f;
    else if (tag == EV_THREAT) thr = 1.0f;
    else if (tag == EV_RESET) rst = 1.0f;
  }
  ef[0] = fail;
  ef[1] = thr;
  ef[2] = rst;
  ef[3] = (float)((double)t / mt);
  ef[4] = (float)((double)n_open / (double)(C.max_tasks > 1 ? C.max_tasks : 1));
}

// calculate_metrics / compute_s_wps / compute_s_esc (mUAV_TA/DroneEnv.py:1231-1337,2002-2011);
// o[0..29] in the order of muav_metric_name()
MUAV_HD inline void metrics_env(const View& V, const muav_config& cfg, double* o) {
  const Layout& L = V.lay();
  const int A = L.D.A;
  const int T = HIv(N_TASKS);
  double td = HFv(TOTAL_DIST);
  // compute_s_wps (DroneEnv.py:1321-1337)
  double dist_term = 0.01 * td / dmax(cfg.max_coord, 1.0);
  double rematch = cfg.reassign_penalty * (double)HIv(N_SWITCH);
  double s_wps = 12.0 * (double)HIv(N_ON_TIME) - 30.0 * (double)HIv(N_MISSED) - dist_term - rematch;
  int req = HIv(ESC_REQ_STEPS);
  double cov = (double)HIv(ESC_COV_STEPS) / (double)(req > 1 ? req : 1);
  double s_esc = s_wps + 20.0 * (double)HIv(PROT_REC_DONE) - 30.0 * (double)HIv(RECON_LOSSES) + 20.0 * cov;
  int losses = 0;
  for (int a = 0; a < A; ++a) losses += V.a_state()[a] == -1;
  int kills = 0;
  for (int i = 0; i < HIv(N_ACTIVE); ++i) kills += V.h_status()[V.h_order()[i]] == 2;
  double fq = 0.0;  // final_quality is -1 (counted as 0) or 0.0 (DroneEnv.py:1249,1566)
  o[0] = 1.0 / (double)HIv(CONCLUSION) * (double)cfg.max_time_steps;
  o[1] = td > 0 ? 1.0 / td * cfg.max_coord : 0.0;
  o[2] = T > 0 ? fq / (double)T : nan("");
  o[3] = HFv(F_REWARD);
  o[4] = s_wps;
  o[5] = s_esc;
  o[6] = losses;
  o[7] = kills;
  o[8] = (double)HIv(CONCLUSION);
  o[9] = td;
  o[10] = HIv(N_REALLOC);
  o[11] = HIv(N_SWITCH);
  o[12] = HIv(N_ARRIVALS);
  o[13] = T;
  o[14] = HIv(N_REACHED);
  o[15] = HIv(N_MISSED);
  o[16] = HIv(N_ON_TIME);
  o[17] = HIv(N_WINDOWED);
  int den = HIv(N_ON_TIME) + HIv(N_MISSED);
  o[18] = (double)HIv(N_ON_TIME) / (double)(den > 1 ? den : 1);
  int den2 = HIv(T) * (A > 1 ? A : 1);
  o[19] = (double)HIv(IDLE_RESERVE) / (double)(den2 > 1 ? den2 : 1);
  o[20] = cov;
  o[21] = HIv(PROT_REC_DONE);
  o[22] = HIv(RECON_LOSSES);
  o[23] = HIv(ESCORT_LOSSES);
  o[24] = HIv(INTERCEPTED);
  o[25] = HIv(MUTUAL);
  o[26] = HIv(BREACHES);
  o[27] = HIv(ESC_REQUESTS);
  o[28] = HIv(ESC_COMPLETED);
  o[29] = HIv(ESC_FAILED);
}

}  // namespace muav
