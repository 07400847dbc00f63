// Local / Coalition Hungarian allocator over one environment record
// (TaskAllocation/OptimizationBased/HungarianAllocator.py:14-208) and the rectangular LSAP it
// calls (scipy.optimize.linear_sum_assignment at :181; SciPy's shortest-augmenting-path solver,
// restated with its exact scan order and tie-breaking -- see oracle/lsap.py and SURVEY.md
// Appendix B).  Cost arithmetic is float64, one rounding per operation, no contraction.
#pragma once
#include "muav_core.cuh"

namespace muav {

struct AllocScratch {
  double *cost, *u, *v, *spc, *resid;
  int16_t *path, *col4row, *row4col, *remaining, *SR, *SC, *free_agents, *round_tasks, *open_t, *tokcol, *col_of_row;
};

MUAV_HD inline AllocScratch carve_scratch(char* p, int A, int TC) {
  AllocScratch S;
  int M = A > TC ? A : TC;
  double* d = (double*)p;
  S.cost = d; d += A * TC;
  S.u = d; d += M;
  S.v = d; d += M;
  S.spc = d; d += M;
  S.resid = d; d += TC;
  int16_t* s = (int16_t*)d;
  S.path = s; s += M;
  S.col4row = s; s += M;
  S.row4col = s; s += M;
  S.remaining = s; s += M;
  S.SR = s; s += M;
  S.SC = s; s += M;
  S.free_agents = s; s += A;
  S.col_of_row = s; s += A;
  S.round_tasks = s; s += TC;
  S.open_t = s; s += TC;
  S.tokcol = s; s += TC;
  return S;
}

// Rectangular LSAP.  cost: row-major nr x nc (ld = nc).  col_of_row[i] = column of row i or -1.
// Sequential part on lane 0; the column scan is spread over the lanes with a warp arg-min that
// reproduces SciPy's tie rule: among equal minima prefer the LAST scanned unassigned column,
// otherwise the FIRST scanned one.
MUAV_HD inline bool lsap_solve(const double* cost, int nr, int nc, AllocScratch& S, int16_t* col_of_row) {
  const bool tr = nc < nr;
  const int R = tr ? nc : nr;
  const int Cn = tr ? nr : nc;
  const int si = tr ? 1 : nc;   // stride of the (possibly transposed) row index
  const int sj = tr ? nc : 1;   // stride of the (possibly transposed) column index
  double* u = S.u;
  double* v = S.v;
  double* spc = S.spc;
  int16_t* path = S.path;
  int16_t* col4row = S.col4row;
  int16_t* row4col = S.row4col;
  int16_t* remaining = S.remaining;
  int16_t* SR = S.SR;
  int16_t* SC = S.SC;
  for (int i = 0; i < R; ++i) { u[i] = 0.0; col4row[i] = -1; }
  for (int j = 0; j < Cn; ++j) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
  for (int cur = 0; cur < R; ++cur) {
    double min_val = 0.0;
    int num_remaining = Cn;
    for (int it = 0; it < Cn; ++it) remaining[it] = (int16_t)(Cn - it - 1);
    for (int i = 0; i < R; ++i) SR[i] = 0;
    for (int j = 0; j < Cn; ++j) { SC[j] = 0; spc[j] = INFINITY; }
    int i = cur;
    int sink = -1;
    while (sink == -1) {
      int index = -1;
      double lowest = INFINITY;
      SR[i] = 1;
      const double ui = u[i];
      const double* crow = cost + i * si;
      for (int it = 0; it < num_remaining; ++it) {
        int j = remaining[it];
        double r = min_val + crow[j * sj] - ui - v[j];
        if (r < spc[j]) {
          path[j] = (int16_t)i;
          spc[j] = r;
        }
        if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) {
          lowest = spc[j];
          index = it;
        }
      }
      min_val = lowest;
      if (index < 0 || min_val == INFINITY) return false;
      int j = remaining[index];
      if (row4col[j] == -1) sink = j;
      else i = row4col[j];
      SC[j] = 1;
      --num_remaining;
      remaining[index] = remaining[num_remaining];
    }
    u[cur] += min_val;
    for (int r = 0; r < R; ++r)
      if (SR[r] && r != cur) u[r] += min_val - spc[col4row[r]];
    for (int j = 0; j < Cn; ++j)
      if (SC[j]) v[j] -= min_val - spc[j];
    int j = sink;
    for (;;) {
      int r = path[j];
      row4col[j] = (int16_t)r;
      int tmp = col4row[r];
      col4row[r] = (int16_t)j;
      j = tmp;
      if (r == cur) break;
    }
  }
  for (int r = 0; r < nr; ++r) col_of_row[r] = -1;
  if (tr) {
    for (int vv = 0; vv < R; ++vv) col_of_row[col4row[vv]] = (int16_t)vv;
  } else {
    for (int r = 0; r < R; ++r) col_of_row[r] = col4row[r];
  }
  return true;
}

// is_escort / residual_demand (HungarianAllocator.py:94-111) == _task_residual (paper_eval.py:85-93)
MUAV_HD inline bool is_coalition(const Sim& S, int k) { return S.V.k_kind()[k] == 1 || S.V.k_req_agents()[k] > 0; }
MUAV_HD inline double residual_demand(const Sim& S, int k) {
  if (is_coalition(S, k)) {
    int ra = S.V.k_req_agents()[k];
    double required = (double)(ra != 0 ? ra : 1);
    double r = required - (double)S.details_count(k + 1);
    return r > 0.0 ? r : 0.0;
  }
  int ti = S.V.k_type()[k];
  int TC = S.V.L->D.TC;
  double r = S.V.k_cur()[ti * TC + k] - S.V.k_alloc()[ti * TC + k];
  return r > 0.0 ? r : 0.0;
}

// HungarianAllocator.allocate_tasks (:72-208) fused with _open_tasks (paper_eval.py:96-101) and
// _apply_assign (wps_eval.py:55-61).  Writes ordered (agent, task id) pairs; returns their count.
MUAV_HD inline int allocate_tasks(Sim& S, const muav_alloc_opts& O, int e, int16_t* out_agent, int16_t* out_tid) {
  View& V = S.V;
  const muav_config& C = S.C();
  const int A = V.L->D.A, TC = V.L->D.TC;
  const int t = HIv(T);
  HIv(N_CALLS) += 1;
  const bool ev_hit = (HIv(EV_TAGMASK) & O.event_mask) != 0;
  const int interval = O.replan_interval > 0 ? O.replan_interval : 1;
  if (O.mode == 1) {
    if (!((t - HIv(LAST_PLAN_STEP)) >= interval || ev_hit)) return 0;
  } else if (O.mode == 2) {
    if (!(t == 0 || (t % interval) == 0 || ev_hit)) return 0;
  } else {
    return 0;
  }
  AllocScratch W = carve_scratch(S.scratch, A, TC);
  const uint8_t* reserved = O.d_reserved ? O.d_reserved + (size_t)e * A : nullptr;
  const double* pri = O.d_priorities ? O.d_priorities + (size_t)e * TC : nullptr;
  const float* scores = O.d_edge_scores ? O.d_edge_scores + (size_t)e * O.score_rows * O.score_cols : nullptr;
  const int n_tasks = HIv(N_TASKS);

  int n_free = 0;
  for (int a = 0; a < A; ++a)
    if (V.a_state()[a] != -1 && !(reserved && reserved[a])) W.free_agents[n_free++] = (int16_t)a;
  // open task list (+ residuals, + pair-token column of each task)
  int n_open = 0;
  int tok_j = 0;
  for (int k = 0; k < n_tasks; ++k) {
    if (V.k_status()[k] == 2) continue;
    int col = -1;
    if (O.pair_tokens) {
      int ti = V.k_type()[k];
      if (!(V.k_alloc()[ti * TC + k] < V.k_cur()[ti * TC + k])) continue;  // AttentionRAH.py:67-71
      if (tok_j >= O.score_cols) break;                                    // open_tasks[:max_tasks]
      col = tok_j++;
    }
    double r = residual_demand(S, k);
    if (r > 0) {
      W.open_t[n_open] = (int16_t)k;
      W.resid[n_open] = r;
      W.tokcol[n_open] = (int16_t)col;
      ++n_open;
    }
  }
  if (n_free == 0 || n_open == 0) return 0;
  const double mc = O.max_coord > 1.0 ? O.max_coord : 1.0;
  int n_pairs = 0;
  while (n_free > 0) {
    int nc = 0;
    for (int q = 0; q < n_open; ++q)
      if (W.resid[q] > 1e-9) W.round_tasks[nc++] = (int16_t)q;
    if (nc == 0) break;
    const int nr = n_free;
    for (int i = 0; i < nr; ++i) {
      const int a = W.free_agents[i];
      const double ax = V.a_posx()[a], ay = V.a_posy()[a];
      const int at = V.a_type()[a];
      int live_row = -1;
      if (O.pair_tokens && scores) {
        live_row = 0;
        for (int b = 0; b < a; ++b)
          if (V.a_state()[b] != -1) ++live_row;
      }
      for (int j = 0; j < nc; ++j) {
        const int q = W.round_tasks[j];
        const int k = W.open_t[q];
        double cst = 1e6;
        bool vis_ok = !O.use_visibility || S.known_bit(a, k);
        int el = V.k_elig()[k];
        bool el_ok = (el == 0) || ((el >> at) & 1);
        if (vis_ok && el_ok) {
          double urgency = 0.0;
          int dl = V.k_deadline()[k];
          if (dl >= 0) {
            int rem = dl - t;
            if (rem < 0) rem = 0;
            urgency = 1.0 - dmin((double)rem / 40.0, 1.0);
          }
          int ti = V.k_type()[k];
          double delivered = is_coalition(S, k) ? 1.0 : S.cap(a, ti);
          double base = 1e6;
          if (delivered > 0) {
            double dist = norm2(ax - V.k_posx()[k], ay - V.k_posy()[k]);
            double missing = dmax(W.resid[q], 1e-6);
            double p = pri ? pri[k] : 0.0;
            base = dist / mc - 0.5 * dmin(delivered, missing) - 0.4 * p - 0.6 * urgency;
          }
          if (base < 1e5 / 2) {
            double sc = 0.0;
            if (scores) {
              if (O.pair_tokens) {
                // edge_score_dict (PairCostHybrid.py:280-291): only valid edges carry a score
                int col = W.tokcol[q];
                bool valid = live_row >= 0 && live_row < O.score_rows && col >= 0 && S.known_bit(a, k) &&
                             S.cap(a, ti) > 0;
                if (valid) sc = (double)scores[live_row * O.score_cols + col];
              } else if (a < O.score_rows && k < O.score_cols) {
                sc = (double)scores[a * O.score_cols + k];
              }
            }
            cst = base - sc;
          }
        }
        W.cost[i * nc + j] = cst;
      }
    }
    if (!lsap_solve(W.cost, nr, nc, W, W.col_of_row)) {
      HIv(ERRFLAGS) |= ERR_LSAP_INFEASIBLE;
      break;
    }
    HIv(N_LSAP) += 1;
    int n_acc = 0;
    for (int i = 0; i < nr; ++i) {
      int j = W.col_of_row[i];
      if (j < 0) continue;
      if (W.cost[i * nc + j] >= 1e5 / 2) { W.col_of_row[i] = -1; continue; }
      const int a = W.free_agents[i];
      const int q = W.round_tasks[j];
      const int k = W.open_t[q];
      double delivered = is_coalition(S, k) ? 1.0 : S.cap(a, V.k_type()[k]);
      out_agent[n_pairs] = (int16_t)a;
      out_tid[n_pairs] = (int16_t)(k + 1);
      ++n_pairs;
      W.resid[q] = dmax(W.resid[q] - delivered, 0.0);
      ++n_acc;
    }
    if (n_acc == 0) break;
    int w = 0;
    for (int i = 0; i < nr; ++i)
      if (W.col_of_row[i] < 0) W.free_agents[w++] = W.free_agents[i];
    n_free = w;
  }
  HIv(LAST_PLAN_STEP) = t;
  HIv(N_REPLANS) += 1;
  return n_pairs;
}

}  // namespace muav
