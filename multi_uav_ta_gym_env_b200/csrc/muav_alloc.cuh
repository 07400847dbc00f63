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
  int32_t* ctrl;  // small control block shared by the lanes
  int16_t *path, *col4row, *row4col, *remaining, *SR, *SC, *free_agents, *round_tasks, *open_t, *tokcol, *col_of_row,
      *live_row;
  double *plan_pri, *plan_score;  // planner: task priorities [TC], lock scores [A]
  uint8_t* plan_reserved;         // planner: reserved agents [A]
};

MUAV_HD inline AllocScratch carve_scratch(char* p, int A, int TC, int IC = 0) {
  AllocScratch S;
  int M = A > TC ? A : TC;
  double* d = (double*)p;
  S.cost = d; d += A * TC;
  S.u = d; d += M;
  S.v = d; d += M;
  S.spc = d; d += M;
  S.resid = d; d += TC;
  S.ctrl = (int32_t*)d; d += 4;
  int16_t* s = (int16_t*)d;
  S.path = s; s += M;
  S.col4row = s; s += M;
  S.row4col = s; s += M;
  S.remaining = s; s += M;
  S.SR = s; s += M;
  S.SC = s; s += M;
  S.free_agents = s; s += A;
  S.col_of_row = s; s += A;
  S.live_row = s; s += A;
  S.round_tasks = s; s += TC;
  S.open_t = s; s += TC;
  S.tokcol = s; s += TC;
  // planner block starts at the next 8-byte boundary
  uintptr_t q = ((uintptr_t)s + 7) & ~(uintptr_t)7;
  S.plan_pri = (double*)q;
  S.plan_score = S.plan_pri + (IC > TC ? IC : TC);  // priorities are indexed by task id
  S.plan_reserved = (uint8_t*)(S.plan_score + A);
  return S;
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double warp_min_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}
__device__ __forceinline__ int warp_min_i32(int v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_max_i32(int v) { return __reduce_max_sync(0xffffffffu, v); }
#else
inline double warp_min_f64(double v) { return v; }
inline int warp_min_i32(int v) { return v; }
inline int warp_max_i32(int v) { return v; }
#endif

// Rectangular LSAP, warp-collective: every lane of the warp must call it (host: lane 0 of 1).
// cost: row-major nr x nc (ld = nc).  col_of_row[i] = column of row i or -1.
// SciPy's sequential column scan is spread over the lanes; its tie rule -- among equal minima the
// LAST scanned unassigned column wins, otherwise the FIRST scanned column -- is reproduced by
// reducing (min value, first index, last unassigned index) over the lanes.
MUAV_HD MUAV_NI_A inline bool lsap_solve(const double* cost, int nr, int nc, AllocScratch& S, int16_t* col_of_row, int lane,
                               int nlanes) {
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
  for (int i = lane; i < R; i += nlanes) { u[i] = 0.0; col4row[i] = -1; }
  for (int j = lane; j < Cn; j += nlanes) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
  MUAV_WARP_SYNC();
  for (int cur = 0; cur < R; ++cur) {
    for (int it = lane; it < Cn; it += nlanes) {
      remaining[it] = (int16_t)(Cn - it - 1);
      SC[it] = 0;
      spc[it] = INFINITY;
    }
    for (int i = lane; i < R; i += nlanes) SR[i] = 0;
    MUAV_WARP_SYNC();
    double min_val = 0.0;
    int num_remaining = Cn;
    int i = cur;
    int sink = -1;
    while (sink == -1) {
      if (lane == 0) SR[i] = 1;
      const double ui = u[i];
      const double* crow = cost + i * si;
      double bv = INFINITY;
      int first = 0x7fffffff, lastun = -1;
      for (int it = lane; it < num_remaining; it += nlanes) {
        const int j = remaining[it];
        const double r = min_val + crow[j * sj] - ui - v[j];
        double sp = spc[j];
        if (r < sp) {
          path[j] = (int16_t)i;
          spc[j] = r;
          sp = r;
        }
        const bool un = row4col[j] == -1;
        if (sp < bv) {
          bv = sp;
          first = it;
          lastun = un ? it : -1;
        } else if (sp == bv) {
          if (first == 0x7fffffff) first = it;
          if (un) lastun = it;
        }
      }
      const double m = warp_min_f64(bv);
      const bool cand = bv == m;
      const int f = warp_min_i32(cand ? first : 0x7fffffff);
      const int l = warp_max_i32(cand ? lastun : -1);
      const int index = l >= 0 ? l : f;
      min_val = m;
      if (index == 0x7fffffff || m == INFINITY) return false;
      const int j = remaining[index];
      const int rj = row4col[j];
      if (rj == -1) sink = j;
      else i = rj;
      MUAV_WARP_SYNC();
      --num_remaining;
      if (lane == 0) {
        SC[j] = 1;
        remaining[index] = remaining[num_remaining];
      }
      MUAV_WARP_SYNC();
    }
    if (lane == 0) u[cur] += min_val;
    for (int r = lane; r < R; r += nlanes)
      if (SR[r] && r != cur) u[r] += min_val - spc[col4row[r]];
    for (int j = lane; j < Cn; j += nlanes)
      if (SC[j]) v[j] -= min_val - spc[j];
    MUAV_WARP_SYNC();
    if (lane == 0) {
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
    MUAV_WARP_SYNC();
  }
  for (int r = lane; r < nr; r += nlanes) col_of_row[r] = -1;
  MUAV_WARP_SYNC();
  if (tr) {
    for (int vv = lane; vv < R; vv += nlanes) col_of_row[col4row[vv]] = (int16_t)vv;
  } else {
    for (int r = lane; r < R; r += nlanes) col_of_row[r] = col4row[r];
  }
  MUAV_WARP_SYNC();
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
  int TC = S.V.lay().D.TC;
  double r = S.V.k_cur_ti()[k] - S.V.k_alloc_ti()[k];
  return r > 0.0 ? r : 0.0;
}

// HungarianAllocator.allocate_tasks (:72-208) fused with _open_tasks (paper_eval.py:96-101).
// Warp-collective (host: lane 0 of 1): lane 0 takes the sequential decisions (replan rule, open list,
// residuals, acceptance in row order), the cost matrix and the LSAP scans are spread over the lanes.
// Writes ordered (agent, task id) pairs; returns their count (same value on every lane).
MUAV_HD inline double coalition_edge_score(const Sim& S, int a, int k, int t);

MUAV_HD MUAV_NI_A inline int allocate_tasks(Sim& S, const muav_alloc_opts& O, int e, int16_t* out_agent, int16_t* out_tid, int lane,
                                  int nlanes, bool local_ptrs = false) {
  View& V = S.V;
  const int A = V.lay().D.A, TC = V.lay().D.TC;
  const int t = HIv(T);
  AllocScratch W = carve_scratch(S.scratch, A, TC);
  const size_t eo = local_ptrs ? 0 : (size_t)e;  // planner-produced arrays live in this env's scratch
  const uint8_t* reserved = O.d_reserved ? O.d_reserved + eo * A : nullptr;
  const int IC = V.lay().D.IC;
  const double* pri = O.d_priorities ? O.d_priorities + eo * IC : nullptr;
  const size_t sc_off = (size_t)e * O.score_rows * O.score_cols;
  const float* scores = (O.d_edge_scores && !O.score_f64) ? (const float*)O.d_edge_scores + sc_off : nullptr;
  const double* scores64 = (O.d_edge_scores && O.score_f64) ? (const double*)O.d_edge_scores + sc_off : nullptr;
  enum { C_GO = 0, C_NFREE = 1, C_NOPEN = 2, C_NC = 3, C_NPAIRS = 4, C_STOP = 5 };
  if (lane == 0) {
    HIv(N_CALLS) += 1;
    const bool ev_hit = (HIv(EV_TAGMASK) & O.event_mask) != 0;
    const int interval = O.replan_interval > 0 ? O.replan_interval : 1;
    bool go;
    if (O.mode == 1) go = (t - HIv(LAST_PLAN_STEP)) >= interval || ev_hit;
    else if (O.mode == 2) go = t == 0 || (t % interval) == 0 || ev_hit;
    else go = O.mode == 3;
    int n_free = 0, n_open = 0;
    if (go) {
      const int n_tasks = HIv(N_TASKS);
      int live = 0;
      for (int a = 0; a < A; ++a) {
        if (V.a_state()[a] == -1) continue;
        if (!(reserved && reserved[a])) {
          W.free_agents[n_free] = (int16_t)a;
          W.live_row[n_free] = (int16_t)live;
          ++n_free;
        }
        ++live;
      }
      // open task list (+ residuals, + pair-token column of each task)
      int tok_j = 0;
      const int32_t* order = O.d_task_order ? O.d_task_order + (size_t)e * IC : nullptr;
      // without a caller-given order the candidates are the ids set in open_mask: the allocator runs at a step
      // boundary, where the mask of the last scan is exact, so closed ids (most of them in WPS_escort) are never touched
      const int KWn = (n_tasks + 31) >> 5;
      int wd = 0;
      uint32_t bits = order ? 0u : (KWn > 0 ? V.open_mask()[0] : 0u);
      for (int it = 0;; ++it) {
        int k;
        if (order) {
          if (it >= IC) break;
          k = order[it];
          if (k < 0) break;
          if (k >= n_tasks) continue;
        } else {
          while (!bits && ++wd < KWn) bits = V.open_mask()[wd];
          if (!bits) break;
          k = (wd << 5) + ctz32(bits);
          bits &= bits - 1;
        }
        if (V.k_status()[k] == 2) continue;
        int col = -1;
        if (O.pair_tokens == 2) {
          col = it < O.score_cols ? it : -1;  // column = position in the escort token list
        } else if (O.pair_tokens) {
          int ti = V.k_type()[k];
          if (!(V.k_alloc_ti()[k] < V.k_cur_ti()[k])) continue;  // AttentionRAH.py:67-71
          if (tok_j >= O.score_cols) break;                                    // open_tasks[:max_tasks]
          col = tok_j++;
        }
        double r = residual_demand(S, k);
        if (r > 0 && n_open < TC) {
          W.open_t[n_open] = (int16_t)k;
          W.resid[n_open] = r;
          W.tokcol[n_open] = (int16_t)col;
          ++n_open;
        }
      }
      if (n_free == 0 || n_open == 0) go = false;  // returns [] before touching last_plan_step (:120-121)
    }
    W.ctrl[C_GO] = go ? 1 : 0;
    W.ctrl[C_NFREE] = n_free;
    W.ctrl[C_NOPEN] = n_open;
    W.ctrl[C_NPAIRS] = 0;
    W.ctrl[C_STOP] = 0;
  }
  MUAV_WARP_SYNC();
  if (!W.ctrl[C_GO]) return 0;
  const double mc = O.max_coord > 1.0 ? O.max_coord : 1.0;
  for (;;) {
    if (lane == 0) {
      int nc = 0;
      if (W.ctrl[C_NFREE] > 0) {
        const int n_open = W.ctrl[C_NOPEN];
        for (int q = 0; q < n_open; ++q)
          if (W.resid[q] > 1e-9) W.round_tasks[nc++] = (int16_t)q;
      }
      W.ctrl[C_NC] = nc;
    }
    MUAV_WARP_SYNC();
    const int nc = W.ctrl[C_NC];
    const int nr = W.ctrl[C_NFREE];
    if (nc == 0 || nr == 0) break;
    // ---- cost matrix: independent entries, spread over the lanes
    for (int idx = lane; idx < nr * nc; idx += nlanes) {
      const int i = idx / nc, j = idx - i * nc;
      const int a = W.free_agents[i];
      const int q = W.round_tasks[j];
      const int k = W.open_t[q];
      const int at = V.a_type()[a];
      double cst = 1e6;
      const bool vis_ok = !O.use_visibility || S.known_bit(a, k);
      const int el = V.k_elig()[k];
      const bool el_ok = (el == 0) || ((el >> at) & 1);
      if (vis_ok && el_ok) {
        double urgency = 0.0;
        const int dl = V.k_deadline()[k];
        if (dl >= 0) {
          int rem = dl - t;
          if (rem < 0) rem = 0;
          urgency = 1.0 - dmin(ddiv((double)rem, 40.0), 1.0);
        }
        const int ti = V.k_type()[k];
        const double delivered = is_coalition(S, k) ? 1.0 : S.cap(a, ti);
        double base = 1e6;
        if (delivered > 0) {
          const double dist = norm2(V.a_posx()[a] - V.k_posx()[k], V.a_posy()[a] - V.k_posy()[k]);
          const double missing = dmax(W.resid[q], 1e-6);
          const double p = pri ? pri[k] : 0.0;
          base = ddiv(dist, mc) - 0.5 * dmin(delivered, missing) - 0.4 * p - 0.6 * urgency;
        }
        if (base < 1e5 / 2) {
          double sc = 0.0;
          if (MUAV_F_PLANNER(O.planner) == 2) {
            sc = coalition_edge_score(S, a, k, t);
          } else if (MUAV_F_PLANNER(O.planner) == 5) {
            // urgency_edge_scores (PairCostHybrid.py:68-86) on the valid edges of the pair tokens; stored as float32
            const int col = W.tokcol[q];
            const int row = W.live_row[i];
            if (row < O.score_rows && col >= 0 && S.known_bit(a, k) && S.cap(a, ti) > 0) {
              double scar = 0.0;
              if (!(!(S.C().sense_radius != 0.0) && !(S.C().threat_delay != 0))) {
                int cnt = 0, n_live = 0;
                for (int b = 0; b < A; ++b) {
                  cnt += S.known_bit(b, k) ? 1 : 0;
                  n_live += V.a_state()[b] != -1;
                }
                scar = 1.0 - dmin(ddiv((double)cnt, (double)(n_live > 1 ? n_live : 1)), 1.0);
              }
              const double d = ddiv(norm2(V.a_posx()[a] - V.k_posx()[k], V.a_posy()[a] - V.k_posy()[k]),
                                    dmax(S.C().max_coord, 1.0));
              const double raw = 0.5 * urgency + 0.3 * scar - 0.4 * d;
              sc = (double)(float)dmin(dmax(raw, -0.35), 0.35);
            }
          } else if (scores64) {
            if (a < O.score_rows && k < O.score_cols) sc = scores64[a * O.score_cols + k];
          } else if (scores) {
            if (O.pair_tokens == 2) {
              // AttentionEscort.edge_score_dict (AttentionEscort.py:478-489): every unpadded (row, column)
              const int col = W.tokcol[q];
              const int row = W.live_row[i];
              if (row < O.score_rows && col >= 0) sc = (double)scores[row * O.score_cols + col];
            } else if (O.pair_tokens) {
              // edge_score_dict (PairCostHybrid.py:280-291): only valid edges carry a score
              const int col = W.tokcol[q];
              const int row = W.live_row[i];
              const bool valid = row < O.score_rows && col >= 0 && S.known_bit(a, k) && S.cap(a, ti) > 0;
              if (valid) sc = (double)scores[row * O.score_cols + col];
            } else if (a < O.score_rows && k < O.score_cols) {
              sc = (double)scores[a * O.score_cols + k];
            }
          }
          cst = base - sc;
        }
      }
      W.cost[idx] = cst;
    }
    MUAV_WARP_SYNC();
    const bool ok = lsap_solve(W.cost, nr, nc, W, W.col_of_row, lane, nlanes);
    if (lane == 0) {
      if (!ok) {
        HIv(ERRFLAGS) |= ERR_LSAP_INFEASIBLE;
        W.ctrl[C_STOP] = 1;
      } else {
        HIv(N_LSAP) += 1;
        int n_pairs = W.ctrl[C_NPAIRS];
        int n_acc = 0;
        for (int i = 0; i < nr; ++i) {
          int j = W.col_of_row[i];
          if (j < 0) continue;
          if (W.cost[i * nc + j] >= 1e5 / 2) { W.col_of_row[i] = -1; continue; }
          const int a = W.free_agents[i];
          const int q = W.round_tasks[j];
          const int k = W.open_t[q];
          const double delivered = is_coalition(S, k) ? 1.0 : S.cap(a, V.k_type()[k]);
          out_agent[n_pairs] = (int16_t)a;
          out_tid[n_pairs] = (int16_t)(k + 1);
          ++n_pairs;
          W.resid[q] = dmax(W.resid[q] - delivered, 0.0);
          ++n_acc;
        }
        W.ctrl[C_NPAIRS] = n_pairs;
        if (n_acc == 0) {
          W.ctrl[C_STOP] = 1;
        } else {
          int w = 0;
          for (int i = 0; i < nr; ++i)
            if (W.col_of_row[i] < 0) {
              W.free_agents[w] = W.free_agents[i];
              W.live_row[w] = W.live_row[i];
              ++w;
            }
          W.ctrl[C_NFREE] = w;
        }
      }
    }
    MUAV_WARP_SYNC();
    if (W.ctrl[C_STOP]) break;
  }
  if (lane == 0) {
    HIv(LAST_PLAN_STEP) = t;
    HIv(N_REPLANS) += 1;
  }
  const int n_pairs = W.ctrl[C_NPAIRS];
  MUAV_WARP_SYNC();
  return n_pairs;
}

// ---- Performance-Impact market allocator (TaskAllocation/MarketBased/PerformanceImpact.py:59-224) --------------------
// as the reference's drivers call it: max_tasks_per_agent = 1 (experiments/wps_eval.py:147-156, escort_eval.py:162-170).
// With one task per agent every path has at most one slot, so
//   IPI(agent, slot) = provisional RPI = RPI of the winner = _path_cost([slot])                       (:243-261,263-301)
//                    = start [+ 200 + (start - deadline) if start > deadline] - 5 cap',   start = max(nft, t) + dist / speed
// is a constant of the call: it is computed once per (agent, open task) over the lanes, and the inclusion loop (:106-165)
// is a repeated lane-parallel arg-min of (IPI, agent id, slot key) under the steal rule (:132-140).  The consensus phase
// (:168-205) cannot change anything in this case (each slot has one claimant and the deadline filter repeats the test the
// inclusion phase already made).  Slot keys are compared as the reference compares them: as the STRINGS "<id>#c<k>" /
// "<id>#r<k>" (CBBA.py:46-65), i.e. ids in lexicographic decimal order.

// "<x>#" < "<y>#" as strings ('#' sorts before every digit)
MUAV_HD inline bool slot_id_less(int x, int y) {
  int nx = 1, ny = 1, px = 10, py = 10;
  while (x >= px) { px *= 10; ++nx; }
  while (y >= py) { py *= 10; ++ny; }
  if (nx == ny) return x < y;
  if (nx < ny) {
    int sh = 1;
    for (int i = nx; i < ny; ++i) sh *= 10;
    return x <= y / sh;   // equal prefix: the shorter string is smaller
  }
  int sh = 1;
  for (int i = ny; i < nx; ++i) sh *= 10;
  return x / sh < y;
}

// ---- bundles (max_tasks_per_agent > 1): the inclusion phase over agent PATHS (PerformanceImpact.py:106-165).
// A path is a list of slots in visiting order; _schedule / _path_cost (:227-261) walk it: start = now + dist / speed,
// now = start + duration[type].  IPI of (agent, slot) = min over the feasible insertion points of cost(path + slot) -
// cost(path) (_best_inclusion_impact :263-283, first minimum within 1e-9); its provisional RPI (:285-301) is the same two
// costs subtracted the same way, i.e. the same number.
constexpr int PI_MAX_BUNDLE = 4;

// cost of the path p[0..n) of agent a; *first_bad = index of the first entry that starts later than its deadline + 1e-6
// (n when every entry is feasible: _filter_feasible :303-311)
MUAV_HD inline double pi_path_cost(Sim& S, const int16_t* open_t, const int16_t* slot_q, int a, const int16_t* p, int n, int t,
                                   int* first_bad) {
  View& V = S.V;
  double px = V.a_posx()[a], py = V.a_posy()[a];
  double now = dmax(V.a_nft()[a], (double)t);
  const double speed = dmax(S.speed_of(a) != 0.0 ? S.speed_of(a) : 1.0, 1e-6);
  double cost = 0.0;
  int bad = n;
  for (int j = 0; j < n; ++j) {
    const int k = open_t[slot_q[p[j]]];
    const double tx = V.k_posx()[k], ty = V.k_posy()[k];
    const double start = now + ddiv(norm2(px - tx, py - ty), speed);
    const int dl = V.k_deadline()[k];
    if (bad == n && dl >= 0 && start > (double)dl + 1e-6) bad = j;
    cost = cost + start;
    if (dl >= 0 && start > (double)dl) {
      const double late = 200.0 + (start - (double)dl);
      cost = cost + late;
    }
    const int ti = V.k_type()[k];
    const double cp = S.cap(a, ti);
    const double bonus = 5.0 * (is_coalition(S, k) ? dmax(cp, 0.5) : cp);
    cost = cost - bonus;
    px = tx;
    py = ty;
    now = start + (double)S.C().duration[ti];
  }
  *first_bad = bad;
  return cost;
}

// paths: [live row][PI_MAX_BUNDLE] slots, plen: [live row]; slot_rpi: RPI of the slot's winner when it won the slot
MUAV_HD inline int pi_bundles(Sim& S, const muav_alloc_opts& O, int e, const AllocScratch& W, int nr, int nq, int ns,
                              const int16_t* slot_q, const int16_t* slot_rank, int16_t* slot_win, int16_t* out_agent,
                              int16_t* out_tid, int lane, int nlanes) {
  View& V = S.V;
  const int A = V.lay().D.A;
  const int t = HIv(T);
  int MB = O.max_tasks_per_agent;
  if (MB > PI_MAX_BUNDLE) MB = PI_MAX_BUNDLE;
  int16_t* paths = W.path;         // the six LSAP index arrays are contiguous: room for nr * PI_MAX_BUNDLE entries
  int16_t* plen = W.live_row;
  double* slot_rpi = W.cost;       // the cost matrix is not used by this form
  enum { C_BEST = 4, C_AT = 5 };
  if (A * V.lay().D.TC < ns) {   // the RPI array lives in the cost matrix: tiny fleets only
    if (lane == 0) HIv(ERRFLAGS) |= ERR_NO_SPACE;
    MUAV_WARP_SYNC();
    return 0;
  }
  for (int i = lane; i < nr; i += nlanes) plen[i] = 0;
  MUAV_WARP_SYNC();
  const int max_it = ns * (nr > 1 ? nr : 1);
  for (int it = 0; it < max_it; ++it) {
    double bc = INFINITY;
    int bkey = 0x7fffffff, bidx = -1, bat = 0;
    for (int idx = lane; idx < nr * ns; idx += nlanes) {
      const int i = idx / ns, s = idx - i * ns;
      const int n = plen[i];
      if (n >= MB) continue;
      const int w = slot_win[s];
      if (w == i) continue;
      const int q = slot_q[s];
      const int16_t* path = paths + i * PI_MAX_BUNDLE;
      bool owned = false;
      for (int j = 0; j < n; ++j) owned = owned || slot_q[path[j]] == q;
      if (owned) continue;
      const int a = W.free_agents[i];
      const int k = W.open_t[q];
      {   // agent_eligible (CBBA.py:27-43)
        const int el = V.k_elig()[k];
        bool ok = !O.use_visibility || S.known_bit(a, k);
        ok = ok && ((el == 0) || ((el >> V.a_type()[a]) & 1));
        ok = ok && S.qfind(a, k + 1) < 0;
        ok = ok && (is_coalition(S, k) || S.cap(a, V.k_type()[k]) > 0);
        if (!ok) continue;
      }
      int bad = 0;
      const double base = pi_path_cost(S, W.open_t, slot_q, a, path, n, t, &bad);
      double best = INFINITY;
      int at = 0;
      for (int pos = 0; pos <= n; ++pos) {
        int16_t mapped[PI_MAX_BUNDLE + 1];
        for (int j = 0; j < pos; ++j) mapped[j] = path[j];
        mapped[pos] = (int16_t)s;
        for (int j = pos; j < n; ++j) mapped[j + 1] = path[j];
        const double c = pi_path_cost(S, W.open_t, slot_q, a, mapped, n + 1, t, &bad);
        if (bad != n + 1) continue;
        const double ipi = c - base;
        if (ipi < best - 1e-9) { best = ipi; at = pos; }
      }
      if (!(best < INFINITY) || !(best > -INFINITY)) continue;
      if (w >= 0) {
        const double cur = slot_rpi[s];
        if (best < cur - 1e-9) continue;
        if (fabs(best - cur) <= 1e-9 && a >= W.free_agents[w]) continue;
      }
      const int key = (a << 16) | slot_rank[s];
      if (best < bc || (best == bc && key < bkey)) { bc = best; bkey = key; bidx = idx; bat = at; }
    }
    const double m = warp_min_f64(bc);
    const int kmin = warp_min_i32((bidx >= 0 && bc == m) ? bkey : 0x7fffffff);
    if (kmin == 0x7fffffff) break;
    if (bidx >= 0 && bc == m && bkey == kmin) { W.ctrl[C_BEST] = bidx; W.ctrl[C_AT] = bat; }   // (agent, slot rank) is unique
    MUAV_WARP_SYNC();
    if (lane == 0) {
      const int idx = W.ctrl[C_BEST], at = W.ctrl[C_AT];
      const int i = idx / ns, s = idx - i * ns;
      const int prev = slot_win[s];
      if (prev >= 0 && prev != i) {
        int16_t* pp = paths + prev * PI_MAX_BUNDLE;
        int n2 = 0;
        for (int j = 0; j < plen[prev]; ++j)
          if (pp[j] != s) pp[n2++] = pp[j];
        plen[prev] = (int16_t)n2;
      }
      int16_t* path = paths + i * PI_MAX_BUNDLE;
      const int n = plen[i];
      for (int j = n; j > at; --j) path[j] = path[j - 1];
      path[at] = (int16_t)s;
      plen[i] = (int16_t)(n + 1);
      slot_win[s] = (int16_t)i;
      // _removal_impact of the slot in its new path (:285-301)
      int16_t rest[PI_MAX_BUNDLE];
      int n2 = 0, bad = 0;
      for (int j = 0; j <= n; ++j)
        if (j != at) rest[n2++] = path[j];
      const int a = W.free_agents[i];
      slot_rpi[s] = pi_path_cost(S, W.open_t, slot_q, a, path, n + 1, t, &bad) - pi_path_cost(S, W.open_t, slot_q, a, rest, n2, t, &bad);
    }
    MUAV_WARP_SYNC();
  }
  // consensus clean-up (:168-205): every slot has one claimant here (a stolen slot leaves its loser's path at once), so
  // the only thing left to do is to cut paths back to their feasible prefix until nothing changes
  if (lane == 0) {
    for (int it = 0; it < 40; ++it) {
      bool changed = false;
      for (int i = 0; i < nr; ++i) {
        int bad = 0;
        pi_path_cost(S, W.open_t, slot_q, W.free_agents[i], paths + i * PI_MAX_BUNDLE, plen[i], t, &bad);
        if (bad < plen[i]) {
          for (int j = bad; j < plen[i]; ++j)
            if (slot_win[paths[i * PI_MAX_BUNDLE + j]] == i) slot_win[paths[i * PI_MAX_BUNDLE + j]] = -1;
          plen[i] = (int16_t)bad;
          changed = true;
        }
      }
      if (!changed) break;
    }
    int nb = 0;
    int32_t* bp = O.d_bundle_pairs ? O.d_bundle_pairs + (size_t)e * A * O.max_tasks_per_agent : nullptr;
    for (int i = 0; i < nr; ++i)
      for (int j = 0; j < plen[i]; ++j) {
        if (bp) bp[nb] = ((int)W.free_agents[i] << 16) | (W.open_t[slot_q[paths[i * PI_MAX_BUNDLE + j]]] + 1);
        ++nb;
      }
    if (O.d_n_bundle_pairs) O.d_n_bundle_pairs[e] = nb;
  }
  MUAV_WARP_SYNC();
  // the step's pairs: the first task of every path (_apply_assign keeps the first pair of an agent)
  int n_pairs = 0;
  for (int i = 0; i < nr; ++i) {
    if (plen[i] == 0) continue;
    if (lane == 0) {
      out_agent[n_pairs] = W.free_agents[i];
      out_tid[n_pairs] = (int16_t)(W.open_t[slot_q[paths[i * PI_MAX_BUNDLE]]] + 1);
    }
    ++n_pairs;
  }
  MUAV_WARP_SYNC();
  return n_pairs;
}

MUAV_HD MUAV_NI_A inline int pi_allocate(Sim& S, const muav_alloc_opts& O, int e, int16_t* out_agent, int16_t* out_tid, int lane,
                               int nlanes) {
  View& V = S.V;
  const int A = V.lay().D.A, TC = V.lay().D.TC;
  const int t = HIv(T);
  AllocScratch W = carve_scratch(S.scratch, A, TC);
  const int M = A > TC ? A : TC;
  const int SMAX = 3 * M + TC;               // u, v, spc, resid reused as four int16 arrays of SMAX entries
  int16_t* slot_q = (int16_t*)W.u;           // open-task column of the slot
  int16_t* slot_rank = slot_q + SMAX;        // position of the slot key in string order
  int16_t* slot_win = slot_rank + SMAX;      // live row of the current winner or -1
  int16_t* slot_k = slot_win + SMAX;         // k of "<id>#?k"
  int16_t* q_rank = W.round_tasks;           // string rank of the task id among the open tasks
  int16_t* row_slot = W.col_of_row;          // slot held by a live row or -1
  const uint8_t* reserved = O.d_reserved ? O.d_reserved + (size_t)e * A : nullptr;
  enum { C_GO = 0, C_NFREE = 1, C_NOPEN = 2, C_NSLOT = 3, C_BEST = 4 };
  if (lane == 0) {
    HIv(N_CALLS) += 1;
    const bool ev_hit = (HIv(EV_TAGMASK) & O.event_mask) != 0;
    const int interval = O.replan_interval > 0 ? O.replan_interval : 1;
    bool go;
    if (O.mode == 1) go = (t - HIv(LAST_PLAN_STEP)) >= interval || ev_hit;
    else if (O.mode == 2) go = t == 0 || (t % interval) == 0 || ev_hit;
    else go = O.mode == 3;
    int n_free = 0, n_open = 0, n_slot = 0;
    if (go) {
      // every call that passes the rule counts as a replan, also the empty ones (:80-93,222-223)
      HIv(LAST_PLAN_STEP) = t;
      HIv(N_REPLANS) += 1;
      for (int a = 0; a < A; ++a)
        if (V.a_state()[a] != -1 && !(reserved && reserved[a])) W.free_agents[n_free++] = (int16_t)a;
      // _open_tasks (paper_eval.py:96-101) + expand_slot_keys (CBBA.py:46-65)
      const int n_tasks = HIv(N_TASKS);
      const int IC = V.lay().D.IC;
      const int32_t* order = O.d_task_order ? O.d_task_order + (size_t)e * IC : nullptr;   // the caller's `tasks` argument
      const int KWn = (n_tasks + 31) >> 5;
      int wd = 0;
      uint32_t bits = order ? 0u : (KWn > 0 ? V.open_mask()[0] : 0u);
      for (int it = 0;; ++it) {
        int k;
        if (order) {
          if (it >= IC) break;
          k = order[it];
          if (k < 0) break;
          if (k >= n_tasks) continue;
        } else {
          while (!bits && ++wd < KWn) bits = V.open_mask()[wd];
          if (!bits) break;
          k = (wd << 5) + ctz32(bits);
          bits &= bits - 1;
        }
        if (V.k_status()[k] == 2) continue;
        const double rem = residual_demand(S, k);
        if (!(rem > 0) || n_open >= TC) continue;
        int ns = is_coalition(S, k) ? (int)ceil(rem) : (int)ceil(dmin(rem, 4.0));
        if (ns < 1) ns = 1;
        if (ns > 10 || n_slot + ns > SMAX) {   // two-digit slot numbers / more slots than the scratch holds: not representable
          HIv(ERRFLAGS) |= ERR_NO_SPACE;
          ns = ns > 10 ? 10 : ns;
          if (n_slot + ns > SMAX) break;
        }
        W.open_t[n_open] = (int16_t)k;
        for (int j = 0; j < ns; ++j) {
          slot_q[n_slot] = (int16_t)n_open;
          slot_k[n_slot] = (int16_t)j;
          slot_win[n_slot] = -1;
          ++n_slot;
        }
        ++n_open;
      }
      if (n_free == 0 || n_slot == 0) go = false;
    }
    W.ctrl[C_GO] = go ? 1 : 0;
    W.ctrl[C_NFREE] = n_free;
    W.ctrl[C_NOPEN] = n_open;
    W.ctrl[C_NSLOT] = n_slot;
    if (O.d_n_bundle_pairs) O.d_n_bundle_pairs[e] = 0;
  }
  MUAV_WARP_SYNC();
  if (!W.ctrl[C_GO]) return 0;
  const int nr = W.ctrl[C_NFREE], nq = W.ctrl[C_NOPEN], ns = W.ctrl[C_NSLOT];
  // ---- string rank of every open task id, then of every slot (slots of one task are adjacent and ordered by k)
  for (int q = lane; q < nq; q += nlanes) {
    const int id = W.open_t[q] + 1;
    int r = 0;
    for (int p = 0; p < nq; ++p) r += (p != q && slot_id_less(W.open_t[p] + 1, id)) ? 1 : 0;
    q_rank[q] = (int16_t)r;
  }
  for (int i = lane; i < nr; i += nlanes) row_slot[i] = -1;
  MUAV_WARP_SYNC();
  for (int s = lane; s < ns; s += nlanes) {
    const int rq = q_rank[slot_q[s]];
    int r = 0;
    for (int p = 0; p < ns; ++p) r += q_rank[slot_q[p]] < rq ? 1 : 0;
    slot_rank[s] = (int16_t)(r + slot_k[s]);
  }
  if (O.max_tasks_per_agent > 1) {
    MUAV_WARP_SYNC();
    return pi_bundles(S, O, e, W, nr, nq, ns, slot_q, slot_rank, slot_win, out_agent, out_tid, lane, nlanes);
  }
  // ---- cost of (live row, open task): independent entries, spread over the lanes
  for (int idx = lane; idx < nr * nq; idx += nlanes) {
    const int i = idx / nq, q = idx - i * nq;
    const int a = W.free_agents[i];
    const int k = W.open_t[q];
    const int at = V.a_type()[a];
    const int ti = V.k_type()[k];
    const int el = V.k_elig()[k];
    const bool coal = is_coalition(S, k);
    double c = INFINITY;
    // agent_eligible (CBBA.py:27-43)
    bool ok = !O.use_visibility || S.known_bit(a, k);
    ok = ok && ((el == 0) || ((el >> at) & 1));
    ok = ok && S.qfind(a, k + 1) < 0;
    ok = ok && (coal || S.cap(a, ti) > 0);
    if (ok) {
      const double nft = V.a_nft()[a];
      const double now = dmax(nft, (double)t);
      const double speed = dmax(S.speed_of(a) != 0.0 ? S.speed_of(a) : 1.0, 1e-6);
      const double start = now + ddiv(norm2(V.a_posx()[a] - V.k_posx()[k], V.a_posy()[a] - V.k_posy()[k]), speed);
      const int dl = V.k_deadline()[k];
      if (!(dl >= 0 && start > (double)dl + 1e-6)) {   // _best_inclusion_impact's deadline test (:276-283)
        double cost = start;
        if (dl >= 0 && start > (double)dl) {
          const double late = 200.0 + (start - (double)dl);
          cost = cost + late;
        }
        const double cp = S.cap(a, ti);
        const double bonus = 5.0 * (coal ? dmax(cp, 0.5) : cp);
        c = cost - bonus;
      }
    }
    W.cost[idx] = c;
  }
  MUAV_WARP_SYNC();
  // ---- inclusion loop (:106-165)
  const int max_it = ns * (nr > 1 ? nr : 1);
  for (int it = 0; it < max_it; ++it) {
    double bc = INFINITY;
    int bkey = 0x7fffffff, bidx = -1;
    for (int idx = lane; idx < nr * ns; idx += nlanes) {
      const int i = idx / ns, s = idx - i * ns;
      if (row_slot[i] >= 0) continue;
      const int q = slot_q[s];
      const double c = W.cost[i * nq + q];
      if (!(c < INFINITY)) continue;
      const int a = W.free_agents[i];
      const int w = slot_win[s];
      if (w >= 0) {
        const double cur = W.cost[w * nq + q];
        if (c < cur - 1e-9) continue;
        if (fabs(c - cur) <= 1e-9 && a >= W.free_agents[w]) continue;
      }
      const int key = (a << 16) | slot_rank[s];
      if (c < bc || (c == bc && key < bkey)) { bc = c; bkey = key; bidx = idx; }
    }
    const double m = warp_min_f64(bc);
    const int kmin = warp_min_i32((bidx >= 0 && bc == m) ? bkey : 0x7fffffff);
    if (kmin == 0x7fffffff) break;
    if (bidx >= 0 && bc == m && bkey == kmin) W.ctrl[C_BEST] = bidx;   // (agent, slot rank) is unique
    MUAV_WARP_SYNC();
    if (lane == 0) {
      const int idx = W.ctrl[C_BEST];
      const int i = idx / ns, s = idx - i * ns;
      const int prev = slot_win[s];
      if (prev >= 0 && prev != i) row_slot[prev] = -1;
      row_slot[i] = (int16_t)s;
      slot_win[s] = (int16_t)i;
    }
    MUAV_WARP_SYNC();
  }
  int n_pairs = 0;
  for (int i = 0; i < nr; ++i) {
    const int s = row_slot[i];
    if (s < 0) continue;
    if (lane == 0) {
      out_agent[n_pairs] = W.free_agents[i];
      out_tid[n_pairs] = (int16_t)(W.open_t[slot_q[s]] + 1);
    }
    ++n_pairs;
  }
  MUAV_WARP_SYNC();
  return n_pairs;
}

// urgency (_urgency, AttentionRAH.py:29-34)
MUAV_HD inline double task_urgency(const View& V, int k, int t) {
  int dl = V.k_deadline()[k];
  if (dl < 0) return 0.0;
  int rem = dl - t;
  if (rem < 0) rem = 0;
  return 1.0 - dmin(ddiv((double)rem, 40.0), 1.0);
}

// UrgencyCoalition edge score (AttentionEscort.py:727-751) incl. _threat_stats pressure (:46-65)
MUAV_HD inline double coalition_edge_score(const Sim& S, int a, int k, int t) {
  const View& V = S.V;
  const double mc = S.C().max_coord;
  double ax = V.k_posx()[k], ay = V.k_posy()[k];
  int prot = V.k_prot_agent()[k];
  if (prot >= 0) { ax = V.a_posx()[prot]; ay = V.a_posy()[prot]; }
  double best = mc;
  const int na = V.hi()[HI_N_ACTIVE];
  for (int i = 0; i < na; ++i) {
    int hid = V.h_order()[i];
    if (V.h_status()[hid] == 2) continue;
    double d = norm2(V.h_posx()[hid] - ax, V.h_posy()[hid] - ay);
    best = dmin(best, d);
  }
  const double pressure = 1.0 - dmin(ddiv(best, mc), 1.0);
  const double urg = task_urgency(V, k, t);
  const int ti = V.k_type()[k];
  const int at = V.a_type()[a];
  const double is_escort = V.k_kind()[k] == 1 ? 1.0 : 0.0;
  const double cp = S.cap(a, ti) > 0 ? S.cap(a, ti) : 0.0;
  const double dist = ddiv(norm2(V.a_posx()[a] - V.k_posx()[k], V.a_posy()[a] - V.k_posy()[k]), mc);
  double score = 0.45 * urg + 0.35 * pressure * (0.5 + 0.5 * is_escort) + 0.3 * dmin(cp, 1.0) - 0.25 * dist;
  if (is_fighter(at) && (is_escort != 0.0 || ti == TT_INT)) score += 0.2;
  if (is_recon(at) && ti == TT_REC) score += 0.2;
  return dmin(dmax(score, 0.0), 1.0);
}

// apply_agent_commits (AttentionCommit.py:33-46): only agents that currently hold a real task are locked
MUAV_HD inline void apply_commit(Sim& S, int a, int horizon) {
  View& V = S.V;
  if (horizon <= 0 || V.a_state()[a] == -1) return;
  if (V.a_qlen()[a] > 0) V.a_commit()[a] = HIv(T) + horizon;
}

// Planner front ends (muav_alloc_opts.planner) + allocate_tasks.  Warp-collective.
}  // namespace muav
#include "muav_cbba.cuh"
namespace muav {

MUAV_HD inline int plan_and_allocate(Sim& S, const muav_alloc_opts& O, int e, int16_t* out_agent, int16_t* out_tid, int lane,
                                     int nlanes) {
  if (MUAV_F_PLANNER(O.planner) == 0) return allocate_tasks(S, O, e, out_agent, out_tid, lane, nlanes);
  if (O.planner == 6) return pi_allocate(S, O, e, out_agent, out_tid, lane, nlanes);
  if (O.planner == 7) return cbba_allocate(S, O, e, out_agent, out_tid, lane, nlanes);
  View& V = S.V;
  const muav_config& C = S.C();
  const int A = V.lay().D.A, TC = V.lay().D.TC;
  const int t = HIv(T);
  AllocScratch W = carve_scratch(S.scratch, A, TC, V.lay().D.IC);
  // the caller's cadence (wps_eval.py:64-73 / escort_eval.py:52-58), then plan(force=True)
  const int interval = O.replan_interval > 0 ? O.replan_interval : 1;
  const bool go = O.mode == 3 || t == 0 || (t % interval) == 0 || (HIv(EV_TAGMASK) & O.event_mask) != 0;
  if (!go) return 0;
  const bool vis_none = !(C.sense_radius != 0.0) && !(C.threat_delay != 0);
  if (lane == 0) {
    // committed_names (AttentionCommit.py:24-30)
    for (int a = 0; a < A; ++a) W.plan_reserved[a] = (V.a_state()[a] != -1 && V.a_commit()[a] > t) ? 1 : 0;
  }
  if (O.planner == 3) {
    // token column of each task in build_att_tokens' open list (AttentionRAH.py:67-71); -1 = no learned priority
    if (lane == 0) {
      const int n = HIv(N_TASKS);
      int col = 0;
      for (int k = 0; k < n; ++k) {
        double c = -1.0;
        if (V.k_status()[k] != 2) {
          const int ti = V.k_type()[k];
          if (V.k_alloc_ti()[k] < V.k_cur_ti()[k]) {
            if (col < O.score_cols) c = (double)col;
            ++col;
          }
        }
        W.plan_pri[k] = c;
      }
    }
    MUAV_WARP_SYNC();
    int n_live = 0;
    for (int a = 0; a < A; ++a) n_live += V.a_state()[a] != -1;
    const double nl = (double)(n_live > 1 ? n_live : 1);
    const int n = HIv(N_TASKS);
    const float* pv = O.d_plan_pri + (size_t)e * O.score_cols;
    for (int k = lane; k < n; k += nlanes) {
      const double c = W.plan_pri[k];
      double p = 0.0;
      if (c >= 0.0) {
        double scar = 0.0;
        if (!vis_none) {
          int cnt = 0;
          for (int a = 0; a < A; ++a) cnt += S.known_bit(a, k) ? 1 : 0;
          scar = 1.0 - dmin(ddiv((double)cnt, nl), 1.0);
        }
        p = 0.35 * task_urgency(V, k, t) + 0.40 * (double)pv[(int)c] + 0.25 * scar;
      }
      W.plan_pri[k] = p;
    }
  }
  if (O.planner == 1) {
    int n_live = 0;
    for (int a = 0; a < A; ++a) n_live += V.a_state()[a] != -1;
    const double nl = (double)(n_live > 1 ? n_live : 1);
    const int n = HIv(N_TASKS);
    for (int k = lane; k < n; k += nlanes) {
      double p = 0.0;
      if (V.k_status()[k] != 2) {
        double scar = 0.0;
        if (!vis_none) {
          int cnt = 0;
          for (int a = 0; a < A; ++a) cnt += S.known_bit(a, k) ? 1 : 0;
          scar = 1.0 - dmin(ddiv((double)cnt, nl), 1.0);
        }
        p = 0.6 * task_urgency(V, k, t) + 0.4 * scar;
      }
      W.plan_pri[k] = p;
    }
  }
  MUAV_WARP_SYNC();
  muav_alloc_opts P = O;
  P.mode = 3;
  P.d_reserved = W.plan_reserved;
  P.d_edge_scores = nullptr;
  P.d_task_order = nullptr;
  if (O.planner == 1 || O.planner == 3) {
    P.d_priorities = W.plan_pri;
    P.pair_tokens = 1;      // task list = build_att_tokens' open list (alloc < cur), not truncated
    P.score_cols = V.lay().D.IC;
    P.score_rows = 0;
    P.use_visibility = 1;
  } else if (O.planner == 5) {
    P.d_priorities = nullptr;
    P.pair_tokens = 1;      // build_pair_tokens' kept list: open_tasks[:max_tasks]
    P.score_rows = O.score_rows > 0 ? O.score_rows : 16;
    P.score_cols = O.score_cols > 0 ? O.score_cols : 32;
    P.use_visibility = 1;
    P.d_reserved = nullptr;  // UrgencyPair.plan passes no reserved agents
  } else if (O.planner == 4) {
    P.d_priorities = nullptr;
    P.d_edge_scores = O.d_edge_scores;
    P.score_f64 = 0;
    P.d_task_order = O.d_task_order;
    P.pair_tokens = 2;
    P.use_visibility = 1;
  } else {
    P.d_priorities = nullptr;
    P.pair_tokens = 0;
    P.use_visibility = 1;
  }
  const int np = allocate_tasks(S, P, e, out_agent, out_tid, lane, nlanes, true);
  if (lane == 0 && np > 0) {
    if (O.planner == 5) {
      // UrgencyPair.plan takes no commit locks
    } else if (O.planner == 2 || O.planner == 4) {
      for (int i = 0; i < np; ++i) apply_commit(S, out_agent[i], C.commit_horizon);
    } else if (O.planner == 3) {
      // AttentionCommit._plan_from_scores (AttentionCommit.py:289-298): gate by the commit head, row = live-agent index
      uint64_t assigned = 0;
      for (int i = 0; i < np; ++i) assigned |= (uint64_t)1 << out_agent[i];
      const float* cv = O.d_plan_commit + (size_t)e * O.score_rows;
      const int horizon = C.commit_horizon != 0 ? C.commit_horizon : 25;
      int row = 0;
      for (int a = 0; a < A && row < O.score_rows; ++a) {
        if (V.a_state()[a] == -1) continue;
        const int r = row++;
        if (W.plan_reserved[a] || !((assigned >> a) & 1)) continue;
        if ((double)cv[r] >= O.commit_threshold) apply_commit(S, a, horizon);
      }
    } else {
      // UrgencyCommit lock ranking (AttentionCommit.py:334-355)
      const double thr = 1.0 - 12.0 / 40.0;
      const int n = HIv(N_TASKS);
      uint64_t assigned = 0;
      int n_free = 0;
      for (int i = 0; i < np; ++i) assigned |= (uint64_t)1 << out_agent[i];
      for (int a = 0; a < A; ++a) {
        W.plan_score[a] = -1.0;
        if (!((assigned >> a) & 1) || V.a_state()[a] == -1 || W.plan_reserved[a]) continue;
        double dmn = 0.0;
        bool any = false;
        for (int k = 0; k < n; ++k) {
          if (V.k_status()[k] == 2 || V.k_deadline()[k] < 0) continue;
          int ti = V.k_type()[k];
          if (!(V.k_alloc_ti()[k] < V.k_cur_ti()[k])) continue;
          if (!vis_none && !S.known_bit(a, k)) continue;
          if (!(task_urgency(V, k, t) >= thr)) continue;
          double d = norm2(V.a_posx()[a] - V.k_posx()[k], V.a_posy()[a] - V.k_posy()[k]);
          if (!any || d < dmn) dmn = d;
          any = true;
        }
        W.plan_score[a] = dmn + (V.a_type()[a] == UT_F2 ? 500.0 : 0.0);
        ++n_free;
      }
      int n_lock = (int)rint(O.commit_fraction * (double)(n_free > 1 ? n_free : 1));
      if (n_lock < 1) n_lock = 1;
      const int horizon = C.commit_horizon != 0 ? C.commit_horizon : 25;
      // top n_lock by (score, name) descending
      for (int r = 0; r < n_lock && r < n_free; ++r) {
        int best = -1;
        for (int a = 0; a < A; ++a) {
          if (W.plan_score[a] < 0.0) continue;
          if (best < 0 || W.plan_score[a] > W.plan_score[best] ||
              (W.plan_score[a] == W.plan_score[best] && V.a_name_rank()[a] > V.a_name_rank()[best]))
            best = a;
        }
        if (best < 0) break;
        W.plan_score[best] = -1.0;
        apply_commit(S, best, horizon);
      }
    }
  }
  MUAV_WARP_SYNC();
  return np;
}

}  // namespace muav
