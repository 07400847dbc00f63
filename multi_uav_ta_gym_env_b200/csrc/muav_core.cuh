// Per-environment simulation core: one env step (MultiUAVEnv.step, mUAV_TA/DroneEnv.py:774-1206),
// the visibility-masked Local/Coalition Hungarian allocator
// (TaskAllocation/OptimizationBased/HungarianAllocator.py:72-208) and the rectangular LSAP
// it calls (scipy.optimize.linear_sum_assignment, HungarianAllocator.py:181).
//
// The code operates on a View of ONE environment record (shared memory inside the kernel).
// All arithmetic is IEEE float64 with one rounding per written operation; the translation
// unit is compiled with -fmad=false and the single FMA NumPy's BLAS ddot performs inside
// np.linalg.norm of a 2-vector is written explicitly.
//
// Control flow is sequential per environment because the reference's semantics are
// order-dependent (dict-order actions, agent-index-order FSM with cross-agent side effects,
// data-dependent RNG draws).  Inside the kernel the sequential sections run on the warp's
// lane 0; the data-parallel loops (sensing, cost matrix, LSAP column scan) take
// (lane, nlanes) and are spread over the warp.
#pragma once
#include <math.h>
#include <stdint.h>
#include "muav_layout.h"

// Lean instantiation of the step kernel (muav_kernels.cu compiles it a second time with MUAV_LEAN): the configuration of
// most registered scenarios -- no escorts, no obstacles, the plain Hungarian allocator -- as compile-time constants, so the
// escort life cycle, obstacle avoidance and the planner front ends drop out of the instruction stream of a kernel that is
// bound by instruction supply.  The launcher picks it only when the configuration really has those values.
#if defined(MUAV_LEAN)
#if defined(MUAV_LEAN_ESCORT)   // second lean instantiation: escorts always on (WPS_escort), otherwise the same
#define MUAV_F_ESCORT(x) true
#else
#define MUAV_F_ESCORT(x) false
#endif
#define MUAV_F_NOBS(x) 0
#if defined(MUAV_LEAN_PLANNER)   // lean instantiations that keep the planner front ends and the market allocators
#define MUAV_F_PLANNER(x) (x)
#else
#define MUAV_F_PLANNER(x) 0
#endif
#else
#define MUAV_F_ESCORT(x) ((x) != 0)
#define MUAV_F_NOBS(x) (x)
#define MUAV_F_PLANNER(x) (x)
#endif

namespace muav {

#define HIv(name) V.hi()[HI_##name]
#define HFv(name) V.hf()[HF_##name]

enum { TT_HOLD = 0, TT_REC = 1, TT_ATT = 2, TT_DEF = 3, TT_INT = 4, TT_DET = 5 };
enum { UT_R1 = 0, UT_R2 = 1, UT_E1 = 2, UT_F1 = 3, UT_F2 = 4, UT_T1 = 5, UT_T2 = 6 };
enum { EV_RESET = 0, EV_FAIL = 1, EV_THREAT = 2, EV_ESC_CREATED = 3, EV_ESC_RETIRED = 4 };

// Out of line on purpose: float64 sqrt / divide expand to ~25 SASS instructions each; the kernel is
// instruction-fetch bound (profiles/r01_step_kernel_ncu.md), so all call sites share one copy.
// index of the lowest set bit (x != 0)
MUAV_HD inline int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
MUAV_HD MUAV_NOINLINE_LEAF inline double norm2(double x, double y) { return sqrt(fma(y, y, x * x)); }
MUAV_HD MUAV_NOINLINE_LEAF inline double norm2_rows(double x, double y) { return sqrt(x * x + y * y); }
MUAV_HD MUAV_NOINLINE_LEAF inline double ddiv(double a, double b) { return a / b; }
MUAV_HD inline double dmax(double a, double b) { return a > b ? a : b; }
MUAV_HD inline double dmin(double a, double b) { return a < b ? a : b; }
MUAV_HD inline bool is_fighter(int ut) { return ut == UT_F1 || ut == UT_F2; }
MUAV_HD inline bool is_recon(int ut) { return ut == UT_R1 || ut == UT_R2; }

#if defined(__CUDA_ARCH__)
#define MUAV_WARP_SYNC() __syncwarp()
// Phase alignment of the warps (= environments) of one CTA: the kernel is instruction-fetch bound, so
// warps that run the same phase at the same time share the fetched lines.  Purely a scheduling hint:
// no data is exchanged between the warps.
#define MUAV_CTA_SYNC(on) \
  do {                    \
    if (on) __syncthreads(); \
  } while (0)
#else
#define MUAV_WARP_SYNC() ((void)0)
#define MUAV_CTA_SYNC(on) ((void)0)
#endif

#if defined(MUAV_PHASE_TIMING) && defined(__CUDA_ARCH__)
#define MUAV_TICK(slot)                                   \
  do {                                                    \
    long long _now = clock64();                           \
    if (phase_cycles) phase_cycles[slot] += _now - _tick; \
    _tick = _now;                                         \
  } while (0)
#define MUAV_TICK_START() long long _tick = clock64()
#else
#define MUAV_TICK(slot) ((void)0)
#define MUAV_TICK_START() ((void)0)
#endif

struct StepResult {
  double reward;
  int terminated, truncated;
};

struct Sim {
  View V;
  const muav_config* Cp;
  const uint32_t* tape;  // this env's tapes: [agent | tgt | mission]
  char* scratch;
  int32_t* out_events;   // drained events of this step (may be null)
  int n_out_events;
  double step_reward;
  long long* phase_cycles;  // MUAV_PHASE_TIMING only: per-warp cycle sums by phase

  MUAV_HD const muav_config& C() const { return *Cp; }
  MUAV_HD int A() const { return V.lay().D.A; }
  MUAV_HD int QC() const { return V.lay().D.QC; }

  // ------------------------------------------------------------------ RNG (oracle/rng.py rules)
  MUAV_HD uint32_t rng_word(int stream) {
    int32_t* cur = &V.hi()[HI_CUR_AGENT + stream];
    int off = 0;
    for (int s = 0; s < stream; ++s) off += C().tape_words[s];
    if (*cur >= C().tape_words[stream]) {
      HIv(ERRFLAGS) |= ERR_TAPE_OVERFLOW;
      return 0u;
    }
    uint32_t w = tape[off + *cur];
    *cur += 1;
    return w;
  }
  MUAV_HD double rng_random(int stream) {
    uint32_t a = rng_word(stream) >> 5;
    uint32_t b = rng_word(stream) >> 6;
    return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
  }
  MUAV_HD double rng_uniform(int stream, double a, double b) { return a + (b - a) * rng_random(stream); }
  MUAV_HD int rng_below(int stream, int n) {
    int k = 0;
    while ((n >> k) != 0) ++k;
    uint32_t r = rng_word(stream) >> (32 - k);
    int guard = 0;
    while ((int)r >= n && guard++ < 64) r = rng_word(stream) >> (32 - k);
    return (int)r;
  }

  // ------------------------------------------------------------------ small accessors
  MUAV_HD double cap(int a, int c) const { return V.a_caps()[c * A() + a]; }
  MUAV_HD void set_cap(int a, int c, double v) { V.a_caps()[c * A() + a] = v; }
  MUAV_HD double speed_of(int a) const { return C().speed[V.a_type()[a]]; }
  MUAV_HD double engage_of(int a) const { return C().engage[V.a_type()[a]]; }
  MUAV_HD int qlen(int a) const { return V.a_qlen()[a]; }
  MUAV_HD int qat(int a, int s) const { return V.a_queue()[s * A() + a]; }
  MUAV_HD int qhead(int a) const { return V.a_qlen()[a] > 0 ? V.a_queue()[a] : 0; }
  MUAV_HD int qfind(int a, int tid) const {
    int n = V.a_qlen()[a];
    _Pragma("unroll 1") for (int s = 0; s < n; ++s)
      if (V.a_queue()[s * A() + a] == tid) return s;
    return -1;
  }
  MUAV_HD double qremove(int a, int slot) {
    int n = V.a_qlen()[a];
    int16_t* q = V.a_queue();
    double* qt = V.a_qtime();
    int Aa = A();
    double t0 = qt[slot * Aa + a];
    _Pragma("unroll 1") for (int s = slot; s + 1 < n; ++s) {
      q[s * Aa + a] = q[(s + 1) * Aa + a];
      qt[s * Aa + a] = qt[(s + 1) * Aa + a];
    }
    V.a_qlen()[a] = (int16_t)(n - 1);
    return t0;
  }
  MUAV_HD void qpush(int a, int tid, double time_at) {
    int n = V.a_qlen()[a];
    if (n >= QC()) {
      HIv(ERRFLAGS) |= ERR_QUEUE_OVERFLOW;
      return;
    }
    V.a_queue()[n * A() + a] = (int16_t)tid;
    V.a_qtime()[n * A() + a] = time_at;
    V.a_qlen()[a] = (int16_t)(n + 1);
  }
  MUAV_HD bool known_bit(int a, int k) const { return (V.known()[(k >> 5) * A() + a] >> (k & 31)) & 1u; }
  MUAV_HD void set_known(int a, int k) { V.known()[(k >> 5) * A() + a] |= (1u << (k & 31)); }
  MUAV_HD void push_event(int tag, int arg) {
    int n = HIv(N_EVENTS);
    if (n >= V.lay().D.EVC) {
      HIv(ERRFLAGS) |= ERR_EVENT_OVERFLOW;
      return;
    }
    V.events()[n] = ((arg + 1) << 8) | tag;
    HIv(N_EVENTS) = n + 1;
  }
  // allocationDetails[k] is represented by the agents whose queue holds task k+1
  MUAV_HD int details_count(int tid) const {
    int c = 0;
    _Pragma("unroll 1") for (int a = 0; a < A(); ++a)
      if (qfind(a, tid) >= 0) ++c;
    return c;
  }

  // A closed task keeps whatever allocationDetails entries it had when it closed (removeAgentCap is a no-op once
  // status == 2, DroneEnvComponents.py:282-285); only their number is observable (the replay's assigned_agents).
  MUAV_HD void close_task(int k) {
    if (V.k_status()[k] != 2) V.k_det_frozen()[k] = (int16_t)details_count(k + 1);  // some sites re-close every step
    V.k_status()[k] = 2;
  }

  // hot copies of component [task type] of the requirement vectors (muav_layout.h): refreshed after every update
  MUAV_HD void sync_req(int k) {
    const int ti = V.k_type()[k];
    V.k_cur_ti()[k] = V.k_cur2(ti, k);
    V.k_alloc_ti()[k] = V.k_alloc2(ti, k);
  }

  // ------------------------------------------------------------------ Task.add/removeAgentCap
  // DroneEnvComponents.py:280-301.  `t0` is the time stored with the entry that was just removed.
  MUAV_HD MUAV_NI_H void remove_agent_cap(int k, int a, double t0) {
    if (V.k_status()[k] == 2) return;
    int TC = V.lay().D.TC;
    _Pragma("unroll 1") for (int c = 0; c < 6; ++c) V.k_alloc2(c, k) = V.k_alloc2(c, k) - cap(a, c);
    sync_req(k);
    int tid = k + 1;
    int cnt = 0;
    double mn = 0.0, mx = 0.0;
    _Pragma("unroll 1") for (int b = 0; b < A(); ++b) {
      int s = qfind(b, tid);
      if (s >= 0) {
        double tm = V.a_qtime()[s * A() + b];
        if (cnt == 0) {
          mn = mx = tm;
        } else {
          if (tm < mn) mn = tm;
          if (tm > mx) mx = tm;
        }
        ++cnt;
      }
    }
    double dur = (double)C().duration[V.k_type()[k]];
    if (cnt > 0) {
      if (t0 == V.k_init()[k]) V.k_init()[k] = mn;
      if (t0 + dur == V.k_dtime()[k]) V.k_dtime()[k] = mx + dur;
    } else {
      V.k_init()[k] = -1.0;
      V.k_dtime()[k] = -1.0;
    }
  }
  // DroneEnvComponents.py:306-326 (the queue entry itself is pushed by the caller)
  MUAV_HD MUAV_NI_H void add_agent_cap(int k, int a, double time_at) {
    if (V.k_status()[k] == 2) return;
    int TC = V.lay().D.TC;
    double end = time_at + (double)C().duration[V.k_type()[k]];
    _Pragma("unroll 1") for (int c = 0; c < 6; ++c) V.k_alloc2(c, k) = V.k_alloc2(c, k) + cap(a, c);
    sync_req(k);
    if (time_at < V.k_init()[k] || V.k_init()[k] == -1.0) {
      V.k_init()[k] = time_at;
      if (V.k_dtime()[k] == -1.0) V.k_dtime()[k] = end;
    }
    if (end > V.k_dtime()[k]) V.k_dtime()[k] = end;
    V.k_status()[k] = 1;
  }

  // ------------------------------------------------------------------ UAV methods
  // UAV.allocate (DroneEnvComponents.py:55-95), task.id != 0
  MUAV_HD MUAV_NI_H bool allocate(int a, int tid) {
    int k = tid - 1;
    if (qfind(a, tid) >= 0 || V.k_status()[k] == 2) return false;
    V.a_re_eval()[a] = 0;
    V.a_last_task()[a] = -1;
    double t = (double)HIv(T);
    double time_to = ddiv(norm2(V.a_nfpx()[a] - V.k_posx()[k], V.a_nfpy()[a] - V.k_posy()[k]), speed_of(a));
    double start = (V.a_nft()[a] - t) > 0 ? V.a_nft()[a] : t;
    double end = start + time_to + (double)C().duration[V.k_type()[k]];
    if (qlen(a) == 0) {
      V.a_task_start()[a] = -1;
      V.a_state()[a] = 1;
    }
    qpush(a, tid, time_to);
    V.a_nft()[a] = end;
    V.a_nfpx()[a] = V.k_posx()[k];
    V.a_nfpy()[a] = V.k_posy()[k];
    add_agent_cap(k, a, time_to);
    return true;
  }
  // UAV.desAllocate (DroneEnvComponents.py:97-113)
  MUAV_HD MUAV_NI_H bool des_allocate(int a, int tid) {
    if (tid <= 0) return false;
    int s = qfind(a, tid);
    if (s < 0) return false;
    double t0 = qremove(a, s);
    V.a_nft()[a] = (double)HIv(T);
    V.a_nfpx()[a] = V.a_posx()[a];
    V.a_nfpy()[a] = V.a_posy()[a];
    V.a_commit()[a] = 0;
    remove_agent_cap(tid - 1, a, t0);
    return true;
  }
  // UAV.desallocateAll (DroneEnvComponents.py:115-119): the reference iterates the list it mutates,
  // so only the entries at even positions are removed.
  MUAV_HD MUAV_NI_H void des_allocate_all(int a) {
    int i = 0;
    _Pragma("unroll 1") while (i < qlen(a)) {
      des_allocate(a, qat(a, i));
      ++i;
    }
    V.a_commit()[a] = 0;
  }
  // UAV.outOfService (DroneEnvComponents.py:122-127)
  MUAV_HD MUAV_NOINLINE void out_of_service(int a) {
    V.a_state()[a] = -1;
    V.a_commit()[a] = 0;
    int i = 0;
    while (i < qlen(a)) {
      des_allocate(a, qat(a, i));
      ++i;
    }
  }
  // EnvUtils.desallocateAll (MultiDroneEnvUtils.py:183-205), single-task mode
  MUAV_HD MUAV_NOINLINE void env_desallocate_all(int a) {
    while (qlen(a) > 0) {
      int tid = qat(a, 0);
      des_allocate(a, tid);
      int k = tid - 1;
      if (a < 32) V.k_tbl_lo()[k] &= ~(1u << a);
      else V.k_tbl_hi()[k] &= ~(1u << (a - 32));
    }
    V.a_nft()[a] = (double)HIv(T);
    V.a_nfpx()[a] = V.a_posx()[a];
    V.a_nfpy()[a] = V.a_posy()[a];
  }
  // UAV.taskDone (DroneEnvComponents.py:143-179); *t0 receives the popped entry's allocation time
  MUAV_HD MUAV_NI_H bool task_done(int a, int tid, double* t0) {
    if (qlen(a) == 0 || qat(a, 0) != tid) return false;
    *t0 = qremove(a, 0);
    V.a_task_start()[a] = -1;
    int k = tid - 1;
    if (V.k_type()[k] == TT_ATT) {
      V.a_ammo()[a] -= 1;
      if (V.a_ammo()[a] <= 0) set_cap(a, TT_ATT, 0.0);
    }
    while (qlen(a) > 0 && V.k_status()[qat(a, 0) - 1] == 2) qremove(a, 0);
    if (qlen(a) == 0) {
      if (V.a_re_eval()[a]) {
        V.a_last_task()[a] = -1;
        V.a_re_eval()[a] = 0;
      }
      V.a_nft()[a] = 0.0;
      V.a_nfpx()[a] = V.a_posx()[a];
      V.a_nfpy()[a] = V.a_posy()[a];
      V.a_state()[a] = 0;
    } else {
      V.a_state()[a] = 1;
    }
    return true;
  }
  // _is_task_action_valid (DroneEnv.py:341-363)
  MUAV_HD bool is_valid(int a, int tid) const {
    int k = tid - 1;
    if (V.k_status()[k] == 2) return false;
    if (qlen(a) > 0 && qat(a, 0) == tid) return true;
    int el = V.k_elig()[k];
    if (el != 0 && !((el >> V.a_type()[a]) & 1)) return false;
    int ti = V.k_type()[k];
    int TC = V.lay().D.TC;
    if (C().capability_mask && cap(a, ti) <= 0) return false;
    if (C().saturate_mask && V.k_alloc_ti()[k] >= V.k_org_ti()[k]) return false;
    return true;
  }

  // ------------------------------------------------------------------ WPS bookkeeping
  // _wps_mark_window_outcome (DroneEnv.py:1543-1555)
  MUAV_HD MUAV_NI_H void mark_outcome(int k, bool success) {
    if (V.k_deadline()[k] < 0 || V.k_counted()[k]) return;
    V.k_counted()[k] = 1;
    if (success && HIv(T) <= V.k_deadline()[k]) {
      HIv(N_ON_TIME) += 1;
      HFv(F_REWARD) += C().on_time_bonus;
    } else {
      HIv(N_MISSED) += 1;
      HFv(F_REWARD) -= C().miss_penalty;
    }
  }
  // Task ctor + env.tasks.append (DroneEnvComponents.py:224-263); returns task id or 0 on overflow
  MUAV_HD MUAV_NOINLINE int new_task(double px, double py, int ti) {
    int k = HIv(N_TASKS);
    const int TC = V.lay().D.TC, IC = V.lay().D.IC;
    if (k >= IC) {
      HIv(ERRFLAGS) |= ERR_TASK_OVERFLOW;
      return 0;
    }
    int slot = -1;
    _Pragma("unroll 1") for (int sidx = 0; sidx < TC; ++sidx)
      if (V.s_used()[sidx] == 0) { slot = sidx; break; }
    if (slot < 0) {
      HIv(ERRFLAGS) |= ERR_TASK_OVERFLOW;
      return 0;
    }
    HIv(N_TASKS) = k + 1;
    HIv(N_SLOTS_USED) += 1;
    V.s_used()[slot] = (int16_t)(k + 1);
    V.k_slot()[k] = (int16_t)slot;
    V.k_posx()[k] = px;
    V.k_posy()[k] = py;
    V.k_type()[k] = (int16_t)ti;
    V.k_status()[k] = 0;
    for (int c = 0; c < 6; ++c) {
      V.k_cur2(c, k) = 0.0;
      V.k_alloc2(c, k) = 0.0;
    }
    V.k_cur_ti()[k] = 0.0;
    V.k_alloc_ti()[k] = 0.0;
    V.k_done_ti()[k] = 0.0;
    V.k_org_ti()[k] = 0.0;
    V.k_init()[k] = -1.0;
    V.k_dtime()[k] = -1.0;
    V.k_created()[k] = 0;
    V.k_deadline()[k] = -1;
    V.k_counted()[k] = 0;
    V.k_fq()[k] = -1;
    V.k_kind()[k] = 0;
    V.k_req_agents()[k] = 0;
    V.k_elig()[k] = 0;
    V.k_threat()[k] = -1;
    V.k_prot_agent()[k] = -1;
    V.k_prot_task()[k] = 0;
    V.k_reveal()[k] = -1;
    V.k_tbl_lo()[k] = 0;
    V.k_tbl_hi()[k] = 0;
    V.k_reached()[k] = 0;
    V.k_det_frozen()[k] = 0;
    return k + 1;
  }
  // A closed task keeps its slot only while something still refers to it: an agent queue or last_task (the
  // switch penalty reads its type and position, DroneEnv.py:852,859), the escort map (_sync_escorts keeps visiting
  // stale entries, :1977-2000) or a threat that is not destroyed (update_threats keeps writing its position, :1740).
  // Everything else about a closed task is dead data in the reference (closed tasks never reopen, :1460), so the
  // slot is recycled.  Spread over the lanes: lane <-> slot.
  MUAV_HD void free_dead_tasks(int lane, int nlanes) {
    const int TC = V.lay().D.TC, Aa = A();
    for (int sidx = lane; sidx < TC; sidx += nlanes) {
      const int tid = V.s_used()[sidx];
      if (tid == 0 || V.k_status()[tid - 1] != 2) continue;
      bool ref = false;
      for (int a = 0; a < Aa && !ref; ++a) {
        if (V.a_last_task()[a] == tid || V.a_escort()[a] == tid) ref = true;
        const int n = V.a_qlen()[a];
        for (int q = 0; q < n && !ref; ++q) ref = V.a_queue()[q * Aa + a] == tid;
      }
      const int na = V.hi()[HI_N_ACTIVE];
      for (int i = 0; i < na && !ref; ++i) {
        const int hid = V.h_order()[i];
        ref = V.h_status()[hid] != 2 && V.h_task()[hid] == tid;
      }
      if (ref) continue;
      if (V.k_tbl_lo_raw()[sidx] == 0 && V.k_tbl_hi_raw()[sidx] == 0) {
#if defined(__CUDA_ARCH__)
        atomicAdd(&V.hi()[HI_N_FREED_EMPTY_TBL], 1);
#else
        V.hi()[HI_N_FREED_EMPTY_TBL] += 1;
#endif
      }
#if defined(__CUDA_ARCH__)
      atomicSub(&V.hi()[HI_N_SLOTS_USED], 1);
#else
      V.hi()[HI_N_SLOTS_USED] -= 1;
#endif
      V.s_used()[sidx] = 0;
      V.k_slot()[tid - 1] = -1;
    }
  }
  // _register_dynamic_task (DroneEnv.py:1491-1504)
  MUAV_HD MUAV_NI_H void register_dynamic(int tid) {
    int k = tid - 1;
    if (C().hard_windows && V.k_deadline()[k] < 0) {
      V.k_deadline()[k] = (int16_t)(HIv(T) + C().window_length);
      HIv(N_WINDOWED) += 1;
    }
    if (C().threat_delay > 0 || C().sense_radius > 0) {
      int d = C().threat_delay > 0 ? C().threat_delay : 0;
      V.k_reveal()[k] = (int16_t)(HIv(T) + d);
    } else {
      for (int a = 0; a < A(); ++a) set_known(a, k);
    }
  }
  // _counts_for_mission_done (DroneEnv.py:1878-1886)
  MUAV_HD MUAV_NI_H bool all_done() const {
    int n = V.hi()[HI_N_TASKS];
    for (int k = 0; k < n; ++k) {
      if (V.k_status()[k] == 2) continue;  // closed tasks never block (and may have given their slot back)
      int ti = V.k_type()[k];
      if (V.k_kind()[k] == 1 || ti == TT_DET || ti == TT_HOLD) continue;
      return false;
    }
    return true;
  }
  MUAV_HD void mark_reached(int k) {
    if (!V.k_reached()[k]) {
      V.k_reached()[k] = 1;
      HIv(N_REACHED) += 1;
    }
  }

  // releaseAllTasks (DroneEnv.py:1442-1480); for_type == -1 indexes the Det column
  MUAV_HD MUAV_NOINLINE void release_all(int for_type) {
    int col = for_type < 0 ? 6 + for_type : for_type;
    uint32_t avail = 0;
    _Pragma("unroll 1") for (int a = 0; a < A(); ++a) {
      if (cap(a, col) > 0) {
        if (V.a_state()[a] != -1) {
          V.a_re_eval()[a] = 1;
          V.a_last_task()[a] = qhead(a);
          des_allocate_all(a);
          avail |= 1u << V.a_type()[a];
        }
      }
    }
    // can any remaining agent type still serve this task type?  (the same answer for every task of the type)
    bool any = false;
    if (for_type >= 0)
      for (int ut = 0; ut < MUAV_N_UAV_TYPES; ++ut)
        if (((avail >> ut) & 1u) && C().cap_table[ut][for_type] != 0.0) any = true;
    // only the tasks that were open at the last scan can be open now (events are drained first thing in a step)
    const int KWn = (HIv(N_TASKS) + 31) >> 5;
    for (int wd = 0; wd < KWn; ++wd)
    for (uint32_t bits = V.open_mask()[wd]; bits; bits &= bits - 1) {
      const int k = (wd << 5) + ctz32(bits);
      if (V.k_status()[k] != 2 && V.k_type()[k] == for_type) {
        if (!any) {
          close_task(k);
          if (!V.k_reached()[k]) {
            V.k_reached()[k] = 1;
            HIv(N_REACHED) += 1;
            if (HIv(N_REACHED) == C().n_tasks_cfg) HIv(CONCLUSION) = HIv(T);
          }
        } else {
          V.k_status()[k] = 0;
          V.k_tbl_lo()[k] = 0;
          V.k_tbl_hi()[k] = 0;
        }
      }
    }
  }

  // get_closest_agent (DroneEnv.py:1691-1723)
  MUAV_HD MUAV_NI_H int closest_agent(double px, double py) const {
    double min_f = INFINITY, min_w = INFINITY;
    int cf = -1, cw = -1;
    _Pragma("unroll 1") for (int a = 0; a < A(); ++a) {
      int st = V.a_state()[a];
      if (st != -1 && st != 4) {
        double d = norm2(V.a_posx()[a] - px, V.a_posy()[a] - py);
        if (is_fighter(V.a_type()[a])) {
          if (d < min_f) { min_f = d; cf = a; }
        } else {
          if (d < min_w) { min_w = d; cw = a; }
        }
      }
    }
    return cw != -1 ? cw : cf;
  }

  // generate_threat + TaskFromThreat (DroneEnv.py:1601-1643,1861-1876)
  MUAV_HD MUAV_NI_H void generate_threat() {
    int t = HIv(T);
    int TC = V.lay().D.TC;
    for (int g = 0; g < C().n_groups; ++g) {
      int* gnext = &V.hi()[HI_GROUP_NEXT0 + g];
      int gend = C().group_start[g + 1];
      int left = gend - *gnext;
      if (left > 0 && t > 40 && t % 10 == 0) {
        if (rng_random(0) < C().threat_gen_prob) {
          int n_spawn = 1;
          if (C().burst_mode) n_spawn = C().burst_size < left ? C().burst_size : left;
          for (int bi = 0; bi < n_spawn; ++bi) {
            if (*gnext >= gend) break;
            int hid = *gnext;
            *gnext += 1;
            if (C().dual_region_bursts) {
              double mid = C().area_w * 0.5;
              double wide = dmax(C().threat_wide, 40.0);
              double x;
              if ((HIv(BURST_TOGGLE) + bi) % 2 == 0) x = rng_uniform(0, wide, mid - wide);
              else x = rng_uniform(0, mid + wide, C().area_w - wide);
              V.h_posx()[hid] = x;
            }
            double hx = V.h_posx()[hid], hy = V.h_posy()[hid];
            int tgt = closest_agent(hx, hy);
            V.h_target()[hid] = (int16_t)tgt;
            V.h_mission()[hid] = (int16_t)tgt;
            int tid = new_task(hx, hy, TT_INT);
            if (tid == 0) return;
            int k = tid - 1;
            int ht = V.h_type()[hid];
            V.k_cur2(TT_INT, k) = 2.0;
            V.k_cur2(TT_ATT, k) = C().cap_table[ht][3] * 2;
            V.k_cur2(TT_DEF, k) = C().cap_table[ht][2] * 2;
            sync_req(k);
            V.k_org_ti()[k] = 2.0;
            V.k_threat()[k] = (int16_t)hid;
            V.k_created()[k] = (int16_t)t;
            if (ht == UT_T1) {
              V.k_req_agents()[k] = 2;
              V.k_elig()[k] = (int16_t)C().escort_type_mask;
            }
            V.h_task()[hid] = (int16_t)tid;
            V.h_spawned()[hid] = 1;
            V.h_order()[HIv(N_ACTIVE)] = (int16_t)hid;
            HIv(N_ACTIVE) += 1;
            int dk = V.h_det_task()[hid] - 1;
            V.k_cur2(TT_DET, dk) = V.k_cur2(TT_DET, dk) - 1.0;
            sync_req(dk);
            register_dynamic(tid);
            push_event(EV_THREAT, tid);
            push_event(EV_RESET, TT_INT);
            HIv(PENDING_RESET) = 1;
          }
          if (C().dual_region_bursts && n_spawn > 0) HIv(BURST_TOGGLE) = (HIv(BURST_TOGGLE) + 1) % 2;
        }
      }
    }
  }

  // _escort_fighters_near (DroneEnv.py:1746-1764): ids sorted by distance (stable) into out[], returns count
  MUAV_HD MUAV_NOINLINE int fighters_near(int prot, double radius, int16_t* out, double* dtmp) const {
    if (prot < 0) return 0;
    int esc = V.a_escort()[prot];
    if (esc == 0 || V.k_status()[esc - 1] == 2) return 0;
    double px = V.a_posx()[prot], py = V.a_posy()[prot];
    int n = 0;
    _Pragma("unroll 1") for (int a = 0; a < A(); ++a) {
      if (V.a_state()[a] == -1 || !((C().escort_type_mask >> V.a_type()[a]) & 1)) continue;
      if (qlen(a) == 0 || qat(a, 0) != esc) continue;
      double d = norm2(V.a_posx()[a] - px, V.a_posy()[a] - py);
      if (d <= radius) {
        int j = n;
        while (j > 0 && dtmp[j - 1] > d) {
          dtmp[j] = dtmp[j - 1];
          out[j] = out[j - 1];
          --j;
        }
        dtmp[j] = d;
        out[j] = (int16_t)a;
        ++n;
      }
    }
    return n;
  }
  // first element of _escort_fighters_near's stable distance sort (DroneEnv.py:1746-1764), or -1 when the list is empty
  MUAV_HD int nearest_escort_fighter(int prot, double radius) const {
    if (prot < 0) return -1;
    const int esc = V.a_escort()[prot];
    if (esc == 0 || V.k_status()[esc - 1] == 2) return -1;
    const double px = V.a_posx()[prot], py = V.a_posy()[prot];
    int best = -1;
    double bd = 0.0;
    _Pragma("unroll 1") for (int a = 0; a < A(); ++a) {
      if (V.a_state()[a] == -1 || !((C().escort_type_mask >> V.a_type()[a]) & 1)) continue;
      if (qlen(a) == 0 || qat(a, 0) != esc) continue;
      const double d = norm2(V.a_posx()[a] - px, V.a_posy()[a] - py);
      if (d <= radius && (best < 0 || d < bd)) { best = a; bd = d; }
    }
    return best;
  }
  MUAV_HD double* near_d() const { return (double*)scratch + 3 * V.lay().D.A; }
  MUAV_HD int16_t* near_i() const { return (int16_t*)((double*)scratch + 4 * V.lay().D.A); }

  // _retarget_threat_via_escort (DroneEnv.py:1766-1779)
  MUAV_HD void retarget_via_escort(int hid) {
    int mission = V.h_mission()[hid] >= 0 ? V.h_mission()[hid] : V.h_target()[hid];
    if (mission < 0 || V.a_state()[mission] == -1) return;
    if (!is_recon(V.a_type()[mission])) return;
    int n = fighters_near(mission, C().escort_intercept_radius, near_i(), near_d());
    if (n == 0) {
      V.h_target()[hid] = (int16_t)mission;
      V.h_intercept()[hid] = -1;
      return;
    }
    V.h_target()[hid] = near_i()[0];
    V.h_intercept()[hid] = near_i()[0];  // Threat.intercepting_agent: read only by the replay (DroneEnv.py:1775-1791)
  }
  // _release_escort_agents (DroneEnv.py:1919-1936)
  MUAV_HD MUAV_NOINLINE void release_escort_agents(int esc) {
_Pragma("unroll 1") for (int a = 0; a < A(); ++a) {
      if (V.a_state()[a] == -1) continue;
      if (qfind(a, esc) >= 0) {
        des_allocate(a, esc);
        if (qlen(a) == 0) {
          V.a_state()[a] = 0;
          V.a_commit()[a] = 0;
          V.a_nft()[a] = (double)HIv(T);
          V.a_nfpx()[a] = V.a_posx()[a];
          V.a_nfpy()[a] = V.a_posy()[a];
        }
      }
    }
  }
  // _retire_escort (DroneEnv.py:1938-1950)
  MUAV_HD MUAV_NOINLINE void retire_escort(int esc, bool failed) {
    if (esc == 0 || V.k_status()[esc - 1] == 2) return;
    release_escort_agents(esc);
    close_task(esc - 1);
    int recon = V.k_prot_agent()[esc - 1];
    if (recon >= 0) V.a_escort()[recon] = 0;
    if (failed) HIv(ESC_FAILED) += 1;
    else HIv(ESC_COMPLETED) += 1;
    push_event(EV_ESC_RETIRED, esc);
  }
  // _create_escort_for (DroneEnv.py:1888-1917)
  MUAV_HD MUAV_NOINLINE void create_escort_for(int a, int rec_tid) {
    if (!MUAV_F_ESCORT(C().escort_enabled)) return;
    if (V.a_escort()[a] != 0) return;
    int tid = new_task(V.a_posx()[a], V.a_posy()[a], TT_DEF);
    if (tid == 0) return;
    int k = tid - 1;
    int TC = V.lay().D.TC;
    V.k_cur2(TT_DEF, k) = C().escort_requirement;
    sync_req(k);
    V.k_org_ti()[k] = C().escort_requirement;
    V.k_kind()[k] = 1;
    V.k_prot_agent()[k] = (int16_t)a;
    V.k_prot_task()[k] = (int16_t)rec_tid;
    V.k_elig()[k] = (int16_t)C().escort_type_mask;
    V.k_req_agents()[k] = (int16_t)C().escort_required_agents;
    V.k_created()[k] = (int16_t)HIv(T);
    register_dynamic(tid);
    V.a_escort()[a] = tid;
    HIv(ESC_REQUESTS) += 1;
    push_event(EV_ESC_CREATED, tid);
    push_event(EV_RESET, TT_DEF);
    HIv(PENDING_RESET) = 1;
  }

  // handle_threat_engagement (DroneEnv.py:1781-1858)
  MUAV_HD MUAV_NOINLINE void engage(int hid) {
    int nd = 0;
    int primary = V.h_target()[hid];
    int mission = V.h_mission()[hid] >= 0 ? V.h_mission()[hid] : primary;
    int16_t* defenders = near_i();
    if (MUAV_F_ESCORT(C().escort_enabled) && mission >= 0 && is_recon(V.a_type()[mission])) {
      nd = fighters_near(mission, C().mutual_support_radius, defenders, near_d());
      if (nd > 0) {
        primary = defenders[0];
        V.h_target()[hid] = (int16_t)primary;
        V.h_intercept()[hid] = (int16_t)primary;
      }
    }
    if (primary < 0) return;
    int ht = V.h_type()[hid];
    double att = C().cap_table[ht][2], dfn = C().cap_table[ht][3], erng = C().engage[ht];
    double att_d, def_d, eng_d;
    if (nd >= 2) {
      HIv(MUTUAL) += 1;
      double att_sum = 0.0, def_sum = 0.0, eng_sum = 0.0;
      for (int i = 0; i < nd; ++i) att_sum = att_sum + cap(defenders[i], 2);
      for (int i = 0; i < nd; ++i) def_sum = def_sum + cap(defenders[i], 3);
      for (int i = 0; i < nd; ++i) eng_sum = eng_sum + engage_of(defenders[i]);
      eng_sum = eng_sum / (double)nd;
      att_d = att_sum / dmax(att, 1e-6);
      def_d = def_sum / dmax(dfn, 1e-6);
      eng_d = eng_sum / dmax(erng, 1e-6);
    } else {
      att_d = cap(primary, 2) / dmax(att, 1e-6);
      def_d = cap(primary, 3) / dmax(dfn, 1e-6);
      eng_d = engage_of(primary) / dmax(erng, 1e-6);
    }
    double avg = (att_d + def_d + eng_d) / 3;
    double prob = avg / (avg + 1);
    double rnd = rng_random(0);
    int tid = V.h_task()[hid];
    int k = tid - 1;
    if (rnd < prob) {
      V.h_status()[hid] = 2;
      close_task(k);
      mark_outcome(k, true);
      HIv(INTERCEPTED) += 1;
      V.a_ammo()[primary] -= 1;
      if (V.a_ammo()[primary] <= 0) set_cap(primary, 3, 0.0);
      if (qlen(primary) > 0 && qat(primary, 0) == tid) {
        double t0;
        task_done(primary, tid, &t0);
      }
      step_reward += 1.0;
    } else {
      V.h_ammo()[hid] -= 1;
      V.a_ammo()[primary] -= 1;
      if (V.a_ammo()[primary] <= 0) {
        set_cap(primary, 3, 0.0);
        int pt = V.a_type()[primary];
        bool was_recon = is_recon(pt);
        bool was_escort = (C().escort_type_mask >> pt) & 1;
        out_of_service(primary);
        if (was_recon) {
          HIv(RECON_LOSSES) += 1;
          HIv(BREACHES) += 1;
          retire_escort(V.a_escort()[primary], true);
        } else if (was_escort) {
          HIv(ESCORT_LOSSES) += 1;
        }
        step_reward -= 1.0;
      }
      if (V.h_ammo()[hid] <= 0) {
        V.h_status()[hid] = 0;
        close_task(k);
        mark_outcome(k, false);
      } else {
        int tgt = closest_agent(V.h_posx()[hid], V.h_posy()[hid]);
        V.h_target()[hid] = (int16_t)tgt;
        V.h_mission()[hid] = (int16_t)tgt;
      }
    }
  }

  // update_threats (DroneEnv.py:1725-1744), sequential form
  MUAV_HD void update_threats() {
    int n = HIv(N_ACTIVE);
    for (int i = 0; i < n; ++i) update_threat(i);
  }

  // ------------------------------------------------------------------ threat-per-lane pursuit
  // One iteration of update_threats moves the threat (drift, or pursuit of its target agent) and copies the position to
  // its task; only two things reach beyond the threat itself: an engagement (RNG draw, ammunition, kills, DroneEnv.py:
  // 1781-1858) and the window outcome when it leaves the area at y <= 0.  threat_plan(i) computes the move without
  // writing; the moves of all threats commit one per lane, the engagements / exits then run in spawn order on one lane
  // (update_threat's own arithmetic: one norm2 + two divisions for the pursuit, one norm2 for the range test).
  // Without escorts a move reads nothing an engagement writes (target, status and position are the threat's own, agents
  // do not move in this phase), so one round covers all threats; with escorts _retarget_threat_via_escort reads agent
  // and escort-task state, so the threats behind an engagement are planned again after it.
#define MUAV_MAX_THREATS_HOST 512
  struct ThreatPlan {
    double hx, hy;
    int hid, k;
    int16_t target, intercept;
    bool live, engage, exit_area, retarget;
  };

  MUAV_HD ThreatPlan threat_plan(int i, bool lane_on = true) {
    ThreatPlan P;
    P.live = P.engage = P.exit_area = P.retarget = false;
    const int n = V.hi()[HI_N_ACTIVE];
    const bool in_range = lane_on && i < n;
    const int hid = in_range ? (int)V.h_order()[i] : 0;
    P.hid = hid;
    const bool live = in_range && V.h_status()[hid] != 2;
    P.live = live;
    const int ht = V.lay().D.HC > 0 ? (int)V.h_type()[hid] : UT_T1;
    const double sp = C().speed[ht];
    double hx = V.lay().D.HC > 0 ? V.h_posx()[hid] : 0.0, hy = V.lay().D.HC > 0 ? V.h_posy()[hid] : 0.0;
    int tg = V.lay().D.HC > 0 ? (int)V.h_target()[hid] : -1;
    P.target = (int16_t)tg;
    P.intercept = V.lay().D.HC > 0 ? V.h_intercept()[hid] : (int16_t)-1;
    const bool drift = !live || V.h_status()[hid] == 0 || tg < 0;
    if (MUAV_F_ESCORT(C().escort_enabled) && live && !drift) {
      // _retarget_threat_via_escort (DroneEnv.py:1766-1779) on a private copy of (target, intercepting agent)
      const int mission = V.h_mission()[hid] >= 0 ? (int)V.h_mission()[hid] : tg;
      if (mission >= 0 && V.a_state()[mission] != -1 && is_recon(V.a_type()[mission])) {
        const int nf = nearest_escort_fighter(mission, C().escort_intercept_radius);
        P.retarget = true;
        if (nf < 0) { P.target = (int16_t)mission; P.intercept = -1; }
        else { P.target = (int16_t)nf; P.intercept = (int16_t)nf; }
        tg = P.target;
      }
    }
    const int tga = (!drift && tg >= 0) ? tg : 0;
    const double ax = V.a_posx()[tga], ay = V.a_posy()[tga];
    const double dx = ax - hx, dy = ay - hy;
    const double mag = norm2(dx, dy);
    const bool mag_ok = !drift && mag != 0;
    const double den = mag_ok ? mag : 1.0;
    const double qx = ddiv(dx, den), qy = ddiv(dy, den);
    double nx = mag_ok ? qx : 0.0, ny = mag_ok ? qy : 0.0;
    if (drift) { nx = 0.0; ny = -1.0; }
    hx = hx + sp * nx;
    hy = hy + sp * ny;
    const double dist = norm2(ax - hx, ay - hy);
    P.hx = hx;
    P.hy = hy;
    P.engage = live && !drift && dist < C().engage[ht];
    P.k = live ? V.h_task()[hid] - 1 : 0;
    // leaving the area re-closes the task every step (DroneEnv.py:1741-1744); only the first time has any effect
    P.exit_area = live && hy <= 0 &&
                  !(V.k_status()[P.k] == 2 && (V.k_deadline()[P.k] < 0 || V.k_counted()[P.k] != 0));
    return P;
  }

  MUAV_HD void threat_commit(const ThreatPlan& P) {
    if (P.retarget) {
      V.h_target()[P.hid] = P.target;
      V.h_intercept()[P.hid] = P.intercept;
    }
    V.h_posx()[P.hid] = P.hx;
    V.h_posy()[P.hid] = P.hy;
    V.k_posx()[P.k] = P.hx;
    V.k_posy()[P.k] = P.hy;
  }
  // the part of an update_threats iteration that reaches beyond the threat (after its move has been committed)
  MUAV_HD void threat_effects(int hid, bool eng) {
    if (eng) engage(hid);
    if (V.h_posy()[hid] <= 0) {
      const int k = V.h_task()[hid] - 1;
      close_task(k);
      mark_outcome(k, false);
    }
  }

  MUAV_HD void update_threats_lanes(int lane, int nlanes, bool alive_env) {
#if defined(__CUDA_ARCH__)
    const int n = alive_env ? HIv(N_ACTIVE) : 0;   // warp-uniform: a warp steps one environment
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      int lo = base;
      const int hi = base + 32 < n ? base + 32 : n;
      while (lo < hi) {
        const bool mine = i >= lo && i < hi;
        const ThreatPlan P = threat_plan(mine ? i : 0, mine);
        unsigned em = __ballot_sync(0xffffffffu, mine && (P.engage || P.exit_area));
        unsigned gm = __ballot_sync(0xffffffffu, mine && P.engage);
        // without escorts every move of the round commits; with escorts only up to the first engagement (see above)
        const int fe = (MUAV_F_ESCORT(C().escort_enabled) && gm) ? base + (__ffs((int)gm) - 1) : hi - 1;
        if (mine && i <= fe && P.live) threat_commit(P);
        __syncwarp();
        em &= fe - base >= 31 ? 0xffffffffu : ((2u << (fe - base)) - 1u);
        if (em) {
          if (lane == 0) {
            while (em) {
              const int b = __ffs((int)em) - 1;
              em &= em - 1;
              threat_effects(V.h_order()[base + b], (gm >> b) & 1u);
            }
          }
          __syncwarp();
        }
        lo = fe + 1;
      }
    }
#else
    // the same rounds with the lanes emulated one after the other (all plans of a round before its commits)
    (void)lane; (void)nlanes;
    if (!alive_env) return;
    const int n = HIv(N_ACTIVE);
    int lo = 0;
    while (lo < n) {
      ThreatPlan P[MUAV_MAX_THREATS_HOST];
      for (int i = lo; i < n; ++i) P[i] = threat_plan(i);
      int fe = n - 1;
      if (MUAV_F_ESCORT(C().escort_enabled))
        for (int i = n - 1; i >= lo; --i)
          if (P[i].engage) fe = i;
      for (int i = lo; i <= fe; ++i)
        if (P[i].live) threat_commit(P[i]);
      for (int i = lo; i <= fe; ++i)
        if (P[i].engage || P[i].exit_area) threat_effects(P[i].hid, P[i].engage);
      lo = fe + 1;
    }
#endif
  }
  // one iteration of update_threats (i indexes env.threats, spawn order)
  MUAV_HD MUAV_NI_H void update_threat(int i) {
    {
      int hid = V.h_order()[i];
      if (V.h_status()[hid] == 2) return;
      double sp = C().speed[V.h_type()[hid]];
      double hx = V.h_posx()[hid], hy = V.h_posy()[hid];
      if (V.h_status()[hid] == 0 || V.h_target()[hid] < 0) {
        hx = hx + sp * 0.0;
        hy = hy + sp * -1.0;
        V.h_posx()[hid] = hx;
        V.h_posy()[hid] = hy;
      } else {
        if (MUAV_F_ESCORT(C().escort_enabled)) retarget_via_escort(hid);
        int tg = V.h_target()[hid];
        double dx = V.a_posx()[tg] - hx, dy = V.a_posy()[tg] - hy;
        double mag = norm2(dx, dy);
        double nx = 0.0, ny = 0.0;
        if (mag != 0) { nx = ddiv(dx, mag); ny = ddiv(dy, mag); }
        hx = hx + sp * nx;
        hy = hy + sp * ny;
        V.h_posx()[hid] = hx;
        V.h_posy()[hid] = hy;
        if (norm2(V.a_posx()[tg] - hx, V.a_posy()[tg] - hy) < C().engage[V.h_type()[hid]]) engage(hid);
      }
      int k = V.h_task()[hid] - 1;
      V.k_posx()[k] = V.h_posx()[hid];
      V.k_posy()[k] = V.h_posy()[hid];
      if (V.h_posy()[hid] <= 0) {
        close_task(k);
        mark_outcome(k, false);
      }
    }
  }

  // random_position (DroneEnv.py:1371-1410) for mission-area draws of the target stream
  MUAV_HD bool random_position_area(int stream, int area, double* ox, double* oy) {
    const double* m = &V.hf()[HF_M0_X + 4 * area];
    int nobs = MUAV_F_NOBS(V.lay().D.NOBS);
    for (int tries = 0; tries < 100; ++tries) {
      double x = rng_uniform(stream, m[0], m[0] + m[2]);
      double y = rng_uniform(stream, m[1], m[1] + m[3]);
      bool ok = true;
      for (int o = 0; o < nobs; ++o) {
        const double* ob = &V.obst()[3 * o];
        double d = norm2(x - ob[0], y - ob[1]) - 3;
        if (d < ob[2] + 20) { ok = false; break; }
      }
      if (ok) { *ox = x; *oy = y; return true; }
    }
    HIv(ERRFLAGS) |= ERR_NO_SPACE;
    return false;
  }

  // inject_dynamic_arrivals (DroneEnv.py:1646-1689); the rate draw precedes the capacity gate
  MUAV_HD MUAV_NI_H void inject_arrivals() {
    if (C().arrival_rate <= 0 || HIv(T) < 5) return;
    if (rng_random(1) >= C().arrival_rate) return;
    if (HIv(N_TASKS) >= C().max_tasks - 1) return;
    int ti = rng_below(1, 2) == 0 ? TT_ATT : TT_REC;
    int area = rng_below(2, 3);
    double x, y;
    if (C().dual_region_bursts) {
      double mid = C().area_w * 0.5;
      double wide = 40.0;
      if (rng_random(1) < 0.5) x = rng_uniform(1, wide, mid - wide);
      else x = rng_uniform(1, mid + wide, C().area_w - wide);
      y = rng_uniform(1, C().area_h * 0.2, C().area_h * 0.8);
    } else {
      if (!random_position_area(1, area, &x, &y)) return;
    }
    int tid = new_task(x, y, ti);
    if (tid == 0) return;
    int k = tid - 1;
    int TC = V.lay().D.TC;
    V.k_cur2(ti, k) = 1.0;
    sync_req(k);
    V.k_org_ti()[k] = 1.0;
    V.k_created()[k] = (int16_t)HIv(T);
    HIv(N_ARRIVALS) += 1;
    register_dynamic(tid);
    push_event(EV_THREAT, tid);
    push_event(EV_RESET, ti);
    HIv(PENDING_RESET) = 1;
  }

  // _sync_escorts (DroneEnv.py:1964-2000)
  MUAV_HD MUAV_NOINLINE void sync_escorts() {
    for (int a = 0; a < A(); ++a) {
      if (V.a_state()[a] == -1 || !is_recon(V.a_type()[a])) continue;
      if (qlen(a) == 0) continue;
      int cur = qat(a, 0);
      if (V.k_type()[cur - 1] == TT_REC && V.k_status()[cur - 1] != 2 && V.a_escort()[a] == 0) create_escort_for(a, cur);
    }
    // dict insertion order == ascending escort task id
    int prev = 0;
    for (;;) {
      int esc = 0, recon = -1;
      for (int a = 0; a < A(); ++a) {
        int e = V.a_escort()[a];
        if (e > prev && (esc == 0 || e < esc)) { esc = e; recon = a; }
      }
      if (esc == 0) break;
      prev = esc;
      int k = esc - 1;
      int rec_task = V.k_prot_task()[k];
      bool dead = V.a_state()[recon] == -1;
      int st = V.a_state()[recon];
      bool idle = qlen(recon) == 0 || st == 0 || st == 3;
      bool rec_done = rec_task != 0 && V.k_status()[rec_task - 1] == 2;
      bool wrong = qlen(recon) > 0 && (rec_task == 0 || qat(recon, 0) != rec_task);
      if (dead || idle || rec_done || wrong) {
        retire_escort(esc, dead);
        continue;
      }
      V.k_posx()[k] = V.a_posx()[recon];
      V.k_posy()[k] = V.a_posy()[recon];
      HIv(ESC_REQ_STEPS) += 1;
      if (fighters_near(recon, C().escort_radius, near_i(), near_d()) > 0) HIv(ESC_COV_STEPS) += 1;
    }
  }

  // _wps_update_sensing (DroneEnv.py:1506-1523): independent per (agent, task) -> spread over lanes
  MUAV_HD void update_sensing(int lane, int nlanes) {
    if (C().sense_radius <= 0) return;
    int n = HIv(N_TASKS);
    int Aa = A();
    // candidates = tasks that are open and dynamic, compacted first (WPS_escort creates ~250 ids of which ~25 are open)
    int16_t* cand = (int16_t*)(scratch + ((8 * 4 * Aa + 2 * Aa + 16 + 15) & ~15));
    int m = 0;
#if defined(__CUDA_ARCH__)
    for (int base = 0; base < n; base += 32) {
      const int k = base + lane;
      const bool c = k < n && V.k_status()[k] != 2 && !(V.k_created()[k] <= 0 && V.k_deadline()[k] < 0);
      const unsigned mk = __ballot_sync(0xffffffffu, c);
      if (c) cand[m + __popc(mk & ((1u << lane) - 1u))] = (int16_t)k;
      m += __popc(mk);
    }
    __syncwarp();
#else
    for (int k = 0; k < n; ++k)
      if (V.k_status()[k] != 2 && !(V.k_created()[k] <= 0 && V.k_deadline()[k] < 0)) cand[m++] = (int16_t)k;
#endif
    for (int idx = lane; idx < Aa * m; idx += nlanes) {
      int a = idx / m, k = cand[idx - a * m];
      if (V.a_state()[a] == -1) continue;
      if (known_bit(a, k)) continue;
      double d = norm2(V.a_posx()[a] - V.k_posx()[k], V.a_posy()[a] - V.k_posy()[k]);
      if (d <= C().sense_radius) {
#if defined(__CUDA_ARCH__)
        atomicOr(&V.known()[(k >> 5) * Aa + a], 1u << (k & 31));
#else
        set_known(a, k);
#endif
      }
    }
  }
  // _wps_process_reveals (DroneEnv.py:1525-1541)
  MUAV_HD void process_reveals() {
    int n = HIv(N_TASKS);
    int t = HIv(T);
    for (int k = 0; k < n; ++k) {
      int rt = V.k_reveal()[k];
      if (rt >= 0 && t >= rt) {
        V.k_reveal()[k] = -1;
        if (C().share_knowledge)
          for (int a = 0; a < A(); ++a) set_known(a, k);
      }
    }
  }
  // _wps_expire_windows (DroneEnv.py:1557-1573)
  MUAV_HD void expire_windows() {
    if (!C().hard_windows) return;
    int n = HIv(N_TASKS);
    int t = HIv(T);
    for (int k = 0; k < n; ++k) {
      if (V.k_status()[k] == 2) continue;
      int dl = V.k_deadline()[k];
      if (dl < 0) continue;
      if (t > dl) expire_one(k);
    }
  }

  // core_sim::SimCore::avoid_obstacles (core_sim/src/sim_core.rs:24-59)
  MUAV_HD MUAV_NOINLINE_LEAF static void avoid_obstacles(const double* obst, int nobs, double px, double py, double mx, double my,
                                                    double* ax, double* ay) {
    const double PI = 3.14159265358979323846;
    double sx = 0.0, sy = 0.0;
    for (int o = 0; o < nobs; ++o) {
      double dx = obst[3 * o] - px, dy = obst[3 * o + 1] - py;
      double d = sqrt(dx * dx + dy * dy);
      double dz = d - obst[3 * o + 2];
      if (dz < 40.0) {
        double nx = dx / dz, ny = dy / dz;
        double f = 0.5 / (1.0 - log(dmax(1.05, dz)));
        double ang = atan2(my, mx) - atan2(dy, dx);
        ang = fmod(ang + PI, 2.0 * PI) - PI;
        double rx, ry;
        if (ang > 0.0) { rx = ny; ry = -nx; }
        else { rx = -ny; ry = nx; }
        sx += rx * f;
        sy += ry * f;
      }
    }
    *ax = sx;
    *ay = sy;
  }

  // NumPy pairwise_sum for n <= 128 (np.sum(dists), DroneEnv.py:1138)
  MUAV_HD static double np_sum(const double* a, int n) {
    if (n < 8) {
      double res = 0.0;
      for (int i = 0; i < n; ++i) res += a[i];
      return res;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    int lim = n - (n % 8);
    for (; i < lim; i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }

  // task id of the idx-th open task of last_tasks_info (DroneEnv.py:492,827-830); 0 if out of range
  MUAV_HD int open_task_at(int idx) const {
    int n_open = V.hi()[HI_N_OPEN];
    if (idx < 0) idx += n_open;  // Python negative indexing
    if (idx < 0 || idx >= n_open) return 0;
    int KW = V.lay().D.KW;
    for (int w = 0; w < KW; ++w) {
      uint32_t m = V.open_mask()[w];
      int c = 0;
      uint32_t mm = m;
      while (mm) { mm &= mm - 1; ++c; }
      if (idx < c) {
        for (int b = 0; b < 32; ++b)
          if ((m >> b) & 1u) {
            if (idx == 0) return w * 32 + b + 1;
            --idx;
          }
      }
      idx -= c;
    }
    return 0;
  }
  MUAV_HD bool in_last_open(int tid) const { return (V.open_mask()[(tid - 1) >> 5] >> ((tid - 1) & 31)) & 1u; }

  // ------------------------------------------------------------------ step: part 1 (lane 0)
  // events drain, actions, kinematics FSM, distances, threats, arrivals, escorts.
  // act_agent/act_tid: ordered (agent, task id) pairs, tid == 0 encodes an out-of-range index.
  struct Acc {
    double action_reward, distance_reward, quality_reward, S_q, time_pen, alloc_reward;
  };

  Acc acc;  // reward accumulators of the current step (meaningful on lane 0)

  // step part 1a: events drain + ordered actions
  MUAV_HD void step_pre_a(const int16_t* act_agent, const int16_t* act_tid, int n_act) {
    acc.action_reward = 0.0;
    acc.distance_reward = 0.0;
    acc.quality_reward = 0.0;
    acc.S_q = 0.0;
    step_reward = 0.0;
    MUAV_TICK_START();
    int Aa = A();
    int TC = V.lay().D.TC;
    HIv(T) += 1;
    // drain events (DroneEnv.py:800-805)
    int nev = HIv(N_EVENTS);
    int tagmask = 0;
    n_out_events = nev;
    HIv(N_EVENTS) = 0;
    // events are processed from a private copy because release_all never appends (asserted by the oracle)
    for (int i = 0; i < nev; ++i) {
      int ev = V.events()[i];
      if (out_events) out_events[i] = ev;
      int tag = ev & 0xff;
      tagmask |= 1 << tag;
    }
    HIv(EV_TAGMASK) = tagmask;
_Pragma("unroll 1") for (int i = 0; i < nev; ++i) {
      int ev = V.events()[i];
      if ((ev & 0xff) == EV_RESET) release_all((ev >> 8) - 1);
    }

    MUAV_TICK(1);
    // ---- actions (DroneEnv.py:810-933)
    _Pragma("unroll 1") for (int i = 0; i < n_act; ++i) {
      int a = act_agent[i];
      if (a < 0 || a >= Aa) continue;
      if (V.a_state()[a] == -1) continue;
      int tid = act_tid[i];
      if (tid <= 0) {
        acc.action_reward += -1;
        continue;
      }
      int k = tid - 1;
      int head = qhead(a);
      if (head != tid) {
        if (head != 0) {
          acc.S_q -= 0.1;
          acc.S_q -= cap(a, V.k_type()[head - 1]);
          HIv(N_REALLOC) += 1;
          HIv(N_SWITCH) += 1;
          V.a_commit()[a] = 0;
          double ax = V.a_posx()[a], ay = V.a_posy()[a];
          double d_old = norm2(ax - V.k_posx()[head - 1], ay - V.k_posy()[head - 1]);
          double d_new = norm2(ax - V.k_posx()[k], ay - V.k_posy()[k]);
          acc.distance_reward += ddiv(d_old - d_new, C().max_coord);
        } else {
          acc.S_q += 0.05;
          if (HIv(PENDING_RESET) && C().dynamic_idle_penalty != 0.0) acc.S_q -= C().dynamic_idle_penalty;
        }
      } else {
        acc.S_q += 0.05;
        continue;
      }
      if (!C().multiple_tasks_per_agent) env_desallocate_all(a);
      if (!is_valid(a, tid)) {
        acc.action_reward += -1;
        continue;
      }
      if (allocate(a, tid)) {
        if (a < 32) V.k_tbl_lo()[k] |= 1u << a;
        else V.k_tbl_hi()[k] |= 1u << (a - 32);
        int ti = V.k_type()[k];
        double cp = cap(a, ti);
        double missing = V.k_cur_ti()[k] - (V.k_alloc_ti()[k] - cp);
        missing = missing > 0 ? missing : 0.0;
        double rest = missing - cp;
        double added = missing - (rest > 0 ? rest : 0.0);
        if (added <= 0) acc.S_q -= 1.5;
        acc.S_q += added;
        V.k_status()[k] = 1;
        // calculate_agent_expected_reward (DroneEnv.py:1216-1229)
        double rx, ry;
        int ql = qlen(a);
        if (ql >= 2) {
          int pk = qat(a, ql - 2) - 1;
          rx = V.k_posx()[pk];
          ry = V.k_posy()[pk];
        } else {
          rx = V.a_posx()[a];
          ry = V.a_posy()[a];
        }
        acc.distance_reward += ddiv(-1.0 * norm2(V.a_nfpx()[a] - rx, V.a_nfpy()[a] - ry), C().max_coord);
        if (V.a_state()[a] != 1 && V.a_state()[a] != -1) V.a_state()[a] = 1;
        if (MUAV_F_ESCORT(C().escort_enabled) && ti == TT_REC && is_recon(V.a_type()[a]) && V.a_escort()[a] == 0) create_escort_for(a, tid);
      }
    }

    MUAV_TICK(2);
  }

  // ------------------------------------------------------------------ agent-per-lane kinematics
  // The loop at DroneEnv.py:965-1129 visits the agents in index order, but almost every visit touches only the agent's
  // own state: fly towards the task / the base, arrive, wait out the task duration, drop a closed task.  fsm_plan(a)
  // evaluates one visit WITHOUT writing anything and classifies it:
  //   simple  -- every write goes to agent a's own fields (state, position, task_start, re_eval / last_task, own queue):
  //              fsm_commit(a, plan) applies it; simple visits of different agents commute, so they run one per lane;
  //   complex -- the visit has cross-agent side effects (agent failure -> events, an Int engagement writing the threat's
  //              target, a task completion -> task / counters / reward / escort retirement): it runs through the
  //              sequential fsm_agent(a), after every earlier agent has committed and before any later agent is planned.
  // The arithmetic of a simple visit is fsm_agent's, operation for operation (one norm2 of the measured vector, its two
  // divisions, the second normalisation of move + avoid, DroneEnv.py:1014-1048,1116-1127); only the control flow is
  // arranged so that all lanes call the out-of-line float64 helpers together.
  struct FsmPlan {
    double px, py;        // position after the visit
    int state;            // agent state after the visit
    int task_start;       // a_task_start after the visit
    bool alive, complex, drop_closed;
    int cur;
  };

  MUAV_HD FsmPlan fsm_plan(int a, bool lane_on = true) const {
    FsmPlan P;
    P.alive = P.complex = P.drop_closed = false;
    P.cur = 0;
    const int t = V.hi()[HI_T];
    const double bx = C().base_x, by = C().base_y;
    const int nobs = MUAV_F_NOBS(V.lay().D.NOBS);
    int st = lane_on ? (int)V.a_state()[a] : -1;
    const bool alive = st != -1;
    if (!alive) a = 0;  // idle lanes run the arithmetic below on agent 0's data and discard it: the calls stay converged
    P.alive = alive;
    P.state = st;
    P.task_start = V.a_task_start()[a];
    double px = V.a_posx()[a], py = V.a_posy()[a];
    P.px = px;
    P.py = py;
    const double speed = speed_of(a);
    const bool re_eval = V.a_re_eval()[a] != 0;
    const int ql = qlen(a);
    const int cur = re_eval ? (int)V.a_last_task()[a] : (ql > 0 ? (int)V.a_queue()[a] : 0);
    P.cur = cur;
    const bool failing = alive && V.a_fail_event()[a] == t;
    const bool closed = cur > 0 && V.k_status()[cur - 1] == 2;
    const bool on_task = alive && cur > 0 && !closed;
    const int k = on_task ? cur - 1 : -1;
    const int ti = on_task ? (int)V.k_type()[k] : -1;
    double tx = 0.0, ty = 0.0;
    if (on_task) { tx = V.k_posx()[k]; ty = V.k_posy()[k]; }
    // the one vector whose length this visit measures: towards the base (idle far-check, returning) or towards the task
    const bool base_vec = alive && ((st == 0 && !re_eval && ql == 0) || st == 3);
    const bool task_vec = on_task && (st == 1 || (st == 2 && ti == TT_INT));
    double vx = 0.0, vy = 0.0;
    if (base_vec) { vx = bx - px; vy = by - py; }      // norm2(P - B) == norm2(B - P) bit for bit (exact negation, squares)
    else if (task_vec) { vx = tx - px; vy = ty - py; }
    const double d = norm2(vx, vy);
    const bool div_ok = (task_vec && st == 1) ? !(fabs(d) < 1e-12) : (base_vec ? d != 0 : false);
    const double den = div_ok ? d : 1.0;
    const double qx = ddiv(vx, den), qy = ddiv(vy, den);
    const double nx = div_ok ? qx : 0.0, ny = div_ok ? qy : 0.0;

    double mvx = 0.0, mvy = 0.0;
    bool moving = false;
    if (failing) P.complex = true;
    if (alive && !failing) {
      if (st == 0 && !re_eval && ql == 0 && d > speed + 5) st = 3;
      if (closed) {
        P.drop_closed = true;
      } else if (on_task) {
        if (st == 1) {
          if (ti == TT_INT) {
            if (d < engage_of(a)) P.complex = true;          // engagement start writes the threat's target
            else { mvx = nx; mvy = ny; moving = true; }
          } else if (d < speed) {
            st = 2;
            P.task_start = t;
            px = tx;
            py = ty;
          } else { mvx = nx; mvy = ny; moving = true; }
        } else if (st == 2) {
          if (ti == TT_INT && d >= engage_of(a)) st = 1;
          if (P.task_start == -1) {
            P.task_start = t;
            px = tx;
            py = ty;
          } else if ((t - P.task_start) >= C().duration[ti] && (ti == TT_REC || ti == TT_ATT)) {
            P.complex = true;                                // task completion
          }
        }
      }
      if (st == 3) {
        if (d < speed + 5) st = 0;
        else { mvx = nx; mvy = ny; moving = true; }
      }
    }
    double avx = 0.0, avy = 0.0;
    if (nobs > 0 && moving && !P.complex) avoid_obstacles(V.obst(), nobs, px, py, mvx, mvy, &avx, &avy);
    const double sx = mvx + avx, sy = mvy + avy;
    const double mag = norm2(sx, sy);
    const bool mag_ok = mag != 0;
    const double mden = mag_ok ? mag : 1.0;
    const double wx = ddiv(sx, mden), wy = ddiv(sy, mden);
    const double ux = mag_ok ? wx : 0.0, uy = mag_ok ? wy : 0.0;
    px = px + ux * speed;
    py = py + uy * speed;
    px = dmin(dmax(px, 0.0), C().area_w);
    py = dmin(dmax(py, 0.0), C().area_h);
    P.px = px;
    P.py = py;
    P.state = st;
    return P;
  }

  MUAV_HD void fsm_commit(int a, const FsmPlan& P) {
    if (P.drop_closed) {
      des_allocate(a, P.cur);
      V.a_re_eval()[a] = 0;
      V.a_last_task()[a] = -1;
    }
    V.a_state()[a] = P.state;
    V.a_task_start()[a] = P.task_start;
    V.a_posx()[a] = P.px;
    V.a_posy()[a] = P.py;
  }

  // step part 1b: kinematics FSM (DroneEnv.py:965-1129).  Warp: rounds of {plan every remaining agent on its own lane,
  // commit the simple ones in front of the first complex visit, run that one sequentially}; a step without complex
  // visits (about five in six) is a single round.  Host build: the same plan / commit / fsm_agent calls in agent order,
  // which is what the golden replays of tests/test_hostcheck_golden.py exercise.
  MUAV_HD void step_pre_b(int lane, int nlanes, bool alive_env) {
    const int Aa = A();
#if defined(__CUDA_ARCH__)
    if (!alive_env) return;   // warp-uniform; a warp without an environment has no staged record to read
    for (int base = 0; base < Aa; base += 32) {
      const int a = base + lane;
      int lo = base;
      const int hi = base + 32 < Aa ? base + 32 : Aa;
      while (lo < hi) {   // warp-uniform
        const bool mine = alive_env && a >= lo && a < hi;
        const FsmPlan P = fsm_plan(mine ? a : 0, mine);
        const unsigned cm = __ballot_sync(0xffffffffu, mine && P.complex);
        const int fc = cm ? base + (__ffs((int)cm) - 1) : hi;
        if (mine && a < fc && P.alive) fsm_commit(a, P);
        __syncwarp();
        if (fc < hi) {
          if (lane == 0) fsm_agent(fc);
          __syncwarp();
        }
        lo = fc + 1;
      }
    }
#else
    // the same rounds with the lanes emulated one after the other: every plan of a round is made BEFORE any commit of
    // that round, exactly what the warp does
    (void)lane; (void)nlanes;
    if (!alive_env) return;
    int lo = 0;
    while (lo < Aa) {
      FsmPlan P[MUAV_MAX_AGENTS];
      int fc = Aa;
      for (int a = lo; a < Aa; ++a) P[a] = fsm_plan(a);
      for (int a = Aa - 1; a >= lo; --a)
        if (P[a].complex) fc = a;
      for (int a = lo; a < fc; ++a)
        if (P[a].alive) fsm_commit(a, P[a]);
      if (fc < Aa) fsm_agent(fc);
      lo = fc + 1;
    }
#endif
  }

  // kinematics FSM of one agent (one iteration of the loop at DroneEnv.py:965-1129)
  MUAV_HD void fsm_agent(int a) {
    const int TC = V.lay().D.TC;
    const int t = HIv(T);
    const double bx = C().base_x, by = C().base_y;
    const int nobs = MUAV_F_NOBS(V.lay().D.NOBS);
    {
      if (V.a_state()[a] == -1) return;
      if (V.a_fail_event()[a] == t) {
        V.a_state()[a] = -1;
        des_allocate_all(a);
        push_event(EV_RESET, -1);
        push_event(EV_FAIL, a);
        HIv(PENDING_RESET) = 1;
        return;
      }
      double mvx = 0.0, mvy = 0.0, avx = 0.0, avy = 0.0;
      double px = V.a_posx()[a], py = V.a_posy()[a];
      double speed = speed_of(a);
      if (V.a_state()[a] == 0 && !V.a_re_eval()[a]) {
        if (qlen(a) == 0 && norm2(px - bx, py - by) > speed + 5) V.a_state()[a] = 3;
      }
      int cur = V.a_re_eval()[a] ? V.a_last_task()[a] : qhead(a);
      if (cur > 0 && V.k_status()[cur - 1] == 2) {
        des_allocate(a, cur);
        V.a_re_eval()[a] = 0;
        V.a_last_task()[a] = -1;
      } else if (cur > 0) {
        int k = cur - 1;
        int ti = V.k_type()[k];
        double tx = V.k_posx()[k], ty = V.k_posy()[k];
        if (V.a_state()[a] == 1) {
          double dx = tx - px, dy = ty - py;
          double d = norm2(dx, dy);
          double nx = 0.0, ny = 0.0;
          if (!(fabs(d) < 1e-12)) { nx = ddiv(dx, d); ny = ddiv(dy, d); }
          if (ti == TT_INT) {
            if (d < engage_of(a)) {
              V.a_state()[a] = 2;
              V.h_target()[V.k_threat()[k]] = (int16_t)a;
              V.a_task_start()[a] = t;
            } else {
              mvx = nx; mvy = ny;
              if (nobs > 0) avoid_obstacles(V.obst(), nobs, px, py, mvx, mvy, &avx, &avy);
            }
          } else if (d < speed) {
            V.a_state()[a] = 2;
            V.a_task_start()[a] = t;
            V.a_posx()[a] = px = tx;
            V.a_posy()[a] = py = ty;
          } else {
            mvx = nx; mvy = ny;
            if (nobs > 0) avoid_obstacles(V.obst(), nobs, px, py, mvx, mvy, &avx, &avy);
          }
        } else if (V.a_state()[a] == 2) {
          if (ti == TT_INT) {
            double d = norm2(tx - px, ty - py);
            if (d >= engage_of(a)) V.a_state()[a] = 1;
          }
          if (V.a_task_start()[a] == -1) {
            V.a_task_start()[a] = t;
            V.a_posx()[a] = px = tx;
            V.a_posy()[a] = py = ty;
          } else if ((t - V.a_task_start()[a]) >= C().duration[ti] && (ti == TT_REC || ti == TT_ATT) && V.k_status()[k] != 2) {
            double t0 = 0.0;
            bool popped = task_done(a, cur, &t0);
            V.k_done_ti()[k] = V.k_done_ti()[k] + cap(a, ti);
            _Pragma("unroll 1") for (int c = 0; c < 6; ++c) V.k_cur2(c, k) = V.k_cur2(c, k) - cap(a, c);
            sync_req(k);
            if (popped) {
              remove_agent_cap(k, a, t0);
            } else if (qfind(a, cur) >= 0) {
              HIv(ERRFLAGS) |= 64;  // details/queue invariant broken (never observed)
            }
            if (V.k_done_ti()[k] >= V.k_org_ti()[k]) {
              if (V.k_kind()[k] != 1) mark_reached(k);
              if (V.k_status()[k] != 2) {
                acc.quality_reward += V.k_org_ti()[k] * 2;
                HFv(F_REWARD) += ddiv(V.k_org_ti()[k] * 1, HFv(NORM_FACTOR));
                if (V.k_kind()[k] != 1) mark_outcome(k, true);
                close_task(k);
                if (ti == TT_REC && is_recon(V.a_type()[a])) {
                  HIv(PROT_REC_DONE) += 1;
                  retire_escort(V.a_escort()[a], false);
                }
                if (all_done()) HIv(CONCLUSION) = t;
              }
            } else {
              acc.quality_reward += cap(a, ti);
            }
          }
        }
      }
      if (V.a_state()[a] == 3) {
        px = V.a_posx()[a];
        py = V.a_posy()[a];
        if (norm2(px - bx, py - by) < speed + 5) {
          V.a_state()[a] = 0;
        } else {
          double dx = bx - px, dy = by - py;
          double mag = norm2(dx, dy);
          if (mag == 0) { mvx = 0.0; mvy = 0.0; }
          else { mvx = ddiv(dx, mag); mvy = ddiv(dy, mag); }
          if (nobs > 0) avoid_obstacles(V.obst(), nobs, px, py, mvx, mvy, &avx, &avy);
        }
      }
      double sx = mvx + avx, sy = mvy + avy;
      double mag = norm2(sx, sy);
      double ux = 0.0, uy = 0.0;
      if (mag != 0) { ux = ddiv(sx, mag); uy = ddiv(sy, mag); }
      px = V.a_posx()[a] + ux * speed;
      py = V.a_posy()[a] + uy * speed;
      px = dmin(dmax(px, 0.0), C().area_w);
      py = dmin(dmax(py, 0.0), C().area_h);
      V.a_posx()[a] = px;
      V.a_posy()[a] = py;
    }

  }

  // positions before the agents move (DroneEnv.py:796-797) / per-agent travelled distance (DroneEnv.py:1131-1137):
  // independent per agent, one agent per lane
  MUAV_HD void snapshot_positions(int lane, int nlanes) {
    const int Aa = A();
    double* prev_x = (double*)scratch;
    double* prev_y = prev_x + Aa;
    for (int a = lane; a < Aa; a += nlanes) {
      prev_x[a] = V.a_posx()[a];
      prev_y[a] = V.a_posy()[a];
    }
  }
  MUAV_HD void travelled_distances(int lane, int nlanes) {
    const int Aa = A();
    double* prev_x = (double*)scratch;
    double* prev_y = prev_x + Aa;
    double* dists = prev_y + Aa;
    for (int a = lane; a < Aa; a += nlanes) {
      dists[a] = norm2_rows(V.a_posx()[a] - prev_x[a], V.a_posy()[a] - prev_y[a]);
      V.a_dist()[a] += dists[a];
    }
  }

  // step part 1c: total distance, time penalty, threat generation
  MUAV_HD void step_pre_c() {
    MUAV_TICK_START();
    const int Aa = A();
    const int t = HIv(T);
    double* dists = (double*)scratch + 2 * Aa;
    HFv(TOTAL_DIST) += np_sum(dists, Aa);   // np.sum's pairwise order (DroneEnv.py:1138)

    // time_penaulty / alloc_reward are evaluated here in the reference (DroneEnv.py:1140-1145)
    {
      double nt = (double)C().n_tasks_cfg;
      acc.time_pen = ddiv(-(double)(C().n_tasks_cfg - HIv(N_REACHED)), nt) * ddiv((double)t, (double)C().max_time_steps);
      acc.alloc_reward = 0.0;
      if (t > C().n_tasks_cfg + 1 && C().rw[5] != 0.0) {  // weight 0 (WPS flags): the count cannot reach the reward
        int unalloc = 1 + HIv(N_FREED_EMPTY_TBL);  // bucket 0 (idle) is always empty; recycled tasks are counted
        int n = HIv(N_TASKS);
        for (int k = 0; k < n; ++k)
          if (V.k_slot()[k] >= 0 && V.k_tbl_lo()[k] == 0 && V.k_tbl_hi()[k] == 0) ++unalloc;
        acc.alloc_reward = -(double)unalloc;
      }
    }

    MUAV_TICK(4);
    generate_threat();
  }
  // step part 1d: arrivals, escorts (after update_threats)
  MUAV_HD void step_pre_d() {
    MUAV_TICK_START();
    inject_arrivals();
    if (MUAV_F_ESCORT(C().escort_enabled)) sync_escorts();
    MUAV_TICK(6);
  }

  // ------------------------------------------------------------------ warp-parallel task scans
#if defined(__CUDA_ARCH__)
  // _wps_process_reveals over 32 tasks per iteration (lane <-> task, one known-word per iteration)
  __device__ void process_reveals_warp(int lane) {
    const int n = HIv(N_TASKS), t = HIv(T), Aa = A();
    const int words = (n + 31) >> 5;
    for (int w = 0; w < words; ++w) {
      const int k = (w << 5) + lane;
      const bool rev = k < n && V.k_reveal()[k] >= 0 && t >= V.k_reveal()[k];
      const unsigned m = __ballot_sync(0xffffffffu, rev);
      if (rev) V.k_reveal()[k] = -1;
      if (m != 0u && C().share_knowledge)
        for (int a = lane; a < Aa; a += 32) V.known()[w * Aa + a] |= m;
    }
    __syncwarp();
  }
  // _wps_expire_windows: detection in parallel, the (rare) expiries are applied by lane 0 in task order
  __device__ void expire_windows_warp(int lane) {
    if (!C().hard_windows) return;
    const int n = HIv(N_TASKS), t = HIv(T);
    const int words = (n + 31) >> 5;
    for (int w = 0; w < words; ++w) {
      const int k = (w << 5) + lane;
      const bool ex = k < n && V.k_status()[k] != 2 && V.k_deadline()[k] >= 0 && t > V.k_deadline()[k];
      unsigned m = __ballot_sync(0xffffffffu, ex);
      if (m != 0u) {
        if (lane == 0) {
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            expire_one((w << 5) + b);
          }
        }
        __syncwarp();
      }
    }
  }
  // last_tasks_info mask + _counts_for_mission_done over all tasks; results are warp-uniform
  __device__ void scan_open_warp(int lane, int* n_open_out, bool* all_done_out) {
    const int n = HIv(N_TASKS);
    const int KW = V.lay().D.KW;
    int n_open = 0;
    bool blocking = false;
    for (int w = 0; w < KW; ++w) {
      const int k = (w << 5) + lane;
      const bool open = k < n && V.k_status()[k] != 2;
      const int ti = k < n ? V.k_type()[k] : 0;
      const bool blk = open && !(V.k_kind()[k] == 1 || ti == TT_DET || ti == TT_HOLD);
      const unsigned mo = __ballot_sync(0xffffffffu, open);
      const unsigned mb = __ballot_sync(0xffffffffu, blk);
      if (lane == 0) V.open_mask()[w] = mo;
      n_open += __popc(mo);
      blocking = blocking || mb != 0u;
    }
    *n_open_out = n_open;
    *all_done_out = !blocking;
  }
#endif

  MUAV_HD MUAV_NI_H void expire_one(int k) {
    close_task(k);
    V.k_fq()[k] = 0;
    mark_outcome(k, false);
    mark_reached(k);
    _Pragma("unroll 1") for (int a = 0; a < A(); ++a)
      if (qlen(a) > 0 && qat(a, 0) == k + 1) des_allocate_all(a);
  }

  // ------------------------------------------------------------------ step: part 3 (lane 0)
  // reserve tracking, reward, termination; `alld` / `n_open` come from the task scan
  MUAV_HD StepResult step_post(bool alld_scan, int n_open) {
    int Aa = A();
    // _wps_track_reserve (DroneEnv.py:1575-1580) and the _pending_reset clear (:1156-1160)
    int idle = 0;
    bool any_busy = false;
    for (int a = 0; a < Aa; ++a) {
      if (V.a_state()[a] == -1) continue;
      if (qlen(a) == 0) ++idle;
      else any_busy = true;
    }
    HIv(IDLE_RESERVE) += idle;
    if (HIv(PENDING_RESET) && any_busy) HIv(PENDING_RESET) = 0;

    const double* rw = C().rw;
    double time_reward = 0.0;
    double reward = (rw[0] * acc.action_reward + rw[1] * acc.distance_reward + rw[2] * acc.quality_reward +
                     rw[3] * acc.S_q + rw[4] * (double)C().n_tasks_cfg * time_reward + rw[5] * acc.alloc_reward +
                     rw[6] * acc.time_pen + rw[7] * step_reward);
    reward = ddiv(ddiv(reward, HFv(NORM_FACTOR)), (double)C().max_time_steps);
    int t = HIv(T);
    int n = HIv(N_TASKS);
    bool alld = n > 0 && alld_scan;
    bool timed_out = (t >= C().max_time_steps) && (C().max_time_steps > 0);
    bool done = timed_out || (C().early_terminate && alld);
    if (alld && HIv(CONCLUSION) > C().max_time_steps) HIv(CONCLUSION) = t;
    StepResult r;
    r.terminated = (C().early_terminate && alld && !timed_out) ? 1 : 0;
    r.truncated = timed_out ? 1 : 0;
    HIv(N_OPEN) = n_open;  // last_tasks_info (DroneEnv.py:492); the mask was written by the scan
    if (done) {
      reward = HFv(F_REWARD);
      HIv(DONE) = 1;
    }
    HFv(LAST_REWARD) = reward;
    r.reward = reward;
    return r;
  }

  // whole step.  The ordered phases (events, actions, threat generation, arrivals, escorts, the rare cross-agent visits of
  // the FSM / threat loops) run on lane 0; the per-agent, per-threat and per-task phases are spread over the warp.
  // `alive` = this warp has an environment to step; `sync_mask` = align the CTA's warps between phases.
  MUAV_HD StepResult step(const int16_t* act_agent, const int16_t* act_tid, int n_act, int lane, int nlanes,
                          bool alive = true, int sync_mask = 0) {
    if (alive) snapshot_positions(lane, nlanes);
    if (alive && lane == 0) step_pre_a(act_agent, act_tid, n_act);
    MUAV_WARP_SYNC();
    MUAV_CTA_SYNC(sync_mask & 2);
    {
      MUAV_TICK_START();
      step_pre_b(lane, nlanes, alive);
      MUAV_WARP_SYNC();
      if (alive) travelled_distances(lane, nlanes);
      MUAV_WARP_SYNC();
      MUAV_TICK(3);
      MUAV_CTA_SYNC(sync_mask & 4);
    }
    if (alive && lane == 0) step_pre_c();
    MUAV_WARP_SYNC();
    {
      MUAV_TICK_START();
      update_threats_lanes(lane, nlanes, alive);
      MUAV_WARP_SYNC();
      MUAV_TICK(5);
    }
    if (alive && lane == 0) step_pre_d();
    MUAV_WARP_SYNC();
    MUAV_CTA_SYNC(sync_mask & 8);
    StepResult r;
    r.reward = 0.0;
    r.terminated = r.truncated = 0;
    if (alive) {
      MUAV_TICK_START();
      update_sensing(lane, nlanes);
      MUAV_WARP_SYNC();
      MUAV_TICK(7);
      int n_open = 0;
      bool alld = true;
#if defined(__CUDA_ARCH__)
      process_reveals_warp(lane);
      expire_windows_warp(lane);
      __syncwarp();
      scan_open_warp(lane, &n_open, &alld);
      MUAV_TICK(8);
#else
      process_reveals();
      expire_windows();
      {
        const int n = HIv(N_TASKS);
        const int KW = V.lay().D.KW;
        for (int w = 0; w < KW; ++w) V.open_mask()[w] = 0;
        for (int k = 0; k < n; ++k)
          if (V.k_status()[k] != 2) {
            V.open_mask()[k >> 5] |= 1u << (k & 31);
            ++n_open;
          }
        alld = all_done();
      }
#endif
      if (lane == 0) r = step_post(alld, n_open);
      MUAV_WARP_SYNC();
      // recycle slots lazily: only when the next step could run out of them (new tasks per step <= threats in a
      // burst + one arrival + one escort per agent).  Dead tasks are dead data whether recycled or not.
      if (HIv(N_SLOTS_USED) + A() + V.lay().D.HC + 2 > V.lay().D.TC) {
        free_dead_tasks(lane, nlanes);
        MUAV_WARP_SYNC();
      }
      MUAV_TICK(9);
    }
    MUAV_CTA_SYNC(sync_mask & 16);
    return r;
  }
};

}  // namespace muav
