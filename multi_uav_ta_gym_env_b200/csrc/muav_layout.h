// Record layout of one environment (struct-of-arrays inside the record; the record is one
// contiguous, 16-byte aligned block in HBM so that a warp can stage it into shared memory
// with a single bulk async copy).  The field list is an X-macro so that the device view,
// the offset table exported through muav_field_info() and the Python packer agree by
// construction.
#pragma once
#include <stdint.h>
#include "../../include/muav.h"

#if defined(__CUDACC__)
#define MUAV_HD __host__ __device__
// MUAV_NOINLINE_LEAF: pure math helpers kept out of line (code size).  MUAV_NOINLINE: rarely executed member functions;
// inlined on the device by default because an out-of-line member takes `this`, which forces the whole Sim object
// (and with it the record base pointer) into local memory and turns every record access into a generic load.
#if defined(MUAV_ALL_INLINE)
#define MUAV_NOINLINE_LEAF
#else
#define MUAV_NOINLINE_LEAF __noinline__
#endif
#if defined(MUAV_MEMBERS_OUT_OF_LINE)
#define MUAV_NOINLINE __noinline__
#else
#define MUAV_NOINLINE
#endif
#else
#define MUAV_HD
#define MUAV_NOINLINE __attribute__((noinline))
#define MUAV_NOINLINE_LEAF __attribute__((noinline))
#endif
// tuning switches (see tools/kbench.py): out-of-line copies of the frequently used helpers / the allocator
#if defined(MUAV_NI_HOT)
#define MUAV_NI_H MUAV_NOINLINE
#else
#define MUAV_NI_H
#endif
#if defined(MUAV_NI_ALLOC)
#define MUAV_NI_A MUAV_NOINLINE
#else
#define MUAV_NI_A
#endif

namespace muav {

// ---- header (int32) indices -------------------------------------------------
#define MUAV_HI_LIST(X)                                                                           \
  X(T) X(N_TASKS) X(N_ACTIVE) X(N_EVENTS) X(N_REALLOC) X(N_SWITCH) X(N_ARRIVALS) X(PENDING_RESET) \
  X(N_MISSED) X(N_ON_TIME) X(N_WINDOWED) X(IDLE_RESERVE) X(BURST_TOGGLE) X(ESC_REQUESTS)          \
  X(ESC_COMPLETED) X(ESC_FAILED) X(ESC_REQ_STEPS) X(ESC_COV_STEPS) X(BREACHES) X(INTERCEPTED)     \
  X(RECON_LOSSES) X(ESCORT_LOSSES) X(MUTUAL) X(PROT_REC_DONE) X(N_REACHED) X(CONCLUSION)          \
  X(CUR_AGENT) X(CUR_TGT) X(CUR_MISSION) X(ERRFLAGS) X(DONE) X(LAST_PLAN_STEP) X(N_REPLANS)       \
  X(N_CALLS) X(EV_TAGMASK) X(N_OPEN) X(GROUP_NEXT0) X(GROUP_NEXT1) X(GROUP_NEXT2) X(GROUP_NEXT3)  \
  X(GROUP_NEXT4) X(GROUP_NEXT5) X(GROUP_NEXT6) X(GROUP_NEXT7) X(N_LSAP) X(N_FREED_EMPTY_TBL) X(N_SLOTS_USED) X(PAD2)

enum HdrI {
#define X(n) HI_##n,
  MUAV_HI_LIST(X)
#undef X
      HI_COUNT
};

// ---- header (double) indices ------------------------------------------------
#define MUAV_HF_LIST(X) \
  X(F_REWARD) X(TOTAL_DIST) X(NORM_FACTOR) X(LAST_REWARD) X(M0_X) X(M0_Y) X(M0_W) X(M0_H) X(M1_X) X(M1_Y) X(M1_W) \
  X(M1_H) X(M2_X) X(M2_Y) X(M2_W) X(M2_H)

enum HdrF {
#define X(n) HF_##n,
  MUAV_HF_LIST(X)
#undef X
      HF_COUNT
};

// error flag bits
enum ErrBits {
  ERR_QUEUE_OVERFLOW = 1,
  ERR_TASK_OVERFLOW = 2,
  ERR_EVENT_OVERFLOW = 4,
  ERR_TAPE_OVERFLOW = 8,
  ERR_NO_SPACE = 16,
  ERR_LSAP_INFEASIBLE = 32,
};

// TC = task SLOT capacity (tasks that are open or still referenced), IC = task ID capacity (tasks ever created)
struct Dims {
  int A, TC, IC, HC, QC, EVC, NOBS, KW;
};

MUAV_HD inline Dims dims_of(const muav_config& c) {
  Dims d{};
  d.A = c.n_agents;
  d.TC = c.task_cap;
  d.IC = c.id_cap > c.task_cap ? c.id_cap : c.task_cap;
  d.HC = c.n_threats;
  d.QC = c.queue_cap;
  d.EVC = c.event_cap;
  d.NOBS = c.n_obstacles;
  d.KW = (d.IC + 31) / 32;
  return d;
}

// name, C type, element count.
// X  = plain array; XS = per-task array indexed by SLOT and accessed by task index through k_slot
// (closed tasks that nothing references any more give their slot back, see Sim::free_dead_tasks).
// XC / XCS = the same, COLD: fields the common step never touches (full requirement vectors, allocation times,
// assignment table, pending events).  The record is [hot part | cold part]; the step kernel stages only the hot part
// into shared memory and reaches the cold part in global memory (L2) on the rare events that need it, which is what lets
// twice as many environments stay resident per SM.  Each part is ordered by decreasing alignment (8, 4, 2 bytes).
// k_cur_ti / k_alloc_ti are hot copies of component [task type] of the cold requirement vectors (the only component
// the allocator, the token builders and the validity checks read); Sim::sync_req keeps them equal after every update.
#define MUAV_FIELDS(X, XS, XC, XCS)  \
  X(hf, double, HF_COUNT)           \
  X(a_posx, double, D.A)            \
  X(a_posy, double, D.A)            \
  X(a_nfpx, double, D.A)            \
  X(a_nfpy, double, D.A)            \
  X(a_nft, double, D.A)             \
  X(a_dist, double, D.A)            \
  X(a_caps, double, 6 * D.A)        \
  XS(k_posx, double, D.TC)          \
  XS(k_posy, double, D.TC)          \
  XS(k_cur_ti, double, D.TC)        \
  XS(k_alloc_ti, double, D.TC)      \
  XS(k_done_ti, double, D.TC)       \
  XS(k_org_ti, double, D.TC)        \
  X(h_posx, double, D.HC)           \
  X(h_posy, double, D.HC)           \
  X(obst, double, 3 * D.NOBS)       \
  X(hi, int32_t, HI_COUNT)          \
  X(a_state, int32_t, D.A)          \
  X(a_task_start, int32_t, D.A)     \
  X(a_fail_event, int32_t, D.A)     \
  X(a_ammo, int32_t, D.A)           \
  X(a_last_task, int32_t, D.A)      \
  X(a_commit, int32_t, D.A)         \
  X(a_escort, int32_t, D.A)         \
  X(known, uint32_t, D.KW * D.A)    \
  X(open_mask, uint32_t, D.KW)      \
  X(a_queue, int16_t, D.QC * D.A)   \
  X(a_type, int16_t, D.A)           \
  X(a_re_eval, int16_t, D.A)        \
  X(a_qlen, int16_t, D.A)           \
  X(a_name_rank, int16_t, D.A)      \
  X(k_slot, int16_t, D.IC)          \
  X(k_status, int16_t, D.IC)        \
  X(k_type, int16_t, D.IC)          \
  X(k_reveal, int16_t, D.IC)        \
  X(s_used, int16_t, D.TC)          \
  XS(k_deadline, int16_t, D.TC)     \
  XS(k_created, int16_t, D.TC)      \
  XS(k_prot_task, int16_t, D.TC)    \
  XS(k_kind, int16_t, D.TC)         \
  XS(k_req_agents, int16_t, D.TC)   \
  XS(k_elig, int16_t, D.TC)         \
  XS(k_counted, int16_t, D.TC)      \
  XS(k_fq, int16_t, D.TC)           \
  XS(k_reached, int16_t, D.TC)      \
  XS(k_det_frozen, int16_t, D.TC)   \
  XS(k_threat, int16_t, D.TC)       \
  XS(k_prot_agent, int16_t, D.TC)   \
  X(h_task, int16_t, D.HC)          \
  X(h_det_task, int16_t, D.HC)      \
  X(h_status, int16_t, D.HC)        \
  X(h_type, int16_t, D.HC)          \
  X(h_group, int16_t, D.HC)         \
  X(h_ammo, int16_t, D.HC)          \
  X(h_target, int16_t, D.HC)        \
  X(h_mission, int16_t, D.HC)       \
  X(h_intercept, int16_t, D.HC)     \
  X(h_spawned, int16_t, D.HC)       \
  X(h_order, int16_t, D.HC)         \
  XC(k_cur, double, 6 * D.TC)       \
  XC(k_alloc, double, 6 * D.TC)     \
  XC(a_qtime, double, D.QC * D.A)   \
  XCS(k_init, double, D.TC)         \
  XCS(k_dtime, double, D.TC)        \
  XCS(k_tbl_lo, uint32_t, D.TC)     \
  XCS(k_tbl_hi, uint32_t, D.TC)     \
  XC(events, int32_t, D.EVC)

struct Layout {
#define X(name, type, count) int32_t o_##name;
  MUAV_FIELDS(X, X, X, X)
#undef X
  int32_t record_bytes;
  int32_t hot_bytes;           // [0, hot_bytes) is staged into shared memory by the step kernel; the rest stays in HBM
  int32_t scratch_bytes;
  int32_t step_scratch_bytes;  // scratch of a launch that does not run the allocator
  int32_t plain_scratch_bytes; // scratch of a launch with the plain Hungarian / PI allocator (no planner front end)
  int32_t act_bytes;           // ordered action list of the step (agent ids + task ids), kept behind the scratch
  Dims D;
};

MUAV_HD constexpr inline int32_t align_up(int32_t x, int32_t a) { return (x + a - 1) / a * a; }

// bytes of allocator scratch: cost[A*TC], u/v/spc[M], resid[TC] doubles + 6 M-sized and 2A+3TC int16 arrays
MUAV_HD constexpr inline int32_t alloc_scratch_bytes(int A, int TC, bool planner_block = true) {
  int M = A > TC ? A : TC;
  int b = 8 * (A * TC + 3 * M + TC + 4);
  b += 2 * (6 * M + 3 * A + 3 * TC);
  b = (b + 7) / 8 * 8;
  if (planner_block) b += 8 * TC + 8 * A + ((A + 7) / 8) * 8;  // planner priorities, lock scores, reserved mask
  return (b + 15) / 16 * 16;
}

// layout of a record with the given dimensions (a constant expression when the dimensions are: fixed-shape
// instantiations of the step kernel, MUAV_FIXED_SHAPE)
MUAV_HD constexpr inline Layout make_layout_dims(const Dims D) {
  Layout L{};
  L.D = D;
  int32_t off = 0;
  bool cold = false;
#define X(name, type, count)                 \
  off = align_up(off, (int32_t)sizeof(type)); \
  L.o_##name = off;                          \
  off += (int32_t)sizeof(type) * (int32_t)(count);
#define XCF(name, type, count)                                    \
  if (!cold) { cold = true; off = align_up(off, 16); L.hot_bytes = off; } \
  X(name, type, count)
  MUAV_FIELDS(X, X, XCF, XCF)
#undef X
#undef XCF
  L.record_bytes = align_up(off, 16);
  // per-warp scratch: allocator work arrays (muav_alloc.cuh carve_scratch) or the step's temporaries
  int32_t s = alloc_scratch_bytes(D.A, D.TC) + 8 * (D.IC - D.TC);  // planner priorities are indexed by task id
  int32_t s2 = 8 * 4 * D.A + 2 * D.A + 16;
  if (s2 > s) s = s2;
  L.scratch_bytes = align_up(s, 16);
  // launches that do not run the allocator only need the step's temporaries (+ the token builder's column list)
  // (the fused token emission keeps its column list here: room for max(TC, 64) columns, so that the standalone and
  // the fused builders accept the same max_tasks)
  int32_t s3 = s2 + 2 * ((D.TC > 64 ? D.TC : 64) + 4);
  L.step_scratch_bytes = align_up(s3, 16);
  if (L.step_scratch_bytes > L.scratch_bytes) L.step_scratch_bytes = L.scratch_bytes;
  int32_t s4 = alloc_scratch_bytes(D.A, D.TC, false);
  if (s2 > s4) s4 = s2;
  L.plain_scratch_bytes = align_up(s4, 16);
  if (L.plain_scratch_bytes < L.step_scratch_bytes) L.plain_scratch_bytes = L.step_scratch_bytes;
  L.act_bytes = align_up(4 * D.A, 16);
  return L;
}

MUAV_HD inline Layout make_layout(const muav_config& c) { return make_layout_dims(dims_of(c)); }

#if defined(MUAV_FIXED_SHAPE)
// MUAV_FIXED_SHAPE = A, TC, IC, HC, QC, EVC, NOBS: a step-kernel instantiation for ONE record shape.  Every field offset and
// loop bound is then a constant expression (a third of the general kernel's dynamic instructions are offset arithmetic on
// the run-time layout, profiles/r02_step_kernel.md); the launcher uses it only for configurations of exactly this shape.
MUAV_HD constexpr inline Dims fixed_dims() {
  constexpr int v[7] = {MUAV_FIXED_SHAPE};
  Dims d{};
  d.A = v[0]; d.TC = v[1]; d.IC = v[2] > v[1] ? v[2] : v[1]; d.HC = v[3]; d.QC = v[4]; d.EVC = v[5]; d.NOBS = v[6];
  d.KW = (d.IC + 31) / 32;
  return d;
}
#endif

// array of a slot-indexed task field, addressed by task index (id - 1)
template <class T>
struct SlotRef {
  T* p;
  const int16_t* slot;
  MUAV_HD inline T& operator[](int k) const { return p[slot[k]]; }
};

struct View {
  char* base;    // hot part of the record (shared memory inside the step kernel)
  char* cbase;   // record start as far as the COLD fields are concerned: the record in HBM inside the step kernel, == base
                 // wherever the whole record is addressed in place (standalone kernels, CPU build)
  MUAV_HD inline void at(char* rec) { base = rec; cbase = rec; }
#if defined(MUAV_FIXED_SHAPE)
  MUAV_HD inline void set_layout(const Layout*) {}
  MUAV_HD inline Layout lay() const {
    constexpr Layout k = make_layout_dims(fixed_dims());
    return k;
  }
#else
  const Layout* L;
  MUAV_HD inline void set_layout(const Layout* l) { L = l; }
  MUAV_HD inline const Layout& lay() const { return *L; }
#endif
#define X(name, type, count) \
  MUAV_HD inline type* name() const { return (type*)(base + lay().o_##name); }
#define XS(name, type, count)                                                                       \
  MUAV_HD inline SlotRef<type> name() const {                                                       \
    return SlotRef<type>{(type*)(base + lay().o_##name), (const int16_t*)(base + lay().o_k_slot)};  \
  }                                                                                                 \
  MUAV_HD inline type* name##_raw() const { return (type*)(base + lay().o_##name); }
#define XC(name, type, count) \
  MUAV_HD inline type* name() const { return (type*)(cbase + lay().o_##name); }
#define XCS(name, type, count)                                                                       \
  MUAV_HD inline SlotRef<type> name() const {                                                        \
    return SlotRef<type>{(type*)(cbase + lay().o_##name), (const int16_t*)(base + lay().o_k_slot)};  \
  }                                                                                                  \
  MUAV_HD inline type* name##_raw() const { return (type*)(cbase + lay().o_##name); }
  MUAV_FIELDS(X, XS, XC, XCS)
#undef X
#undef XS
#undef XC
#undef XCS
  // requirement vectors (cold): component c of task index k
  MUAV_HD inline double& k_cur2(int c, int k) const { return k_cur()[c * lay().D.TC + k_slot()[k]]; }
  MUAV_HD inline double& k_alloc2(int c, int k) const { return k_alloc()[c * lay().D.TC + k_slot()[k]]; }
};

}  // namespace muav
