// libmuav_b200.so -- CUDA kernels (sm_100a) and the C ABI declared in include/muav.h.
//
// Execution model: one warp per environment.  The environment record (struct-of-arrays, one
// contiguous 16-byte aligned block, ~10 KB for WPS_hard) is staged HBM -> shared memory with a
// single bulk async copy (cp.async.bulk + mbarrier, the 1-D TMA path; SASS: UBLKCP), the step
// (and optionally the Hungarian allocator in front of it) runs entirely out of shared memory,
// and the record goes back with one bulk store.  Per-step outputs are a few scalars per env.
// Compile with -fmad=false: every float64 operation must round exactly like the reference's.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

#include "muav_alloc.cuh"
#include "muav_views.cuh"

namespace muav {

// ------------------------------------------------------------------ bulk async copy helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(phase)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// wait until the bulk stores have READ their shared-memory source (the CTA may then exit and free it); the global writes
// themselves complete before the grid does
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

struct StepParams {
  muav_config cfg;
  Layout L;
  muav_alloc_opts opts;
  muav_step_out out;
  muav_token_out tok;
  char* records;
  const uint32_t* tapes;
  const int32_t* actions;
  int32_t* actions_out;  // allocate-only mode: ordered (agent, index) list per env
  long long* phase_out;  // MUAV_PHASE_TIMING builds: 16 cycle counters summed over warps
  int n_envs, n_steps, tape_stride, use_bulk, alloc_only, cta_warps, sync_mask;
  int scratch_launch;   // per-environment scratch bytes of this launch (allocator work arrays or step temporaries)
  int stage_bytes;      // bytes of each record staged into shared memory: hot_bytes, or record_bytes (cold part too)
  int actions_are_ids;  // actions hold (agent, task id) instead of (agent, index into last_tasks_info)
};

#define MUAV_MAX_CTA_WARPS 16
#define MUAV_MAX_DEVICES 64

// One warp per environment; a CTA holds `cta_warps` environments whose warps are phase-aligned with
// CTA barriers (no data is shared between them).
// Register budget: the general kernel may be launched with up to 16 warps per CTA (128 registers per thread, at most 16
// resident warps per SM).  A fixed-shape instantiation knows its shared-memory slot and states its own bound
// (MUAV_LB_THREADS threads per CTA at most, MUAV_LB_BLOCKS CTAs per SM): more resident environments per SM.
#if !defined(MUAV_LB_THREADS)
#define MUAV_LB_THREADS (32 * MUAV_MAX_CTA_WARPS)
#define MUAV_LB_BLOCKS 1
#endif
__global__ void __launch_bounds__(MUAV_LB_THREADS, MUAV_LB_BLOCKS) muav_step_kernel(const __grid_constant__ StepParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[MUAV_MAX_CTA_WARPS];
  __shared__ int n_act_s[MUAV_MAX_CTA_WARPS];
  const int W = P.cta_warps;
  const int w = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int slot = blockIdx.x * W + w;
  bool has_env = slot < P.n_envs;
  int e = slot;
  if (has_env && P.out.d_env_order) {
    e = P.out.d_env_order[slot];
    if (e < 0 || e >= P.n_envs) { has_env = false; e = 0; }
  }
  if (P.out.d_env_order && blockIdx.x == 0 && threadIdx.x == 0) {
    // counters of the buffer that the next launch will fill
    ((int32_t*)P.out.d_env_order)[P.n_envs] = 0;
    ((int32_t*)P.out.d_env_order)[P.n_envs + 1] = 0;
  }
  const int sync_mask = W > 1 ? P.sync_mask : 0;
#if defined(MUAV_FIXED_SHAPE)
  constexpr Layout L = make_layout_dims(fixed_dims());   // launch_step checked that the configuration has this shape
#else
  const Layout& L = P.L;
#endif
  // shared-memory slot of this warp's environment: [hot record | scratch | ordered action list]
  // (a launch may stage the cold part as well, stage_bytes == record_bytes: small records, many allocator updates.  The
  // specialised instantiations fix the choice at compile time -- MUAV_STAGE_COLD_FIXED -- so that the compiler knows
  // whether a cold access is a shared-memory or a global one and that it cannot alias the hot part)
#if defined(MUAV_STAGE_COLD_FIXED)
  const int stage_bytes = MUAV_STAGE_COLD_FIXED ? L.record_bytes : L.hot_bytes;
#else
  const int stage_bytes = P.stage_bytes;
#endif
  const int slot_bytes = stage_bytes + P.scratch_launch + L.act_bytes;
  char* rec = (char*)smem + (size_t)w * slot_bytes;
  char* scratch = rec + stage_bytes;
  char* grec = P.records + (size_t)(has_env ? e : 0) * L.record_bytes;
  int16_t* act_agent = (int16_t*)(scratch + P.scratch_launch);
  int16_t* act_tid = act_agent + L.D.A;

  // ---- allocator-only launch: environments whose replan rule does not fire leave after a look at their header
#if defined(MUAV_LEAN) && !defined(MUAV_LEAN_PLANNER)
  if (false) {   // the plain lean kernels are never launched allocator-only (launch_step): no second copy of the allocator
#else
  if (P.alloc_only && has_env) {
#endif
    int32_t* ghi = (int32_t*)(grec + L.o_hi);
    int go = 0;
    if (lane == 0) {
      const muav_alloc_opts& O = P.opts;
      if (!ghi[HI_DONE]) {
        const int t = ghi[HI_T];
        const int iv = O.replan_interval > 0 ? O.replan_interval : 1;
        const bool ev_hit = (ghi[HI_EV_TAGMASK] & O.event_mask) != 0;
        if (O.mode == 1 && (O.planner == 0 || O.planner == 6 || O.planner == 7)) go = (t - ghi[HI_LAST_PLAN_STEP]) >= iv || ev_hit;
        else if (O.mode == 3) go = 1;
        else go = t == 0 || (t % iv) == 0 || ev_hit;
        // allocate_tasks counts every call, also those that return without replanning (HungarianAllocator.py:84)
        if (!go && (O.planner == 0 || O.planner == 6 || O.planner == 7) && O.mode != 0) ghi[HI_N_CALLS] += 1;
      }
      if (!go) {
        if (P.out.d_n_pairs) P.out.d_n_pairs[e] = 0;
        if (P.opts.d_n_bundle_pairs) P.opts.d_n_bundle_pairs[e] = 0;
        if (P.actions_out) P.actions_out[(size_t)e * L.D.A * 2] = -1;
      }
    }
    go = __shfl_sync(0xffffffffu, go, 0);
    if (!go) has_env = false;
  }

  // ---- stage the record into shared memory
  if (has_env) {
    if (P.use_bulk) {
      if (lane == 0) {
        mbar_init(&bar[w], 1);
        fence_mbar_init();
      }
      __syncwarp();
      if (lane == 0) {
        mbar_expect_tx(&bar[w], (uint32_t)stage_bytes);
        bulk_g2s(rec, grec, (uint32_t)stage_bytes, &bar[w]);
      }
      mbar_wait(&bar[w], 0);
    } else {
      const uint4* src = (const uint4*)grec;
      uint4* dst = (uint4*)rec;
      for (int i = lane; i < stage_bytes / 16; i += 32) dst[i] = src[i];
    }
  }
  __syncwarp();

  Sim S;
  S.V.base = rec;     // hot part: shared memory
#if defined(MUAV_STAGE_COLD_FIXED)
  S.V.cbase = MUAV_STAGE_COLD_FIXED ? rec : grec;       // cold part: staged too, or in place in HBM (muav_layout.h)
#else
  S.V.cbase = stage_bytes > L.hot_bytes ? rec : grec;
#endif
  S.V.set_layout(&L);
  S.Cp = &P.cfg;
  S.tape = P.tapes + (size_t)(has_env ? e : 0) * P.tape_stride;
  S.scratch = scratch;
  S.out_events = (P.out.d_events && has_env) ? P.out.d_events + (size_t)e * L.D.EVC : nullptr;
  S.n_out_events = 0;
  S.step_reward = 0.0;
  S.phase_cycles = nullptr;
#if defined(MUAV_PHASE_TIMING)
  __shared__ long long phase_sh[16];
  if (threadIdx.x < 16) phase_sh[threadIdx.x] = 0;
  __syncthreads();
  if (lane == 0 && w == 0) S.phase_cycles = phase_sh;
  long long t_begin = clock64();
#endif
  View& V = S.V;
  const int A = L.D.A;

  // Warm L1 with the few global lines the sequential code will touch (RNG tape at the current cursors, this
  // environment's edge-score tile): lane 0 would otherwise eat a full DRAM/L2 latency per access.
  if (has_env) {
    if (lane < 3 && P.tapes) {
      int off = 0;
      for (int s = 0; s < lane; ++s) off += P.cfg.tape_words[s];
      const uint32_t* p = S.tape + off + V.hi()[HI_CUR_AGENT + lane];
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 32));
    }
    if (P.opts.d_edge_scores && P.opts.mode != 0 && !P.opts.score_f64) {
      const int tile_bytes = P.opts.score_rows * P.opts.score_cols * 4;
      const char* sp = (const char*)P.opts.d_edge_scores + (size_t)e * tile_bytes;
      for (int o = lane * 128; o < tile_bytes; o += 32 * 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + o));
    }
  }

#if defined(MUAV_LEAN) && !defined(MUAV_LEAN_PLANNER)
  if (false) {
#else
  if (P.alloc_only && has_env) {
#endif
    const int np = HIv(DONE) ? 0 : plan_and_allocate(S, P.opts, e, act_agent, act_tid, lane, 32);
    if (lane == 0) {
      if (P.out.d_n_pairs) P.out.d_n_pairs[e] = np;
      int n_act = 0;
      for (int i = 0; i < np; ++i) {
        if (P.out.d_pairs) P.out.d_pairs[(size_t)e * A + i] = ((int)act_agent[i] << 16) | (int)act_tid[i];
        if (P.actions_out && P.actions_are_ids) {
          if (HIv(N_OPEN) > 0 && S.in_last_open(act_tid[i])) {
            P.actions_out[((size_t)e * A + n_act) * 2] = act_agent[i];
            P.actions_out[((size_t)e * A + n_act) * 2 + 1] = act_tid[i];
            ++n_act;
          }
        } else if (P.actions_out && HIv(N_OPEN) > 0 && S.in_last_open(act_tid[i])) {
          // index of the task inside last_tasks_info = number of open tasks with a smaller id
          int k = act_tid[i] - 1, idx = 0;
          for (int ww = 0; ww < (k >> 5); ++ww) idx += __popc(V.open_mask()[ww]);
          idx += __popc(V.open_mask()[k >> 5] & ((1u << (k & 31)) - 1u));
          P.actions_out[((size_t)e * A + n_act) * 2] = act_agent[i];
          P.actions_out[((size_t)e * A + n_act) * 2 + 1] = idx;
          ++n_act;
        }
      }
      if (P.actions_out && n_act < A) P.actions_out[((size_t)e * A + n_act) * 2] = -1;
    }
    __syncwarp();
  }

  for (int s = 0; s < P.n_steps; ++s) {
    const bool alive = has_env && !HIv(DONE);
    int np = 0;
#if defined(MUAV_PHASE_TIMING)
    long long t_al = clock64();
#endif
    if (alive && P.opts.mode != 0) np = plan_and_allocate(S, P.opts, e, act_agent, act_tid, lane, 32);
#if defined(MUAV_PHASE_TIMING)
    if (lane == 0 && w == 0) phase_sh[0] += clock64() - t_al;
#endif
    if (alive && lane == 0) {
      int n_act = 0;
      if (P.opts.mode != 0) {
        if (P.out.d_n_pairs) P.out.d_n_pairs[e] = np;
        if (P.out.d_pairs)
          for (int i = 0; i < np; ++i) P.out.d_pairs[(size_t)e * A + i] = ((int)act_agent[i] << 16) | (int)act_tid[i];
        // _apply_assign (wps_eval.py:55-61): only tasks present in last_tasks_info become actions
        for (int i = 0; i < np; ++i) {
          if (HIv(N_OPEN) > 0 && S.in_last_open(act_tid[i])) {
            act_agent[n_act] = act_agent[i];
            act_tid[n_act] = act_tid[i];
            ++n_act;
          }
        }
      } else if (P.actions) {
        const int32_t* act = P.actions + (size_t)e * A * 2;
        for (int i = 0; i < A; ++i) {
          int a = act[2 * i];
          if (a < 0) break;
          act_agent[n_act] = (int16_t)a;
          act_tid[n_act] = (int16_t)(P.actions_are_ids ? act[2 * i + 1] : S.open_task_at(act[2 * i + 1]));
          ++n_act;
        }
      }
      n_act_s[w] = n_act;
    }
    __syncwarp();
    MUAV_CTA_SYNC(sync_mask & 1);
    StepResult r = S.step(act_agent, act_tid, alive ? n_act_s[w] : 0, lane, 32, alive, sync_mask);
    if (alive && lane == 0) {
      if (P.out.d_reward) P.out.d_reward[e] = r.reward;
      if (P.out.d_terminated) P.out.d_terminated[e] = (uint8_t)r.terminated;
      if (P.out.d_truncated) P.out.d_truncated[e] = (uint8_t)r.truncated;
      if (P.out.d_n_events) P.out.d_n_events[e] = S.n_out_events;
      if (P.out.d_n_open) P.out.d_n_open[e] = HIv(N_OPEN);
    }
    __syncwarp();
    if (P.tok.d_need && has_env) {
      const int iv = P.tok.interval > 0 ? P.tok.interval : 1;
      const bool need = !HIv(DONE) && ((HIv(T) % iv) == 0 || (HIv(EV_TAGMASK) & P.tok.event_mask) != 0);
      if (lane == 0) P.tok.d_need[e] = need ? 1 : 0;
      if (need && P.tok.d_task_feats && MUAV_F_ESCORT(P.cfg.escort_enabled) && P.tok.agent_feat_dim == 16) {
        // escort tokens (build_escort_tokens, AttentionEscort.py:76-241): task features [., 22], agent features [., 16]
        const int mt = P.tok.max_tasks, ma = P.tok.max_agents;
        EscortTokScratch W = carve_escort_tok((char*)scratch, V.lay().D.TC, mt);
        tokens_escort_env(V, P.cfg, mt, ma, P.tok.d_task_feats + (size_t)e * mt * 22, P.tok.d_task_mask + (size_t)e * mt,
                          P.tok.d_agent_feats + (size_t)e * ma * 16, P.tok.d_agent_mask + (size_t)e * ma,
                          P.tok.d_edge_valid + (size_t)e * ma * mt, P.tok.d_task_ids + (size_t)e * mt,
                          P.tok.d_task_order ? P.tok.d_task_order + (size_t)e * V.lay().D.IC : nullptr, W, lane, 32);
      } else if (need && P.tok.d_task_feats) {
        const int mt = P.tok.max_tasks, ma = P.tok.max_agents;
        const int afd = P.tok.agent_feat_dim == 13 ? 13 : 12;   // 13: commit tokens (enrich_commit_tokens)
        tokens_pair_env(V, P.cfg, mt, ma, P.tok.d_task_feats + (size_t)e * mt * 13, P.tok.d_task_mask + (size_t)e * mt,
                        P.tok.d_agent_feats + (size_t)e * ma * afd, P.tok.d_agent_mask + (size_t)e * ma,
                        P.tok.d_edge_valid ? P.tok.d_edge_valid + (size_t)e * ma * mt : nullptr,
                        P.tok.d_task_ids + (size_t)e * mt, (int16_t*)scratch, lane, 32, afd, 0,
                        P.tok.d_context ? P.tok.d_context + (size_t)e * 8 : nullptr);
      }
      __syncwarp();
    }
  }

  // ---- launch slot of this environment in the next launch: replanning environments first
  if (P.out.d_env_order_next && has_env && lane == 0 && !P.alloc_only) {
    const muav_alloc_opts& O = P.opts;
    bool will = false;
    const int rule = O.mode != 0 ? O.mode : O.order_hint_mode;
    if (!HIv(DONE) && rule != 0) {
      const int t = HIv(T);
      const int iv = O.replan_interval > 0 ? O.replan_interval : 1;
      const bool ev_hit = (HIv(EV_TAGMASK) & O.event_mask) != 0;
      if (rule == 1 && (O.planner == 0 || O.planner == 6 || O.planner == 7)) will = (t - HIv(LAST_PLAN_STEP)) >= iv || ev_hit;
      else if (rule == 3) will = true;
      else will = t == 0 || (t % iv) == 0 || ev_hit;
    }
    int32_t* nx = P.out.d_env_order_next;
    const int pos = atomicAdd(&nx[P.n_envs + (will ? 0 : 1)], 1);
    if (pos >= 0 && pos < P.n_envs) nx[will ? pos : P.n_envs - 1 - pos] = e;
  }
#if defined(MUAV_PHASE_TIMING)
  __syncthreads();
  if (lane == 0 && w == 0 && P.phase_out) {
    phase_sh[15] = clock64() - t_begin;
    for (int i = 0; i < 16; ++i) atomicAdd((unsigned long long*)&P.phase_out[i], (unsigned long long)phase_sh[i]);
  }
#endif
  // ---- write the record back
  __syncwarp();
  if (has_env) {
    if (P.use_bulk) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        bulk_s2g(grec, rec, (uint32_t)stage_bytes);
        bulk_wait_all();
      }
    } else {
      const uint4* src = (const uint4*)rec;
      uint4* dst = (uint4*)grec;
      for (int i = lane; i < stage_bytes / 16; i += 32) dst[i] = src[i];
    }
  }
}

#if defined(MUAV_STEP_ONLY)
}  // namespace muav (renamed by the including translation unit)

// launcher of this translation unit's instantiation of the step kernel (see muav_step_lean.cu)
// The opt-in shared-memory size is a per-device attribute of the function: set to the device maximum once per device
// (concurrent first calls write the same value).
static int muav_inst_prepare() {
  static bool done[MUAV_MAX_DEVICES];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < MUAV_MAX_DEVICES && done[dev]) return 0;
  // the opt-in limit covers static + dynamic shared memory of a CTA
  int optin = 0;
  cudaFuncAttributes fa;
  cudaError_t e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, muav::muav_step_kernel);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(muav::muav_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             optin - (int)fa.sharedSizeBytes);
  if (e != cudaSuccess) return -1000 - (int)e;
  if (dev >= 0 && dev < MUAV_MAX_DEVICES) done[dev] = true;
  return 0;
}
extern "C" int MUAV_STEP_LAUNCHER(const void* params, int grid, int threads, size_t smem, void* stream) {
  if (threads > MUAV_LB_THREADS) return -22;
  const int rc = muav_inst_prepare();
  if (rc) return rc;
  muav::muav_step_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(*(const muav::StepParams*)params);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - (int)e;
}
extern "C" int MUAV_STEP_STATIC_SMEM(void) {
  cudaFuncAttributes fa;
  return cudaFuncGetAttributes(&fa, muav::muav_step_kernel) == cudaSuccess ? (int)fa.sharedSizeBytes : 4608;
}
// resident CTAs per SM for a CTA of `threads` threads with `smem` bytes of dynamic shared memory (registers, shared
// memory and warp slots taken into account); 0 when the instantiation cannot be launched that wide
extern "C" int MUAV_STEP_OCC(int threads, size_t smem) {
  if (threads > MUAV_LB_THREADS || muav_inst_prepare() != 0) return 0;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, muav::muav_step_kernel, threads, smem) != cudaSuccess) return 0;
  return nb;
}
#if defined(MUAV_FIXED_SHAPE)
// the one record shape this instantiation was compiled for: A, TC, IC, HC, QC, EVC, NOBS
extern "C" void MUAV_STEP_SHAPE(int* out) {
  const int v[7] = {MUAV_FIXED_SHAPE};
  for (int i = 0; i < 7; ++i) out[i] = v[i];
}
#endif
#else
// ------------------------------------------------------------------ standalone batched LSAP
__global__ void __launch_bounds__(32) muav_lsap_kernel(const double* cost, const int32_t* nr_arr, const int32_t* nc_arr,
                                                       int nr_max, int nc_max, int32_t* col4row, int n) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int b = blockIdx.x;
  if (b >= n) return;
  const int lane = threadIdx.x;
  const int nr = nr_arr[b], nc = nc_arr[b];
  AllocScratch W = carve_scratch((char*)smem, nr_max, nc_max);
  const double* src = cost + (size_t)b * nr_max * nc_max;
  for (int idx = lane; idx < nr * nc; idx += 32) {
    int i = idx / nc, j = idx - i * nc;
    W.cost[idx] = src[i * nc_max + j];
  }
  __syncwarp();
  const bool ok = (nr > 0 && nc > 0) ? lsap_solve(W.cost, nr, nc, W, W.col_of_row, lane, 32) : true;
  for (int i = lane; i < nr_max; i += 32)
    col4row[(size_t)b * nr_max + i] = (ok && i < nr && nr > 0 && nc > 0) ? W.col_of_row[i] : -1;
}

__global__ void muav_avoid_kernel(const double* pos, const double* mv, const double* obst, int nobs, double* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double ax, ay;
  Sim::avoid_obstacles(obst, nobs, pos[2 * i], pos[2 * i + 1], mv[2 * i], mv[2 * i + 1], &ax, &ay);
  out[2 * i] = ax;
  out[2 * i + 1] = ay;
}

__global__ void muav_metrics_kernel(const __grid_constant__ muav_config cfg, const __grid_constant__ Layout L,
                                    const char* records, double* out, int n) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  View V;
  V.at((char*)records + (size_t)e * L.record_bytes);
  V.set_layout(&L);
  metrics_env(V, cfg, out + (size_t)e * MUAV_N_METRICS);
}

__global__ void __launch_bounds__(128) muav_tokens_pair_kernel(const __grid_constant__ muav_config cfg,
                                                               const __grid_constant__ Layout L, const char* records,
                                                               int max_tasks, int max_agents, float* tf, uint8_t* tm,
                                                               float* af, uint8_t* am, float* ev, int32_t* ids, int n,
                                                               int af_dim, int raw, float* ctx) {
  __shared__ int16_t cols[4][MUAV_MAX_TASK_CAP + 2];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * 4 + w;
  if (e >= n) return;
  View V;
  V.at((char*)records + (size_t)e * L.record_bytes);
  V.set_layout(&L);
  const int TD = raw ? 9 : 13, AD = raw ? 11 : af_dim, CD = raw ? 1 : 8;
  tokens_pair_env(V, cfg, max_tasks, max_agents, tf + (size_t)e * max_tasks * TD, tm + (size_t)e * max_tasks,
                  af + (size_t)e * max_agents * AD, am + (size_t)e * max_agents,
                  ev ? ev + (size_t)e * max_agents * max_tasks : nullptr, ids + (size_t)e * max_tasks, cols[w], lane, 32,
                  af_dim, raw, ctx ? ctx + (size_t)e * CD : nullptr);
}

__global__ void __launch_bounds__(128) muav_tokens_escort_kernel(const __grid_constant__ muav_config cfg,
                                                                 const __grid_constant__ Layout L, const char* records,
                                                                 int max_tasks, int max_agents, float* tf, uint8_t* tm,
                                                                 float* af, uint8_t* am, float* ev, int32_t* ids,
                                                                 int32_t* order, int n, int per_warp) {
  extern __shared__ __align__(16) char esc_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * 4 + w;
  if (e >= n) return;
  View V;
  V.at((char*)records + (size_t)e * L.record_bytes);
  V.set_layout(&L);
  EscortTokScratch W = carve_escort_tok(esc_smem + (size_t)w * per_warp, L.D.TC, max_tasks);
  tokens_escort_env(V, cfg, max_tasks, max_agents, tf + (size_t)e * max_tasks * 22, tm + (size_t)e * max_tasks,
                    af + (size_t)e * max_agents * 16, am + (size_t)e * max_agents, ev + (size_t)e * max_agents * max_tasks,
                    ids + (size_t)e * max_tasks, order ? order + (size_t)e * L.D.IC : nullptr, W, lane, 32);
}

// Pair mask of the trainers: mask[i, j] = 1 for every allocator pair (agent, task) whose agent is token row i (i-th live
// agent) and whose task is token column j (task_ids[j]); require_valid keeps only pairs with edge_valid >= 0.5
// (_expert_mask, experiments/train_pair_cost.py:53-70) -- without it this is PairCostHybrid._selected_mask (:293-306).
__global__ void muav_pair_mask_kernel(const __grid_constant__ Layout L, const char* records, const int32_t* pairs,
                                      const int32_t* n_pairs, const int32_t* task_ids, const float* edge_valid,
                                      int max_tasks, int max_agents, int require_valid, float* mask, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  View V;
  V.at((char*)records + (size_t)e * L.record_bytes);
  V.set_layout(&L);
  const int A = L.D.A;
  float* m = mask + (size_t)e * max_agents * max_tasks;
  for (int i = 0; i < max_agents * max_tasks; ++i) m[i] = 0.0f;
  const int32_t* ids = task_ids + (size_t)e * max_tasks;
  const int np = n_pairs[e];
  for (int p = 0; p < np && p < A; ++p) {
    const int v = pairs[(size_t)e * A + p];
    const int a = v >> 16, tid = v & 0xFFFF;
    if (a < 0 || a >= A || V.a_state()[a] == -1) continue;
    int row = 0;
    for (int b = 0; b < a; ++b) row += V.a_state()[b] != -1;
    if (row >= max_agents) continue;
    int col = -1;
    for (int j = 0; j < max_tasks; ++j)
      if (ids[j] == tid && tid != 0) col = j;   // dict built in column order: the last duplicate would win (ids are unique)
    if (col < 0) continue;
    if (require_valid && !(edge_valid[((size_t)e * max_agents + row) * max_tasks + col] >= 0.5f)) continue;
    m[row * max_tasks + col] = 1.0f;
  }
}

__global__ void muav_observe_kernel(const __grid_constant__ muav_config cfg, const __grid_constant__ Layout L,
                                    const char* records, int max_rows, double* ti, uint8_t* pad, uint8_t* legal,
                                    double* ao, float* ef, int32_t* n_rows, int n) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  View V;
  V.at((char*)records + (size_t)e * L.record_bytes);
  V.set_layout(&L);
  int32_t nr = 0;
  observe_env(V, cfg, max_rows, ti + (size_t)e * max_rows * MUAV_OBS_TASK_DIM, pad + (size_t)e * max_rows,
              legal + (size_t)e * L.D.A * max_rows, ao + (size_t)e * L.D.A * MUAV_OBS_AGENT_DIM, ef + (size_t)e * 5, &nr);
  if (n_rows) n_rows[e] = nr;
}

}  // namespace muav

// ====================================================================== C ABI
using namespace muav;

static int cuda_rc(cudaError_t e) { return e == cudaSuccess ? 0 : -1000 - (int)e; }

extern "C" {

#include "muav_abi_common.inl"

// Instantiations of the step kernel in their own translation units (same source, namespaces of their own):
//   lean / lean_escort (muav_step_lean*.cu): no obstacles, plain Hungarian allocator, escorts off / always on;
//   fixed-shape ones (muav_step_hard.cu, ...): the lean feature set AND the record dimensions as compile-time constants.
#define MUAV_DECL_INST(n)                                                                        \
  int muav_step_##n##_launch(const void* params, int grid, int threads, size_t smem, void* stream); \
  int muav_step_##n##_static_smem(void);                                                         \
  int muav_step_##n##_occ(int threads, size_t smem);
#define MUAV_DECL_SHAPED(n) MUAV_DECL_INST(n) void muav_step_##n##_shape(int* out);
MUAV_DECL_INST(lean)
MUAV_DECL_INST(lean_escort)
MUAV_DECL_SHAPED(hard)
MUAV_DECL_SHAPED(hard32)
MUAV_DECL_SHAPED(commit)
MUAV_DECL_SHAPED(escort)
MUAV_DECL_SHAPED(commit_planner)
MUAV_DECL_SHAPED(escort_planner)
MUAV_DECL_SHAPED(hard32_planner)
MUAV_DECL_SHAPED(burst2)
MUAV_DECL_SHAPED(burst4)
MUAV_DECL_SHAPED(burst8)

struct StepInst {
  int (*launch)(const void*, int, int, size_t, void*);
  int (*occ)(int, size_t);
  void (*shape)(int*);  // null: any shape
  int escort;           // value of cfg.escort_enabled this instantiation was compiled for
  int stage_cold;       // MUAV_STAGE_COLD_FIXED of the instantiation: 1 whole record staged, 0 hot part only
  int planners;         // 1: compiled with the planner front ends / market allocators (any muav_alloc_opts.planner) and
                        //    the allocator-only launch (muav_allocate, the split allocator -> step path)
};
static const StepInst kStepInst[] = {
    {muav_step_hard_launch, muav_step_hard_occ, muav_step_hard_shape, 0, 1, 0},
    {muav_step_hard32_launch, muav_step_hard32_occ, muav_step_hard32_shape, 0, 1, 0},
    {muav_step_commit_launch, muav_step_commit_occ, muav_step_commit_shape, 0, 0, 0},
    {muav_step_escort_launch, muav_step_escort_occ, muav_step_escort_shape, 1, 0, 0},
    {muav_step_burst2_launch, muav_step_burst2_occ, muav_step_burst2_shape, 0, 0, 0},
    {muav_step_burst4_launch, muav_step_burst4_occ, muav_step_burst4_shape, 0, 0, 0},
    {muav_step_burst8_launch, muav_step_burst8_occ, muav_step_burst8_shape, 0, 0, 0},
    {muav_step_lean_launch, muav_step_lean_occ, nullptr, 0, 0, 0},
    {muav_step_lean_escort_launch, muav_step_lean_escort_occ, nullptr, 1, 0, 0},
    {muav_step_hard32_planner_launch, muav_step_hard32_planner_occ, muav_step_hard32_planner_shape, 0, 1, 1},
    {muav_step_commit_planner_launch, muav_step_commit_planner_occ, muav_step_commit_planner_shape, 0, 0, 1},
    {muav_step_escort_planner_launch, muav_step_escort_planner_occ, muav_step_escort_planner_shape, 1, 0, 1},
};

// general kernel of this translation unit: same two services
static int general_prepare() {
  static bool done[MUAV_MAX_DEVICES];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < MUAV_MAX_DEVICES && done[dev]) return 0;
  int optin = 0;
  cudaFuncAttributes fa;
  cudaError_t e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, muav_step_kernel);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(muav_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
  if (e != cudaSuccess) return cuda_rc(e);
  if (dev >= 0 && dev < MUAV_MAX_DEVICES) done[dev] = true;
  return 0;
}
static int general_occ(int threads, size_t smem) {
  if (general_prepare() != 0) return 0;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, muav_step_kernel, threads, smem) != cudaSuccess) return 0;
  return nb;
}

// the most specialised instantiation that covers this launch (null: the general kernel of this translation unit)
static const StepInst* pick_inst(const StepParams& P) {
  const bool lean = P.L.D.NOBS == 0 && !getenv("MUAV_NO_LEAN");
  if (!lean) return nullptr;
  const bool no_fixed = getenv("MUAV_NO_FIXED_SHAPE") != nullptr;
  const Dims& D = P.L.D;
  const int have[7] = {D.A, D.TC, D.IC, D.HC, D.QC, D.EVC, D.NOBS};
  for (const StepInst& I : kStepInst) {
    if (I.escort != (P.cfg.escort_enabled ? 1 : 0)) continue;
    if ((P.opts.planner != 0 || P.alloc_only) && !I.planners) continue;   // front ends / the allocator-only service
    if (I.shape) {
      if (no_fixed) continue;
      int want[7];
      I.shape(want);
      if (want[2] < want[1]) want[2] = want[1];
      bool same = true;
      for (int i = 0; i < 7; ++i) same = same && want[i] == have[i];
      if (!same) continue;
    }
    return &I;
  }
  return nullptr;
}

// shared memory per resident environment of this launch: [hot record | scratch | ordered action list]; the scratch holds
// the allocator's work arrays (+ the planner front ends' arrays when one runs) or just the step's temporaries
static size_t slot_bytes_of(StepParams& P) {
  const bool with_alloc = P.alloc_only || P.opts.mode != 0;
  const bool with_planner = P.opts.planner != 0 && P.opts.planner != 6;
  P.scratch_launch = !with_alloc ? P.L.step_scratch_bytes : (with_planner ? P.L.scratch_bytes : P.L.plain_scratch_bytes);
  if (with_alloc && P.opts.planner == 7) {   // CBBA keeps an MT19937 state, a set table and per-agent paths
    const int cb = cbba_scratch_bytes(P.L.D.A);
    if (cb > P.scratch_launch) P.scratch_launch = cb;
  }
  if (P.tok.d_need && P.tok.d_task_feats && P.tok.agent_feat_dim == 16) {   // escort tokens sort their candidates
    const int eb = (int)escort_tok_scratch_bytes(P.L.D.TC, P.tok.max_tasks);
    if (eb > P.scratch_launch) P.scratch_launch = eb;
  }
  return (size_t)P.stage_bytes + (size_t)P.scratch_launch + (size_t)P.L.act_bytes;
}

// Environments (warps) per CTA.  The warps of a CTA are phase-aligned with CTA barriers: they share instruction
// fetches, but a CTA is as slow as its slowest environment.  Measured on B200 (profiles/r01_cta_width.md,
// profiles/r02_step_kernel.md): the number of resident environments per SM decides first (registers, shared memory and
// warp slots: asked from the occupancy calculator per instantiation); at equal residency the widest CTA that still
// leaves two CTAs per SM wins; when only one-warp CTAs or a single wide CTA reach the maximum, the wide CTA wins from
// four environments up and one-warp CTAs below.  Returns W, *envs_per_sm = resident environments per SM with it.
static int choose_width(const StepInst* inst, size_t slot, int* envs_per_sm) {
  // the answer depends on (instantiation, slot size) only: remembered in a small table (written under a lock)
  struct Memo { const StepInst* inst; size_t slot; int W, envs; };
  static Memo memo[32];
  static int n_memo = 0;
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> g(mu);
    for (int i = 0; i < n_memo; ++i)
      if (memo[i].inst == inst && memo[i].slot == slot) {
        if (envs_per_sm) *envs_per_sm = memo[i].envs;
        return memo[i].W;
      }
  }
  const int WMAX = 12;
  int envs[WMAX + 1], ctas[WMAX + 1], best = 0;
  for (int w = 0; w <= WMAX; ++w) envs[w] = ctas[w] = 0;
  for (int w = 1; w <= WMAX; ++w) {
    if (slot * w > 226 * 1024) break;
    ctas[w] = inst ? inst->occ(32 * w, slot * w) : general_occ(32 * w, slot * w);
    envs[w] = ctas[w] * w;
    if (envs[w] > best) best = envs[w];
  }
  int pick = 0;
  for (int w = WMAX; w >= 2 && !pick; --w)
    if (envs[w] == best && ctas[w] >= 2) pick = w;
  if (!pick) {
    int wide = 0;
    for (int w = WMAX; w >= 2 && !wide; --w)
      if (envs[w] == best) wide = w;
    pick = (wide && (best >= 4 || envs[1] < best)) ? wide : 1;
  }
  std::lock_guard<std::mutex> g(mu);
  if (n_memo < 32) memo[n_memo++] = Memo{inst, slot, pick, envs[pick]};
  if (envs_per_sm) *envs_per_sm = envs[pick];
  return pick;
}

static int launch_step(StepParams& P, void* stream) {
  const StepInst* inst = pick_inst(P);
  // Hot part only, or the whole record?  Staging the cold part costs shared memory (fewer resident environments) but
  // turns the allocator's / the action phase's updates of the requirement vectors and allocation times into
  // shared-memory accesses.  Measured on B200 (profiles/r02_step_kernel.md): while the whole record still leaves >= 12
  // environments per SM (WPS_easy / hard / burst) staging everything is ~15 % faster; for the larger records (WPS_commit,
  // WPS_escort, the scaled bursts) the residency of the hot-only slot wins by 20-40 %.  MUAV_STAGE_COLD=0 / 1 forces it.
  // The specialised instantiations carry the answer for their family; the general kernel decides at run time.
  if (inst) {
    P.stage_bytes = inst->stage_cold ? P.L.record_bytes : P.L.hot_bytes;
  } else {
    P.stage_bytes = P.L.record_bytes;
    int envs_full = 0;
    choose_width(inst, slot_bytes_of(P), &envs_full);
    const char* fc = getenv("MUAV_STAGE_COLD");
    const bool stage_cold = fc ? atoi(fc) != 0 : envs_full >= 12;
    if (!stage_cold) P.stage_bytes = P.L.hot_bytes;
  }
  const size_t slot = slot_bytes_of(P);
  int W = choose_width(inst, slot, nullptr);
  const char* ev = getenv("MUAV_CTA_WARPS");
  if (ev) W = atoi(ev);
  if (W < 1) W = 1;
  if (W > MUAV_MAX_CTA_WARPS) W = MUAV_MAX_CTA_WARPS;
  while (W > 1 && slot * W > 200 * 1024) --W;
  if (slot > 226 * 1024) return -12;
  P.cta_warps = W;
  // phase-alignment barriers (bit 0: before the step, 1..3: after its three sequential parts, 4: end of step).
  // Re-measured on the final kernels (profiles/r02_step_kernel.md): the barriers before the step and after the second and
  // third part pay for themselves (shared instruction fetch), the ones after the first part and at the end of the step
  // only add waiting: 13 beats 31 by 1.6-5.8 % on every BASELINE shape
  P.sync_mask = 13;
  const char* sm = getenv("MUAV_SYNC_MASK");
  if (sm) P.sync_mask = atoi(sm);
  const size_t smem = slot * W;
  if (inst) return inst->launch(&P, (P.n_envs + W - 1) / W, 32 * W, smem, stream);
  {
    const int prc = general_prepare();
    if (prc) return prc;
  }
  const int grid = (P.n_envs + W - 1) / W;
  muav_step_kernel<<<grid, 32 * W, smem, (cudaStream_t)stream>>>(P);
  return cuda_rc(cudaGetLastError());
}

int muav_step(const muav_config* cfg, void* d_records, const uint32_t* d_tapes, const int32_t* d_actions,
              const muav_alloc_opts* opts, const muav_step_out* out, const muav_token_out* tok, int n_envs, int n_steps,
              void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!d_records || n_envs < 0 || n_steps < 0) return -22;
  if (n_envs == 0 || n_steps == 0) return 0;
  StepParams P;
  memset(&P, 0, sizeof(P));
  P.cfg = *cfg;
  P.L = make_layout(*cfg);
  if (opts) P.opts = *opts;
  if (out) P.out = *out;
  if (tok) {
    if (tok->max_tasks < 1 || tok->max_agents < 1 || tok->max_tasks > (cfg->task_cap > 64 ? cfg->task_cap : 64)) return -22;
    if (tok->agent_feat_dim == 16 && tok->d_task_feats && (!cfg->escort_enabled || !tok->d_edge_valid)) return -22;
    P.tok = *tok;
  }
  P.records = (char*)d_records;
  P.tapes = d_tapes;
  P.actions = d_actions;
  P.n_envs = n_envs;
  P.n_steps = n_steps;
  P.tape_stride = cfg->tape_words[0] + cfg->tape_words[1] + cfg->tape_words[2];
  const char* st = getenv("MUAV_STAGE");
  P.use_bulk = !(st && strcmp(st, "ldst") == 0);
#if defined(MUAV_PHASE_TIMING)
  {
    static long long* d_phase = nullptr;
    if (!d_phase) {
      cudaMalloc(&d_phase, 16 * sizeof(long long));
      cudaMemset(d_phase, 0, 16 * sizeof(long long));
    }
    P.phase_out = d_phase;
    const char* dump = getenv("MUAV_PHASE_DUMP");
    if (dump) {
      long long h[16];
      cudaMemcpy(h, d_phase, sizeof(h), cudaMemcpyDeviceToHost);
      for (int i = 0; i < 16; ++i) fprintf(stderr, "phase %d %lld\n", i, h[i]);
      cudaMemset(d_phase, 0, 16 * sizeof(long long));
    }
  }
#endif
  bool split = false;
  if (n_steps == 1 && P.opts.mode != 0 && P.out.d_actions_ws) {
    // One fused kernel, or the allocator for the environments that replan followed by the step for everyone?  The
    // step-only launch needs no allocator scratch, so more environments fit an SM; measured on B200 the split pays
    // when that residency is at least 1.5 x the fused one (WPS_escort, burst x4 / x8), and costs ~15 % when both fit
    // equally (WPS_hard): profiles/r02_configs.md.  MUAV_SPLIT_STEP=0 / 1 forces the answer.
    const char* fs = getenv("MUAV_SPLIT_STEP");
    if (fs) {
      split = atoi(fs) != 0;
    } else {
      StepParams Pf = P, Ps = P;
      Ps.opts.mode = 0;
      const StepInst* inf = pick_inst(Pf);
      const StepInst* ins = pick_inst(Ps);
      Pf.stage_bytes = (inf && inf->stage_cold) ? P.L.record_bytes : P.L.hot_bytes;
      Ps.stage_bytes = (ins && ins->stage_cold) ? P.L.record_bytes : P.L.hot_bytes;
      int envs_f = 0, envs_s = 0;
      choose_width(inf, slot_bytes_of(Pf), &envs_f);
      choose_width(ins, slot_bytes_of(Ps), &envs_s);
      split = envs_f > 0 && 2 * envs_s >= 3 * envs_f;
    }
  }
  if (split) {
    // two kernels: allocator for the environments that replan, then the step for everyone with the small scratch
    StepParams Pa = P;
    Pa.n_steps = 0;
    Pa.alloc_only = 1;
    Pa.actions_out = P.out.d_actions_ws;
    Pa.actions_are_ids = 1;
    Pa.out.d_env_order_next = nullptr;
    memset(&Pa.tok, 0, sizeof(Pa.tok));
    rc = launch_step(Pa, stream);
    if (rc) return rc;
    StepParams Ps = P;
    Ps.actions = P.out.d_actions_ws;
    Ps.actions_are_ids = 1;
    Ps.opts.order_hint_mode = P.opts.mode;
    Ps.opts.mode = 0;
    Ps.out.d_n_pairs = nullptr;  // written by the allocator launch
    Ps.out.d_pairs = nullptr;
    return launch_step(Ps, stream);
  }
  return launch_step(P, stream);
}

int muav_allocate(const muav_config* cfg, void* d_records, const muav_alloc_opts* opts, const muav_step_out* out,
                  int32_t* d_actions_out, int n_envs, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!d_records || !opts || n_envs < 0) return -22;
  if (n_envs == 0) return 0;
  StepParams P;
  memset(&P, 0, sizeof(P));
  P.cfg = *cfg;
  P.L = make_layout(*cfg);
  P.opts = *opts;
  if (out) P.out = *out;
  P.records = (char*)d_records;
  P.actions_out = d_actions_out;
  P.n_envs = n_envs;
  P.n_steps = 0;
  P.alloc_only = 1;
  const char* st = getenv("MUAV_STAGE");
  P.use_bulk = !(st && strcmp(st, "ldst") == 0);
  return launch_step(P, stream);
}

int muav_rollout(const muav_config* cfg, void* d_records, const uint32_t* d_tapes, const muav_alloc_opts* opts,
                 const muav_step_out* out, const muav_token_out* tok, int n_envs, int n_steps, void* stream) {
  if (!opts || opts->mode == 0) return -22;  // a rollout needs the fused allocator
  return muav_step(cfg, d_records, d_tapes, nullptr, opts, out, tok, n_envs, n_steps, stream);
}

size_t muav_state_bytes(const muav_config* cfg, int n_envs) {
  if (check_cfg(cfg) || n_envs < 0) return 0;
  return muav_record_bytes(cfg) * (size_t)n_envs;
}

size_t muav_tape_bytes(const muav_config* cfg, int n_envs) {
  if (check_cfg(cfg) || n_envs < 0) return 0;
  return (size_t)(cfg->tape_words[0] + cfg->tape_words[1] + cfg->tape_words[2]) * sizeof(uint32_t) * (size_t)n_envs;
}

int muav_reset_upload(const muav_config* cfg, void* d_records, uint32_t* d_tapes, const void* h_records,
                      const uint32_t* h_tapes, int n_envs, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (n_envs <= 0) return n_envs == 0 ? 0 : -22;
  if (!d_records || !d_tapes || !h_records || !h_tapes) return -22;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpyAsync(d_records, h_records, muav_state_bytes(cfg, n_envs), cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return cuda_rc(e);
  return cuda_rc(cudaMemcpyAsync(d_tapes, h_tapes, muav_tape_bytes(cfg, n_envs), cudaMemcpyHostToDevice, s));
}

int muav_snapshot(const muav_config* cfg, const void* d_records, int env_index, void* h_record, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!d_records || !h_record || env_index < 0) return -22;
  const size_t rb = muav_record_bytes(cfg);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpyAsync(h_record, (const char*)d_records + rb * (size_t)env_index, rb, cudaMemcpyDeviceToHost, s);
  if (e != cudaSuccess) return cuda_rc(e);
  return cuda_rc(cudaStreamSynchronize(s));
}

// ---- host-buffer entry points.  All staging memory belongs to a caller-owned handle (muav_ctx): the library keeps no
// mutable process-wide state, so two handles (two threads, two devices) never share anything.
struct muav_ctx {
  muav_config cfg;
  int n_envs, device;
  char* d_buf;       // [actions in | reward | terminated | truncated]
  char* h_pin;       // pinned image of the three outputs: one device -> host copy per step
  size_t off_rew, off_term, off_trunc, bytes;
};

static size_t ctx_layout(const muav_config* cfg, int n_envs, size_t* off_rew, size_t* off_term, size_t* off_trunc) {
  const size_t act_bytes = (size_t)n_envs * cfg->n_agents * 2 * sizeof(int32_t);
  *off_rew = (act_bytes + 15) / 16 * 16;
  *off_term = *off_rew + (size_t)n_envs * 8;
  *off_trunc = *off_term + (size_t)n_envs;
  return (*off_trunc + (size_t)n_envs + 15) / 16 * 16;
}

int muav_ctx_create(const muav_config* cfg, int n_envs, int device, muav_ctx** out) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!out || n_envs < 1 || device < 0) return -22;
  *out = nullptr;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_rc(e);
  muav_ctx* c = (muav_ctx*)calloc(1, sizeof(muav_ctx));
  if (!c) { cudaSetDevice(prev); return -12; }
  c->cfg = *cfg;
  c->n_envs = n_envs;
  c->device = device;
  c->bytes = ctx_layout(cfg, n_envs, &c->off_rew, &c->off_term, &c->off_trunc);
  e = cudaMalloc((void**)&c->d_buf, c->bytes);
  if (e == cudaSuccess) e = cudaMallocHost((void**)&c->h_pin, c->bytes - c->off_rew);
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    muav_ctx_destroy(c);
    return cuda_rc(e);
  }
  *out = c;
  return 0;
}

void muav_ctx_destroy(muav_ctx* c) {
  if (!c) return;
  if (c->d_buf) cudaFree(c->d_buf);
  if (c->h_pin) cudaFreeHost(c->h_pin);
  free(c);
}

// shared body: `b` = device staging block laid out by ctx_layout, `h_pin` = optional pinned image of the outputs
static int step_host_impl(const muav_config* cfg, char* b, char* h_pin, size_t off_rew, size_t off_term, size_t off_trunc,
                          void* d_records, const uint32_t* d_tapes, const int32_t* h_actions, const muav_alloc_opts* opts,
                          const muav_token_out* tok, double* h_reward, uint8_t* h_terminated, uint8_t* h_truncated,
                          int n_envs, int n_steps, cudaStream_t s, const int32_t* d_env_order, int32_t* d_env_order_next) {
  const size_t act_bytes = (size_t)n_envs * cfg->n_agents * 2 * sizeof(int32_t);
  if (h_actions) {
    cudaError_t e = cudaMemcpyAsync(b, h_actions, act_bytes, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return cuda_rc(e);
  }
  muav_step_out out;
  memset(&out, 0, sizeof(out));
  out.d_reward = (double*)(b + off_rew);
  out.d_terminated = (uint8_t*)(b + off_term);
  out.d_truncated = (uint8_t*)(b + off_trunc);
  out.d_env_order = d_env_order;
  out.d_env_order_next = d_env_order_next;
  int rc = muav_step(cfg, d_records, d_tapes, h_actions ? (const int32_t*)b : nullptr, opts, &out, tok, n_envs, n_steps,
                     (void*)s);
  if (rc) return rc;
  cudaError_t e = cudaSuccess;
  if (h_pin) {
    if (h_reward || h_terminated || h_truncated) {
      e = cudaMemcpyAsync(h_pin, b + off_rew, off_trunc + (size_t)n_envs - off_rew, cudaMemcpyDeviceToHost, s);
      if (e != cudaSuccess) return cuda_rc(e);
    }
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return cuda_rc(e);
    if (h_reward) memcpy(h_reward, h_pin, (size_t)n_envs * 8);
    if (h_terminated) memcpy(h_terminated, h_pin + (off_term - off_rew), (size_t)n_envs);
    if (h_truncated) memcpy(h_truncated, h_pin + (off_trunc - off_rew), (size_t)n_envs);
    return 0;
  }
  if (h_reward) e = cudaMemcpyAsync(h_reward, out.d_reward, (size_t)n_envs * 8, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && h_terminated)
    e = cudaMemcpyAsync(h_terminated, out.d_terminated, (size_t)n_envs, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && h_truncated)
    e = cudaMemcpyAsync(h_truncated, out.d_truncated, (size_t)n_envs, cudaMemcpyDeviceToHost, s);
  if (e != cudaSuccess) return cuda_rc(e);
  return cuda_rc(cudaStreamSynchronize(s));
}

int muav_ctx_step_host(muav_ctx* c, void* d_records, const uint32_t* d_tapes, const int32_t* h_actions,
                       const muav_alloc_opts* opts, const muav_token_out* tok, double* h_reward, uint8_t* h_terminated,
                       uint8_t* h_truncated, int n_steps, void* stream, const int32_t* d_env_order,
                       int32_t* d_env_order_next) {
  if (!c || !d_records || n_steps < 0) return -22;
  return step_host_impl(&c->cfg, c->d_buf, c->h_pin, c->off_rew, c->off_term, c->off_trunc, d_records, d_tapes, h_actions,
                        opts, tok, h_reward, h_terminated, h_truncated, c->n_envs, n_steps, (cudaStream_t)stream, d_env_order,
                        d_env_order_next);
}

int muav_ctx_allocate_host(muav_ctx* c, void* d_records, const muav_alloc_opts* opts, const muav_step_out* out,
                           int32_t* h_actions_out, void* stream) {
  if (!c || !d_records || !opts || !h_actions_out) return -22;
  int rc = muav_allocate(&c->cfg, d_records, opts, out, (int32_t*)c->d_buf, c->n_envs, stream);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t act_bytes = (size_t)c->n_envs * c->cfg.n_agents * 2 * sizeof(int32_t);
  cudaError_t e = cudaMemcpyAsync(h_actions_out, c->d_buf, act_bytes, cudaMemcpyDeviceToHost, s);
  if (e != cudaSuccess) return cuda_rc(e);
  return cuda_rc(cudaStreamSynchronize(s));
}

// Handle-free form: the staging block is a stream-ordered allocation of this call (cudaMallocAsync), so concurrent
// callers and several devices are safe; muav_ctx_step_host avoids the allocation.
int muav_step_host(const muav_config* cfg, void* d_records, const uint32_t* d_tapes, const int32_t* h_actions,
                   const muav_alloc_opts* opts, const muav_token_out* tok, double* h_reward, uint8_t* h_terminated,
                   uint8_t* h_truncated, int n_envs, int n_steps, void* stream, const int32_t* d_env_order,
                   int32_t* d_env_order_next) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (!d_records || n_envs < 0 || n_steps < 0) return -22;
  if (n_envs == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  size_t off_rew, off_term, off_trunc;
  const size_t need = ctx_layout(cfg, n_envs, &off_rew, &off_term, &off_trunc);
  void* buf = nullptr;
  cudaError_t e = cudaMallocAsync(&buf, need, s);
  if (e != cudaSuccess) return cuda_rc(e);
  rc = step_host_impl(cfg, (char*)buf, nullptr, off_rew, off_term, off_trunc, d_records, d_tapes, h_actions, opts, tok,
                      h_reward, h_terminated, h_truncated, n_envs, n_steps, s, d_env_order, d_env_order_next);
  cudaFreeAsync(buf, s);
  return rc;
}

int muav_lsap(const double* d_cost, const int32_t* d_nr, const int32_t* d_nc, int nr_max, int nc_max, int32_t* d_col4row,
              int n_problems, void* stream) {
  if (!d_cost || !d_nr || !d_nc || !d_col4row || nr_max < 1 || nc_max < 1 || n_problems < 0) return -22;
  if (n_problems == 0) return 0;
  size_t smem = (size_t)alloc_scratch_bytes(nr_max, nc_max);
  if (smem > 200 * 1024) return -7;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(muav_lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_rc(e);
  }
  muav_lsap_kernel<<<n_problems, 32, smem, (cudaStream_t)stream>>>(d_cost, d_nr, d_nc, nr_max, nc_max, d_col4row, n_problems);
  return cuda_rc(cudaGetLastError());
}

int muav_avoid_obstacles(const double* d_pos, const double* d_move, const double* d_obstacles, int n_obstacles, double* d_out,
                         int n, void* stream) {
  if (n < 0 || n_obstacles < 0) return -22;
  if (n == 0) return 0;
  muav_avoid_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_pos, d_move, d_obstacles, n_obstacles, d_out, n);
  return cuda_rc(cudaGetLastError());
}


int muav_metrics(const muav_config* cfg, const void* d_records, double* d_out, int n_envs, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (n_envs <= 0) return n_envs == 0 ? 0 : -22;
  Layout L = make_layout(*cfg);
  muav_metrics_kernel<<<(n_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*cfg, L, (const char*)d_records, d_out, n_envs);
  return cuda_rc(cudaGetLastError());
}

int muav_tokens_pair(const muav_config* cfg, const void* d_records, int max_tasks, int max_agents, float* d_task_feats,
                     uint8_t* d_task_mask, float* d_agent_feats, uint8_t* d_agent_mask, float* d_edge_valid,
                     int32_t* d_task_ids, int n_envs, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (n_envs <= 0) return n_envs == 0 ? 0 : -22;
  if (max_tasks < 1 || max_agents < 1) return -22;
  Layout L = make_layout(*cfg);
  if (max_tasks > MUAV_MAX_TASK_CAP) return -22;
  muav_tokens_pair_kernel<<<(n_envs + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      *cfg, L, (const char*)d_records, max_tasks, max_agents, d_task_feats, d_task_mask, d_agent_feats, d_agent_mask,
      d_edge_valid, d_task_ids, n_envs, 12, 0, nullptr);
  return cuda_rc(cudaGetLastError());
}

int muav_tokens_context(const muav_config* cfg, const void* d_records, int max_tasks, int max_agents, int raw,
                        float* d_task_feats, uint8_t* d_task_mask, float* d_agent_feats, uint8_t* d_agent_mask,
                        float* d_edge_valid, int32_t* d_task_ids, float* d_context, int n_envs, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (n_envs <= 0) return n_envs == 0 ? 0 : -22;
  if (max_tasks < 1 || max_agents < 1 || max_tasks > MUAV_MAX_TASK_CAP) return -22;
  if (!d_task_feats || !d_task_mask || !d_agent_feats || !d_agent_mask || !d_edge_valid || !d_task_ids || !d_context)
    return -22;
  Layout L = make_layout(*cfg);
  muav_tokens_pair_kernel<<<(n_envs + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      *cfg, L, (const char*)d_records, max_tasks, max_agents, d_task_feats, d_task_mask, d_agent_feats, d_agent_mask,
      d_edge_valid, d_task_ids, n_envs, 12, raw ? 1 : 0, d_context);
  return cuda_rc(cudaGetLastError());
}

int muav_tokens_commit(const muav_config* cfg, const void* d_records, int max_tasks, int max_agents, float* d_task_feats,
                       uint8_t* d_task_mask, float* d_agent_feats13, uint8_t* d_agent_mask, int32_t* d_task_ids, int n_envs,
                       void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (n_envs <= 0) return n_envs == 0 ? 0 : -22;
  if (max_tasks < 1 || max_agents < 1 || max_tasks > MUAV_MAX_TASK_CAP) return -22;
  Layout L = make_layout(*cfg);
  muav_tokens_pair_kernel<<<(n_envs + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      *cfg, L, (const char*)d_records, max_tasks, max_agents, d_task_feats, d_task_mask, d_agent_feats13, d_agent_mask,
      nullptr, d_task_ids, n_envs, 13, 0, nullptr);
  return cuda_rc(cudaGetLastError());
}

int muav_tokens_escort(const muav_config* cfg, const void* d_records, int max_tasks, int max_agents, float* d_task_feats22,
                       uint8_t* d_task_mask, float* d_agent_feats16, uint8_t* d_agent_mask, float* d_edge_valid,
                       int32_t* d_task_ids, int32_t* d_task_order, int n_envs, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (n_envs <= 0) return n_envs == 0 ? 0 : -22;
  if (max_tasks < 1 || max_agents < 1 || max_tasks > MUAV_MAX_TASK_CAP) return -22;
  if (!d_task_feats22 || !d_task_mask || !d_agent_feats16 || !d_agent_mask || !d_edge_valid || !d_task_ids) return -22;
  Layout L = make_layout(*cfg);
  const int per_warp = (int)escort_tok_scratch_bytes(L.D.TC, max_tasks);
  muav_tokens_escort_kernel<<<(n_envs + 3) / 4, 128, (size_t)per_warp * 4, (cudaStream_t)stream>>>(
      *cfg, L, (const char*)d_records, max_tasks, max_agents, d_task_feats22, d_task_mask, d_agent_feats16, d_agent_mask,
      d_edge_valid, d_task_ids, d_task_order, n_envs, per_warp);
  return cuda_rc(cudaGetLastError());
}

int muav_pair_mask(const muav_config* cfg, const void* d_records, const int32_t* d_pairs, const int32_t* d_n_pairs,
                   const int32_t* d_task_ids, const float* d_edge_valid, int max_tasks, int max_agents, int require_valid,
                   float* d_mask, int n_envs, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (n_envs <= 0) return n_envs == 0 ? 0 : -22;
  if (!d_records || !d_pairs || !d_n_pairs || !d_task_ids || !d_mask || max_tasks < 1 || max_agents < 1) return -22;
  if (require_valid && !d_edge_valid) return -22;
  Layout L = make_layout(*cfg);
  muav_pair_mask_kernel<<<(n_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      L, (const char*)d_records, d_pairs, d_n_pairs, d_task_ids, d_edge_valid, max_tasks, max_agents, require_valid, d_mask,
      n_envs);
  return cuda_rc(cudaGetLastError());
}

int muav_observe(const muav_config* cfg, const void* d_records, int max_rows, double* d_tasks_info, uint8_t* d_pad_mask,
                 uint8_t* d_legal_mask, double* d_agent_obs, float* d_event_flags, int32_t* d_n_rows, int n_envs,
                 void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (n_envs <= 0) return n_envs == 0 ? 0 : -22;
  if (max_rows < 1) return -22;
  Layout L = make_layout(*cfg);
  muav_observe_kernel<<<(n_envs + 63) / 64, 64, 0, (cudaStream_t)stream>>>(*cfg, L, (const char*)d_records, max_rows,
                                                                         d_tasks_info, d_pad_mask, d_legal_mask,
                                                                         d_agent_obs, d_event_flags, d_n_rows, n_envs);
  return cuda_rc(cudaGetLastError());
}

}  // extern "C"

#endif  // !MUAV_STEP_ONLY
