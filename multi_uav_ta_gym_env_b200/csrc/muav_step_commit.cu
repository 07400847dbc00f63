// Fixed-shape instantiation of muav_step_kernel for the WPS_commit / WPS_attn record (12 agents, 64 task slots, 14 threats):
// the lean feature set of muav_step_lean.cu AND the record dimensions as compile-time constants (MUAV_FIXED_SHAPE,
// muav_layout.h).  Used by the launches of these scenarios that run the plain allocator (Local-Hungarian, the Att-Pair /
// Att-ContextPair hybrids); the planner front ends (Urgency-Commit, Att-Commit) need the general kernel.
#define MUAV_LEAN 1
#define MUAV_FIXED_SHAPE 12, 64, 64, 14, 16, 84, 0
#if !defined(MUAV_LB_THREADS)
#define MUAV_LB_THREADS 384
#define MUAV_LB_BLOCKS 2
#endif
#define MUAV_STAGE_COLD_FIXED 0   // only the hot part of the record is staged: residency first (launch_step, muav_kernels.cu)
#define MUAV_STEP_ONLY 1
#define MUAV_STEP_LAUNCHER muav_step_commit_launch
#define MUAV_STEP_STATIC_SMEM muav_step_commit_static_smem
#define MUAV_STEP_OCC muav_step_commit_occ
#define MUAV_STEP_SHAPE muav_step_commit_shape
#define muav muav_commit
#include "muav_kernels.cu"
