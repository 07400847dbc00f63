// As muav_step_commit.cu (fixed WPS_commit / WPS_attn record shape, no escorts, no obstacles) but WITH the planner front
// ends and the market allocators compiled in (MUAV_LEAN_PLANNER): Urgency-Commit, Att-Commit, Urgency-Pair, Local-PI,
// CBBA-Replan launches of these scenarios.
#define MUAV_LEAN 1
#define MUAV_LEAN_PLANNER 1
#define MUAV_FIXED_SHAPE 12, 64, 64, 14, 16, 84, 0
#if !defined(MUAV_LB_THREADS)
#define MUAV_LB_THREADS 256
#define MUAV_LB_BLOCKS 2
#endif
#define MUAV_STAGE_COLD_FIXED 0   // only the hot part of the record is staged: residency first (launch_step, muav_kernels.cu)
#define MUAV_STEP_ONLY 1
#define MUAV_STEP_LAUNCHER muav_step_commit_planner_launch
#define MUAV_STEP_STATIC_SMEM muav_step_commit_planner_static_smem
#define MUAV_STEP_OCC muav_step_commit_planner_occ
#define MUAV_STEP_SHAPE muav_step_commit_planner_shape
#define muav muav_commit_planner
#include "muav_kernels.cu"
