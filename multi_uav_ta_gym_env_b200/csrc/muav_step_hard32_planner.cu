// Fixed-shape instantiation of muav_step_kernel (WPS_hard record with 32 task slots, planner front ends and market allocators compiled in: Urgency-Pair, Local-PI, CBBA-Replan): the lean feature set AND the record dimensions as
// compile-time constants (MUAV_FIXED_SHAPE, muav_layout.h); see muav_step_hard.cu / muav_step_commit.cu.
#define MUAV_LEAN 1
#define MUAV_LEAN_PLANNER 1
#define MUAV_FIXED_SHAPE 8, 32, 64, 9, 16, 58, 0
#if !defined(MUAV_LB_THREADS)
#define MUAV_LB_THREADS 256
#define MUAV_LB_BLOCKS 2
#endif
#define MUAV_STAGE_COLD_FIXED 1
#define MUAV_STEP_ONLY 1
#define MUAV_STEP_LAUNCHER muav_step_hard32_planner_launch
#define MUAV_STEP_STATIC_SMEM muav_step_hard32_planner_static_smem
#define MUAV_STEP_OCC muav_step_hard32_planner_occ
#define MUAV_STEP_SHAPE muav_step_hard32_planner_shape
#define muav muav_hard32_planner
#include "muav_kernels.cu"
