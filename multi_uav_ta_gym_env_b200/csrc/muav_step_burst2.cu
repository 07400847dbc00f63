// Fixed-shape instantiation of muav_step_kernel (WPS_burst x2 record: 16 agents, 64 task slots, 20 threats): the lean feature set AND the record dimensions as
// compile-time constants (MUAV_FIXED_SHAPE, muav_layout.h); see muav_step_hard.cu / muav_step_commit.cu.
#define MUAV_LEAN 1

#define MUAV_FIXED_SHAPE 16, 64, 64, 20, 16, 112, 0
#if !defined(MUAV_LB_THREADS)
#define MUAV_LB_THREADS 384
#define MUAV_LB_BLOCKS 2
#endif
#define MUAV_STAGE_COLD_FIXED 0
#define MUAV_STEP_ONLY 1
#define MUAV_STEP_LAUNCHER muav_step_burst2_launch
#define MUAV_STEP_STATIC_SMEM muav_step_burst2_static_smem
#define MUAV_STEP_OCC muav_step_burst2_occ
#define MUAV_STEP_SHAPE muav_step_burst2_shape
#define muav muav_burst2
#include "muav_kernels.cu"
