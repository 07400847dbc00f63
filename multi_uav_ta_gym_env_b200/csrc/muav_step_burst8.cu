// Fixed-shape instantiation of muav_step_kernel (WPS_burst x8 record: 64 agents, 160 task slots, 80 threats): the lean feature set AND the record dimensions as
// compile-time constants (MUAV_FIXED_SHAPE, muav_layout.h); see muav_step_hard.cu / muav_step_commit.cu.
#define MUAV_LEAN 1

#define MUAV_FIXED_SHAPE 64, 160, 160, 80, 16, 424, 0
#if !defined(MUAV_LB_THREADS)
#define MUAV_LB_THREADS 384
#define MUAV_LB_BLOCKS 2
#endif
#define MUAV_STAGE_COLD_FIXED 0
#define MUAV_STEP_ONLY 1
#define MUAV_STEP_LAUNCHER muav_step_burst8_launch
#define MUAV_STEP_STATIC_SMEM muav_step_burst8_static_smem
#define MUAV_STEP_OCC muav_step_burst8_occ
#define MUAV_STEP_SHAPE muav_step_burst8_shape
#define muav muav_burst8
#include "muav_kernels.cu"
