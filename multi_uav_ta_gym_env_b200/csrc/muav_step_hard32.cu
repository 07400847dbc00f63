// As muav_step_hard.cu with 32 task slots (BatchedMultiUAVEnv(task_cap=32): at most 24 tasks are alive at once in WPS_hard).
// feature set of muav_step_lean.cu AND the record dimensions as compile-time constants (MUAV_FIXED_SHAPE, muav_layout.h),
// so every field offset is an immediate and the loops over agents / threats / mask words have constant bounds.
#define MUAV_LEAN 1
#define MUAV_FIXED_SHAPE 8, 32, 64, 9, 16, 58, 0
// at most 8 environments per CTA, two CTAs per SM: 128 registers per thread (measured: the 80-register build that would
// allow 24 resident environments per SM is 8 % slower than this one at 16, profiles/r02_step_kernel.md)
#if !defined(MUAV_LB_THREADS)
#define MUAV_LB_THREADS 256
#define MUAV_LB_BLOCKS 2
#endif
#define MUAV_STAGE_COLD_FIXED 1   // the whole record is staged: 8-agent records are small (launch_step, muav_kernels.cu)
#define MUAV_STEP_ONLY 1
#define MUAV_STEP_LAUNCHER muav_step_hard32_launch
#define MUAV_STEP_STATIC_SMEM muav_step_hard32_static_smem
#define MUAV_STEP_OCC muav_step_hard32_occ
#define MUAV_STEP_SHAPE muav_step_hard32_shape
#define muav muav_hard32
#include "muav_kernels.cu"
