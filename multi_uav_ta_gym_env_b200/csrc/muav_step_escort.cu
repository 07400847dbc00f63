// Fixed-shape instantiation of muav_step_kernel for the WPS_escort record (14 agents, 80 task slots, 512 ids, 10 threats):
// the feature set of muav_step_lean_escort.cu (escorts compiled in, no obstacles, plain Hungarian / Coalition-Hungarian
// allocator) AND the record dimensions as compile-time constants (MUAV_FIXED_SHAPE, muav_layout.h).
#define MUAV_LEAN 1
#define MUAV_LEAN_ESCORT 1
#define MUAV_FIXED_SHAPE 14, 80, 512, 10, 16, 84, 0
#if !defined(MUAV_LB_THREADS)
#define MUAV_LB_THREADS 384
#define MUAV_LB_BLOCKS 2
#endif
#define MUAV_STAGE_COLD_FIXED 0   // only the hot part of the record is staged: residency first (launch_step, muav_kernels.cu)
#define MUAV_STEP_ONLY 1
#define MUAV_STEP_LAUNCHER muav_step_escort_launch
#define MUAV_STEP_STATIC_SMEM muav_step_escort_static_smem
#define MUAV_STEP_OCC muav_step_escort_occ
#define MUAV_STEP_SHAPE muav_step_escort_shape
#define muav muav_escort
#include "muav_kernels.cu"
