// Third instantiation of muav_step_kernel: the lean feature set of muav_step_lean.cu with escorts compiled IN
// (escort_enabled fixed to true; no obstacles, plain Hungarian / Coalition-Hungarian allocator) for the WPS_escort family.
#define MUAV_LEAN 1
#define MUAV_LEAN_ESCORT 1
// at most 12 environments per CTA, two CTAs per SM: 80 registers per thread, up to 24 resident environments per SM
#if !defined(MUAV_LB_THREADS)
#define MUAV_LB_THREADS 384
#define MUAV_LB_BLOCKS 2
#endif
#define MUAV_STAGE_COLD_FIXED 0   // only the hot part of the record is staged: residency first (launch_step, muav_kernels.cu)
#define MUAV_STEP_ONLY 1
#define MUAV_STEP_LAUNCHER muav_step_lean_escort_launch
#define MUAV_STEP_STATIC_SMEM muav_step_lean_escort_static_smem
#define MUAV_STEP_OCC muav_step_lean_escort_occ
#define muav muav_lean_escort
#include "muav_kernels.cu"
