// Second instantiation of muav_step_kernel with the feature set of most registered scenarios fixed at compile time
// (MUAV_LEAN, see muav_core.cuh: no escorts, no obstacles, the plain Hungarian allocator in front of the step): half the
// SASS of the general kernel, which matters because the kernel is bound by instruction supply.  Every symbol of this
// translation unit lives in namespace muav_lean, so nothing collides with the general build at link time; the launcher in
// muav_kernels.cu (launch_step) picks this kernel only when the configuration really has those values.
#define MUAV_LEAN 1
// at most 12 environments per CTA, two CTAs per SM: 80 registers per thread, up to 24 resident environments per SM
#if !defined(MUAV_LB_THREADS)
#define MUAV_LB_THREADS 384
#define MUAV_LB_BLOCKS 2
#endif
#define MUAV_STAGE_COLD_FIXED 0   // only the hot part of the record is staged: residency first (launch_step, muav_kernels.cu)
#define MUAV_STEP_ONLY 1
#define MUAV_STEP_LAUNCHER muav_step_lean_launch
#define MUAV_STEP_STATIC_SMEM muav_step_lean_static_smem
#define MUAV_STEP_OCC muav_step_lean_occ
#define muav muav_lean
#include "muav_kernels.cu"
