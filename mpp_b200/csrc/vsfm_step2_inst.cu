// vsfm_step2_inst.cu -- one (INST_LPC, INST_SF) slice of the vsfm_step2_kernel template instances (see step_launch.h).
#define MPP_STEP_KERNEL_TU
#include "step_launch.h"
#include "physics.cuh"
#include "vsfm_kernels.cuh"
#include "vsfm_kernels2.cuh"

namespace mpp {
#define MPP_NAME2(l, s) vsfm2_launch_##l##_##s
#define MPP_NAME(l, s) MPP_NAME2(l, s)
void MPP_NAME(INST_LPC, INST_SF)(const VsfmArgs &A, int variant, int nblocks, cudaStream_t s)
{
  if (variant == 3)      vsfm_step2_kernel<INST_LPC, INST_SF, true, false, true><<<nblocks, VSFM2_THREADS, 0, s>>>(A);
  else if (variant == 2) vsfm_step2_kernel<INST_LPC, INST_SF, true, true><<<nblocks, VSFM2_THREADS, 0, s>>>(A);
  else if (variant == 1) vsfm_step2_kernel<INST_LPC, INST_SF, true><<<nblocks, VSFM2_THREADS, 0, s>>>(A);
  else                   vsfm_step2_kernel<INST_LPC, INST_SF, false><<<nblocks, VSFM2_THREADS, 0, s>>>(A);
}
}  // namespace mpp
