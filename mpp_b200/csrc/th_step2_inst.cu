// th_step2_inst.cu -- one model combination (INST_COMBO) of the th_step2_kernel template instances (see step_launch.h).
#define MPP_STEP_KERNEL_TU
#include "step_launch.h"
#include "physics.cuh"
#include "vsfm_kernels.cuh"
#include "th_kernels.cuh"
#include "th_kernels2.cuh"

namespace mpp {
#if INST_COMBO == 0
void th2_launch_0(const THArgs &A, int nblocks, cudaStream_t s) { th_step2_kernel<16, SATFUNC_VG, DENSITY_TGDPB01, INT_ENERGY_ENTHALPY_CONSTANT><<<nblocks, TH2_THREADS, 0, s>>>(A); }
#elif INST_COMBO == 1
void th2_launch_1(const THArgs &A, int nblocks, cudaStream_t s) { th_step2_kernel<16, SATFUNC_VG, DENSITY_IFC67, INT_ENERGY_ENTHALPY_IFC67><<<nblocks, TH2_THREADS, 0, s>>>(A); }
#elif INST_COMBO == 2
void th2_launch_2(const THArgs &A, int nblocks, cudaStream_t s) { th_step2_kernel<16, SATFUNC_SBC, DENSITY_TGDPB01, INT_ENERGY_ENTHALPY_CONSTANT><<<nblocks, TH2_THREADS, 0, s>>>(A); }
#elif INST_COMBO == 4
void th2_launch_0p(const THArgs &A, int nblocks, cudaStream_t s) { th_step2_kernel<16, SATFUNC_VG, DENSITY_TGDPB01, INT_ENERGY_ENTHALPY_CONSTANT, true><<<nblocks, TH2_THREADS, 0, s>>>(A); }
#elif INST_COMBO == 5
void th2_launch_2p(const THArgs &A, int nblocks, cudaStream_t s) { th_step2_kernel<16, SATFUNC_SBC, DENSITY_TGDPB01, INT_ENERGY_ENTHALPY_CONSTANT, true><<<nblocks, TH2_THREADS, 0, s>>>(A); }
#else
void th2_launch_3(const THArgs &A, int nblocks, cudaStream_t s) { th_step2_kernel<16, -1, -1, -1><<<nblocks, TH2_THREADS, 0, s>>>(A); }
#endif
}  // namespace mpp
