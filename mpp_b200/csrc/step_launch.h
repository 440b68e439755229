// step_launch.h -- launch entry points of the step kernels that are compiled in their own translation units
// (vsfm_step2_inst.cu once per (lanes per column, saturation function), th_step2_inst.cu once per model combination) so that the
// 18 + 4 template instances build in parallel instead of serially inside mppgpu.cu.
#pragma once
#include <cuda_runtime.h>

namespace mpp {
struct VsfmArgs;
struct THArgs;

// variant: 0 plain (no boundary conditions), 1 boundary conditions / down-regulated sink / IFC-67 density, 2 the RETRY specialisation,
// 3 the residual / Jacobian probe (mppgpu_eval)
#define MPP_DECL_VSFM2(LPC, SF) void vsfm2_launch_##LPC##_##SF(const VsfmArgs &A, int variant, int nblocks, cudaStream_t s);
MPP_DECL_VSFM2(8, 0) MPP_DECL_VSFM2(8, 1) MPP_DECL_VSFM2(8, 2) MPP_DECL_VSFM2(16, 0) MPP_DECL_VSFM2(16, 1) MPP_DECL_VSFM2(16, 2)
#undef MPP_DECL_VSFM2

// combo: 0 VG + Tanaka + constant c_p, 1 VG + IFC-67 + IFC-67, 2 smoothed Brooks-Corey + Tanaka + constant c_p, 3 run-time dispatch;
// 0p / 2p (INST_COMBO 4 / 5): combos 0 / 2 with the boundary connection on the padding lane fixed at compile time
void th2_launch_0(const THArgs &A, int nblocks, cudaStream_t s);
void th2_launch_1(const THArgs &A, int nblocks, cudaStream_t s);
void th2_launch_2(const THArgs &A, int nblocks, cudaStream_t s);
void th2_launch_3(const THArgs &A, int nblocks, cudaStream_t s);
void th2_launch_0p(const THArgs &A, int nblocks, cudaStream_t s);
void th2_launch_2p(const THArgs &A, int nblocks, cudaStream_t s);
}  // namespace mpp
