! mppgpu_binding.F90 -- ISO_C_BINDING interfaces of libmppgpu.so (C prototypes: include/mppgpu.h) for MPP-LSM/MPP.
! Drop into src/mpp/soe/ and route the type-bound procedures of sysofeqns_base_type through it as INTEGRATION.md shows.
! Written against the reference's Fortran 2003 sources; no Fortran compiler exists in the build image, so this file has
! been syntax-reviewed only.
module MPPGpuBinding
  use, intrinsic :: iso_c_binding
  implicit none
  private

  interface
     integer(c_int) function mppgpu_create(soe_itype, ncol, nlev, device, handle) bind(C, name="mppgpu_create")
       import :: c_int, c_ptr
       integer(c_int), value :: soe_itype, ncol, nlev, device
       type(c_ptr)           :: handle
     end function
     integer(c_int) function mppgpu_set_mesh(h, orientation, dz, area, col_active) bind(C, name="mppgpu_set_mesh")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value          :: h
       integer(c_int), value       :: orientation
       real(c_double), intent(in)  :: dz(*), area(*)
       type(c_ptr), value          :: col_active          ! c_null_ptr = all columns active
     end function
     integer(c_int) function mppgpu_add_condition(h, ieqn, ss_or_bc, cond_type, region, cond_id) bind(C, name="mppgpu_add_condition")
       import :: c_int, c_ptr
       type(c_ptr), value    :: h
       integer(c_int), value :: ieqn, ss_or_bc, cond_type, region
       integer(c_int)        :: cond_id
     end function
     integer(c_int) function mppgpu_vsfm_set_soils(h, watsat, hksat, bsw, sucsat, residual_sat, satfunc_type, density_type) &
          bind(C, name="mppgpu_vsfm_set_soils")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value         :: h
       real(c_double), intent(in) :: watsat(*), hksat(*), bsw(*), sucsat(*), residual_sat(*)
       integer(c_int), value      :: satfunc_type, density_type
     end function
     integer(c_int) function mppgpu_thermal_add_snow_ssw(h, nlevsno, soil_top_dist_dn) bind(C, name="mppgpu_thermal_add_snow_ssw")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value         :: h
       integer(c_int), value      :: nlevsno
       real(c_double), intent(in) :: soil_top_dist_dn(*)
     end function
     integer(c_int) function mppgpu_set_tolerances(h, atol, rtol, stol, max_it, max_funcs) bind(C, name="mppgpu_set_tolerances")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: h
       real(c_double), value :: atol, rtol, stol
       integer(c_int), value :: max_it, max_funcs
     end function
     integer(c_int) function mppgpu_restart(h, x, n) bind(C, name="mppgpu_restart")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value         :: h
       real(c_double), intent(in) :: x(*)
       integer(c_int), value      :: n
     end function
     integer(c_int) function mppgpu_set_data(h, ieqn, auxvar_type, var_type, cond_id, data, n) bind(C, name="mppgpu_set_data")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value         :: h
       integer(c_int), value      :: ieqn, auxvar_type, var_type, cond_id, n
       real(c_double), intent(in) :: data(*)
     end function
     integer(c_int) function mppgpu_get_data(h, ieqn, auxvar_type, var_type, cond_id, data, n) bind(C, name="mppgpu_get_data")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: h
       integer(c_int), value :: ieqn, auxvar_type, var_type, cond_id, n
       real(c_double)        :: data(*)
     end function
     integer(c_int) function mppgpu_pre_step_dt(h) bind(C, name="mppgpu_pre_step_dt")
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function
     integer(c_int) function mppgpu_step_dt(h, dt, nstep, converged, converged_reason) bind(C, name="mppgpu_step_dt")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: h
       real(c_double), value :: dt
       integer(c_int), value :: nstep
       integer(c_int)        :: converged, converged_reason
     end function
     integer(c_int) function mppgpu_post_step_dt(h) bind(C, name="mppgpu_post_step_dt")
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function
     integer(c_int) function mppgpu_destroy(h) bind(C, name="mppgpu_destroy")
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function
     integer(c_int) function mppgpu_host_register(ptr, nbytes) bind(C, name="mppgpu_host_register")
       import :: c_int, c_ptr, c_long_long
       type(c_ptr), value          :: ptr
       integer(c_long_long), value :: nbytes
     end function
     integer(c_int) function mppgpu_host_unregister(ptr) bind(C, name="mppgpu_host_unregister")
       import :: c_int, c_ptr
       type(c_ptr), value :: ptr
     end function
     type(c_ptr) function mppgpu_last_error() bind(C, name="mppgpu_last_error")
       import :: c_ptr
     end function
  end interface

  public :: mppgpu_create, mppgpu_set_mesh, mppgpu_add_condition, mppgpu_vsfm_set_soils, mppgpu_thermal_add_snow_ssw, mppgpu_set_tolerances, &
            mppgpu_restart, mppgpu_set_data, mppgpu_get_data, mppgpu_pre_step_dt, mppgpu_step_dt, mppgpu_post_step_dt, &
            mppgpu_destroy, mppgpu_last_error, mppgpu_host_register, mppgpu_host_unregister
end module MPPGpuBinding
