"""Loader of the in-tree CUDA library libmppgpu.so (C ABI in include/mppgpu.h).

There is no CPU fallback: if the library is missing the import of any solver class fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPPGPU_LIB_PATH", os.path.join(_HERE, "libmppgpu.so"))   # override: kernel-variant experiments only
_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)

class Xfer(C.Structure):
    """mppgpu_xfer (include/mppgpu.h)."""
    _fields_ = [("ieqn", C.c_int), ("auxvar_type", C.c_int), ("var_type", C.c_int), ("cond_id", C.c_int), ("host", C.POINTER(C.c_double))]


class ElmColumns(C.Structure):
    """mppgpu_elm_columns (include/mppgpu.h)."""
    _fields_ = [("fortran_order", C.c_int), ("npft", C.c_int), ("max_patch_per_col", C.c_int),
                ("col_pfti", c_ip), ("col_npfts", c_ip), ("pft_active", c_ip), ("pft_wtcol", c_dp), ("rootr_pft", c_dp), ("qflx_tran_veg_pft", c_dp),
                ("rootr_col", c_dp),
                ("qflx_tran_veg_col", c_dp), ("qflx_infl", c_dp), ("qflx_dew_snow", c_dp), ("qflx_dew_grnd", c_dp), ("qflx_sub_snow", c_dp), ("frac_h2osfc", c_dp),
                ("snl", c_ip),
                ("qflx_drain", c_dp), ("zwt", c_dp), ("h2osoi_liq", c_dp), ("h2osoi_ice", c_dp), ("mflx_snowlyr_col", c_dp),
                ("mflx_neg_snow_col", c_dp), ("mflx_drain_perched", c_dp),
                ("smp_l", c_dp), ("soilp_col", c_dp), ("qcharge", c_dp), ("abs_mass_error", c_dp), ("iter_count", c_ip), ("status", c_ip)]


class ElmThermalColumns(C.Structure):
    """mppgpu_elm_thermal_columns (include/mppgpu.h)."""
    _fields_ = [("snl", c_ip), ("z", c_dp), ("dz", c_dp), ("zi", c_dp), ("t_soisno", c_dp), ("h2osoi_liq", c_dp), ("h2osoi_ice", c_dp),
                ("frac_sno_eff", c_dp), ("h2osno", c_dp), ("h2osfc", c_dp), ("frac_h2osfc", c_dp), ("t_h2osfc", c_dp), ("sabg_lyr", c_dp),
                ("dhsdT", c_dp), ("hs_soil", c_dp), ("hs_top_snow", c_dp), ("hs_h2osfc", c_dp), ("tvector", c_dp)]


_SIGS = {
    "mppgpu_last_error": (C.c_char_p, []),
    "mppgpu_version": (C.c_int, []),
    "mppgpu_device_count": (C.c_int, []),
    "mppgpu_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mppgpu_destroy": (C.c_int, [C.c_void_p]),
    "mppgpu_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mppgpu_synchronize": (C.c_int, [C.c_void_p]),
    "mppgpu_set_mesh": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp, c_ip]),
    "mppgpu_set_connection_distances": (C.c_int, [C.c_void_p, c_dp, c_dp]),
    "mppgpu_add_condition": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_ip]),
    "mppgpu_vsfm_set_soils": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_int, C.c_int]),
    "mppgpu_thermal_set_soils": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_ip, C.c_int, C.c_int]),
    "mppgpu_thermal_set_cnfac": (C.c_int, [C.c_void_p, C.c_double]),
    "mppgpu_thermal_add_snow_ssw": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "mppgpu_thermal_elm_solve": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.POINTER(ElmThermalColumns), C.c_double]),
    "mppgpu_elm_set_pipeline": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mppgpu_vsfm_elm_set_geometry": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_double, c_ip]),
    "mppgpu_vsfm_elm_set_geometry_f": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_double, c_ip]),
    "mppgpu_vsfm_elm_solve": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.POINTER(ElmColumns), c_ip, c_ip]),
    "mppgpu_th_set_soils": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_int, C.c_int, C.c_int]),
    "mppgpu_th_set_energy_permeability": (C.c_int, [C.c_void_p, c_dp, C.c_int]),
    "mppgpu_set_tolerances": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]),
    "mppgpu_thermal_set_bulk_copy": (C.c_int, [C.c_void_p, C.c_int]),
    "mppgpu_set_step_budget": (C.c_int, [C.c_void_p, C.c_int]),
    "mppgpu_set_column_ordering": (C.c_int, [C.c_void_p, C.c_int]),
    "mppgpu_restart": (C.c_int, [C.c_void_p, c_dp, C.c_int]),
    "mppgpu_set_data": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, C.c_int]),
    "mppgpu_set_idata": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, C.c_int]),
    "mppgpu_get_data": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, C.c_int]),
    "mppgpu_set_data_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "mppgpu_get_data_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "mppgpu_pre_step_dt": (C.c_int, [C.c_void_p]),
    "mppgpu_step_dt": (C.c_int, [C.c_void_p, C.c_double, C.c_int, c_ip, c_ip]),
    "mppgpu_step_dt_async": (C.c_int, [C.c_void_p, C.c_double, C.c_int]),
    "mppgpu_step_result": (C.c_int, [C.c_void_p, c_ip, c_ip]),
    "mppgpu_post_step_dt": (C.c_int, [C.c_void_p]),
    "mppgpu_vsfm_coupled_step": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.c_int, C.POINTER(Xfer), C.c_int, C.POINTER(Xfer), C.c_int, c_ip, c_ip]),
    "mppgpu_get_column_stats": (C.c_int, [C.c_void_p, c_ip, c_ip, c_ip, c_ip]),
    "mppgpu_vsfm_mass_balance": (C.c_int, [C.c_void_p, C.c_double, c_dp, c_dp]),
    "mppgpu_reduction_buffer_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mppgpu_comm_unique_id": (C.c_int, [C.c_void_p]),
    "mppgpu_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mppgpu_global_reduce_async": (C.c_int, [C.c_void_p]),
    "mppgpu_global_mass_balance": (C.c_int, [C.c_void_p, c_dp, c_dp, c_ip]),
    "mppgpu_host_register": (C.c_int, [C.c_void_p, C.c_longlong]),
    "mppgpu_host_unregister": (C.c_int, [C.c_void_p]),
    "mppgpu_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong)]),
    "mppgpu_last_step_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "mppgpu_eval": (C.c_int, [C.c_void_p, C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
}

EXPORTS = tuple(_SIGS)


def lib():
    """Return the loaded CUDA library; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "mpp_b200: %s is missing -- build it with `make -C mpp_b200/csrc` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class MPPError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise MPPError(lib().mppgpu_last_error().decode("utf-8", "replace"))
