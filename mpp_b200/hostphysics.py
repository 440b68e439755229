"""Scalar constitutive relations of the product (mpp_b200/csrc/physics.cuh compiled for the host into libmpp_hostcheck.so), for
problem set-up code that needs them on the CPU the way the reference's drivers call EOSWaterMod / SaturationFunction directly
(manufactured source terms of th_mms_problem.F90:1158-1458).  Not on the hot path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _L():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libmpp_hostcheck.so")
        if not os.path.exists(path):
            raise ImportError("libmpp_hostcheck.so is not built: run `make -C mpp_b200/csrc` (or __graft_entry__.build())")
        _lib = C.CDLL(path)
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class HostPhysics:
    """Same five calls as oracle.OraclePhysics."""
    VISCOSITY = 8.904156e-4                         # EOSWaterMod.F90:582 (constant)

    def density(self, P, T, itype):
        out = np.zeros(3)
        _L().hc_density(int(itype), C.c_double(P), C.c_double(T), _dp(out))
        return out[0], out[1], out[2]

    def viscosity(self, P, T):
        return self.VISCOSITY

    def internal_energy_enthalpy(self, P, T, itype, rho, drho_dT, drho_dP):
        out = np.zeros(6)
        _L().hc_internal_energy_enthalpy(int(itype), C.c_double(P), C.c_double(T), C.c_double(rho), C.c_double(drho_dT), C.c_double(drho_dP), _dp(out))
        return tuple(out)                           # U, H, dU_dT, dH_dT, dU_dP, dH_dP

    def _vg(self, P, sat_res, alpha, m):
        par = np.array([sat_res, alpha, m, 1.0 / (1.0 - m), 0.0, 0.0, 0.0, 0.0])
        out = np.zeros(4)
        _L().hc_sat(0, _dp(par), C.c_double(P), C.c_double(1.0), _dp(out))
        return out

    def vg_sat(self, P, sat_res, alpha, m):
        o = self._vg(P, sat_res, alpha, m)
        return o[0], o[1]

    def vg_relperm(self, P, sat_res, alpha, m):
        o = self._vg(P, sat_res, alpha, m)
        return o[2], o[3]
