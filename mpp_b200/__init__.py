"""mpp_b200 -- B200-native (sm_100a, fp64) column-physics solver behind MPP's system-of-equations API.

The product is the CUDA library libmppgpu.so (mpp_b200/csrc, C ABI in include/mppgpu.h); this package is the
thin Python mirror of the reference's sysofeqns interface used by the tests and the benchmark.
"""
from . import constants  # noqa: F401
from .soe import VSFM, Thermal, ThermalSnow, TH, MPPError, host_register, host_unregister, comm_unique_id  # noqa: F401

__all__ = ["constants", "VSFM", "Thermal", "ThermalSnow", "TH", "MPPError", "host_register", "host_unregister", "comm_unique_id"]
