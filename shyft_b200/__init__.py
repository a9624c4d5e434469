"""shyft_b200 -- B200-native (sm_100a) implementation of Shyft's per-cell region-model time-stepping hot path.

Only what the path needs: `csrc/` (CUDA kernels + the C ABI of include/shyft_b200.h), the ctypes binding (`capi`), the
Python mirror of the reference's region-model surface (`region_model`), the calibration goal-function entry
(`calibration`) and the synthetic regions used by tests and bench (`synthetic`).  No CPU fallback.
"""
from .capi import (COLLECT_ALL, COLLECT_DISCHARGE, COLLECT_NONE, COLLECT_SNOW, COLLECT_STATE, HBV_STACK, PT_GS_K, PT_HPS_K, PT_HS_K, PT_SS_K,  # noqa: F401
                   InterpolationParameter)
from .region_model import (GeoPointSources, GeoPointSourceVector, HbvModel, HbvOptModel, HbvStackModel, HbvStackOptModel, PTGSKModel, PTGSKOptModel, PTHSKModel, PTHSKOptModel, PTSSKModel, PTSSKOptModel, PTHPSKModel, PTHPSKOptModel,  # noqa: F401
                           RegionEnvironment, RegionModel, TimeAxis, geo_cell_data_vector)
from .calibration import Optimizer, TargetSpecification, calendar_period_points  # noqa: F401,E402
from .state_io import StateIoHandler, StateWithIdVector, cell_state_id_of  # noqa: F401,E402
