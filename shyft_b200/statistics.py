"""Statistics readers over the resident per-cell series: the Python mirror of api/api.h:178-1600 (basic_cell_statistics and the
per-method state / response statistics) as shyft/api/pt_gs_k/__init__.py:14-27 attaches them to the models
(`model.statistics.discharge(cids)`, `model.gamma_snow_response.sca(cids)`, ...).  Every method is one call of
sb2_statistics_series / _cells / _geo: the reduction over the selected cells runs on the device (csrc/sb2_stats.cuh)."""
import ctypes as C

import numpy as np

from . import capi

CATCHMENT_IX, CELL_IX = 0, 1          # stat_scope (core/cell_model.h:178-181)
_FORCING, _RESPONSE, _STATE, _AE_POT_RATIO = 0, 1, 2, 3
_SUM, _AVERAGE, _AVERAGE_VALUE = 0, 1, 2


class _Reader:
    def __init__(self, model):
        self._m = model

    def _idx(self, indexes):
        a = np.ascontiguousarray(list(indexes), dtype=np.int64)
        return a, a.ctypes.data_as(capi.c_i64p) if a.size else None

    def _series(self, kind, series, indexes, ix_type, op, start=0, n=None):
        m = self._m
        if n is None:
            n = m.time_axis.n + (1 if kind in (_STATE, _AE_POT_RATIO) else 0) - start
        a, p = self._idx(indexes)
        out = np.zeros(n)
        m._ck(m._L.sb2_statistics_series(m._h, C.c_int(kind), C.c_int(series), p, C.c_int(a.size), C.c_int(ix_type), C.c_int(op), C.c_int64(start),
                                         C.c_int64(n), capi.dptr(out)))
        return out

    def _value(self, kind, series, indexes, i, ix_type, op):
        return float(self._series(kind, series, indexes, ix_type, _AVERAGE_VALUE if op == _AVERAGE else op, i, 1)[0])

    def _cells(self, kind, series, indexes, i, ix_type):
        m = self._m
        a, p = self._idx(indexes)
        out = np.zeros(m.size())
        n_out = C.c_int64(0)
        m._ck(m._L.sb2_statistics_cells(m._h, C.c_int(kind), C.c_int(series), p, C.c_int(a.size), C.c_int(ix_type), C.c_int64(i), capi.dptr(out),
                                        C.byref(n_out)))
        return out[: n_out.value].copy()

    def _geo(self, what, indexes, ix_type):
        m = self._m
        a, p = self._idx(indexes)
        out = C.c_double(0.0)
        m._ck(m._L.sb2_statistics_geo(m._h, C.c_int(what), p, C.c_int(a.size), C.c_int(ix_type), C.byref(out)))
        return out.value


def _feature(name, kind, series_of, op):
    """the three forms the reference gives every feature: f(indexes [, ith_timestep]) and f_value(indexes, ith_timestep)"""
    def series_or_cells(self, indexes=(), ith_timestep=None, ix_type=CATCHMENT_IX):
        s = series_of(self._m)
        if ith_timestep is None:
            return self._series(kind, s, indexes, ix_type, op)
        return self._cells(kind, s, indexes, ith_timestep, ix_type)

    def value(self, indexes, ith_timestep, ix_type=CATCHMENT_IX):
        return self._value(kind, series_of(self._m), indexes, ith_timestep, ix_type, op)
    series_or_cells.__name__, value.__name__ = name, name + "_value"
    return series_or_cells, value


def _attach(cls, name, kind, series_of, op):
    f, v = _feature(name, kind, series_of, op)
    setattr(cls, name, f)
    setattr(cls, name + "_value", v)


class BasicCellStatistics(_Reader):
    """basic_cell_statistics (api/api.h:179-420): geo sums, discharge / charge sums, area-weighted forcing averages"""
    def total_area(self, indexes=(), ix_type=CATCHMENT_IX): return self._geo(0, indexes, ix_type)
    def forest_area(self, indexes=(), ix_type=CATCHMENT_IX): return self._geo(1, indexes, ix_type)
    def glacier_area(self, indexes=(), ix_type=CATCHMENT_IX): return self._geo(2, indexes, ix_type)
    def lake_area(self, indexes=(), ix_type=CATCHMENT_IX): return self._geo(3, indexes, ix_type)
    def reservoir_area(self, indexes=(), ix_type=CATCHMENT_IX): return self._geo(4, indexes, ix_type)
    def unspecified_area(self, indexes=(), ix_type=CATCHMENT_IX): return self._geo(5, indexes, ix_type)
    def snow_storage_area(self, indexes=(), ix_type=CATCHMENT_IX): return self._geo(6, indexes, ix_type)
    def elevation(self, indexes=(), ix_type=CATCHMENT_IX): return self._geo(7, indexes, ix_type)


def _resp(name):
    from .region_model import RESPONSE_NAMES
    return lambda m: RESPONSE_NAMES.index(name)


def _state(name):
    from .region_model import STATE_SERIES_NAMES
    return lambda m: STATE_SERIES_NAMES[m.stack].index(name)


_attach(BasicCellStatistics, "discharge", _RESPONSE, _resp("avg_discharge"), _SUM)
_attach(BasicCellStatistics, "charge", _RESPONSE, _resp("charge_m3s"), _SUM)
for _i, _n in enumerate(capi.FORCING_NAMES):
    _attach(BasicCellStatistics, _n, _FORCING, (lambda k: (lambda m: k))(_i), _AVERAGE)


class KirchnerStateStatistics(_Reader):
    """kirchner_cell_state_statistics (api/api.h:422-445): sum of the instant Kirchner discharge [m3/s]"""


_attach(KirchnerStateStatistics, "discharge", _STATE, _state("kirchner_discharge"), _SUM)


class GammaSnowStateStatistics(_Reader):
    """gamma_snow_cell_state_statistics (api/api.h:447-600): area-weighted averages of the eight state fields"""


for _n in ("albedo", "lwc", "surface_heat", "alpha", "sdc_melt_mean", "acc_melt", "iso_pot_energy", "temp_swe"):
    _attach(GammaSnowStateStatistics, _n, _STATE, _state("gs_" + _n), _AVERAGE)


class GammaSnowResponseStatistics(_Reader):
    """gamma_snow_cell_response_statistics (api/api.h:602-675): sca, swe averaged; outflow, glacier_melt summed"""


_attach(GammaSnowResponseStatistics, "sca", _RESPONSE, _resp("snow_sca"), _AVERAGE)
_attach(GammaSnowResponseStatistics, "swe", _RESPONSE, _resp("snow_swe"), _AVERAGE)
_attach(GammaSnowResponseStatistics, "outflow", _RESPONSE, _resp("snow_outflow"), _SUM)
_attach(GammaSnowResponseStatistics, "glacier_melt", _RESPONSE, _resp("glacier_melt"), _SUM)


class PriestleyTaylorResponseStatistics(_Reader):
    """priestley_taylor_cell_response_statistics: area-weighted potential evapotranspiration"""


_attach(PriestleyTaylorResponseStatistics, "output", _RESPONSE, _resp("pe_output"), _AVERAGE)


class ActualEvapotranspirationResponseStatistics(_Reader):
    """actual_evapotranspiration_cell_response_statistics: area-weighted actual evapotranspiration"""


_attach(ActualEvapotranspirationResponseStatistics, "output", _RESPONSE, _resp("ae_output"), _AVERAGE)
_attach(ActualEvapotranspirationResponseStatistics, "pot_ratio", _AE_POT_RATIO, lambda m: 0, _AVERAGE)  # api/api.h:1519-1564


# ---- the HBV stacks (pt_hs_k, hbv_stack): shyft/api/pt_hs_k/__init__.py:13-19, shyft/api/hbv_stack/__init__.py:13-19 ----------
class HbvSnowStateStatistics(_Reader):
    """hbv_snow_cell_state_statistics (api/api.h:1049-1160): area-weighted swe and sca, and the per-bin series sp[i] / sw[i]
    (one area-weighted series, one per-cell vector or one value per snow bin, as the reference returns them)"""
    N_BINS = 5

    def _bins(self, which, indexes, ith_timestep, ix_type, value):
        out = []
        for i in range(self.N_BINS):
            s = _state(f"snow_{which}_{i}")(self._m)
            if value:
                out.append(self._value(_STATE, s, indexes, ith_timestep, ix_type, _AVERAGE))
            elif ith_timestep is None:
                out.append(self._series(_STATE, s, indexes, ix_type, _AVERAGE))
            else:
                out.append(self._cells(_STATE, s, indexes, ith_timestep, ix_type))
        return out

    def sp(self, indexes=(), ith_timestep=None, ix_type=CATCHMENT_IX): return self._bins("sp", indexes, ith_timestep, ix_type, False)
    def sw(self, indexes=(), ith_timestep=None, ix_type=CATCHMENT_IX): return self._bins("sw", indexes, ith_timestep, ix_type, False)
    def sp_value(self, indexes, ith_timestep, ix_type=CATCHMENT_IX): return self._bins("sp", indexes, ith_timestep, ix_type, True)
    def sw_value(self, indexes, ith_timestep, ix_type=CATCHMENT_IX): return self._bins("sw", indexes, ith_timestep, ix_type, True)


_attach(HbvSnowStateStatistics, "swe", _STATE, _state("snow_swe"), _AVERAGE)
_attach(HbvSnowStateStatistics, "sca", _STATE, _state("snow_sca"), _AVERAGE)


class HbvSnowResponseStatistics(_Reader):
    """hbv_snow_cell_response_statistics (api/api.h:1162-1205): snow outflow and glacier melt, summed [m3/s]"""


_attach(HbvSnowResponseStatistics, "outflow", _RESPONSE, _resp("snow_outflow"), _SUM)
_attach(HbvSnowResponseStatistics, "glacier_melt", _RESPONSE, _resp("glacier_melt"), _SUM)


class HbvSoilStateStatistics(_Reader):
    """hbv_soil_cell_state_statistics (api/api.h:423-443): the reference names the sum of soil_moisture `discharge`"""


_attach(HbvSoilStateStatistics, "discharge", _STATE, _state("soil_moisture"), _SUM)


class HbvTankStateStatistics(_Reader):
    """hbv_tank_cell_state_statistics (api/api.h:445-465): `discharge` = the sum of tank_uz (so marked "to be modified" in the reference)"""


_attach(HbvTankStateStatistics, "discharge", _STATE, _state("tank_uz"), _SUM)


class HbvSoilResponseStatistics(_Reader):
    """hbv_soil_cell_response_statistics (api/api.h:1471-1491): area-weighted soil outflow"""


_attach(HbvSoilResponseStatistics, "output", _RESPONSE, _resp("soil_outflow"), _AVERAGE)


class HbvActualEvapotranspirationResponseStatistics(_Reader):
    """hbv_actual_evapotranspiration_cell_response_statistics (api/api.h:1568-1590): area-weighted actual evapotranspiration"""


_attach(HbvActualEvapotranspirationResponseStatistics, "output", _RESPONSE, _resp("ae_output"), _AVERAGE)
