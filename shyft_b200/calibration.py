"""Python mirror of the calibration goal-function entry (core/model_calibration.h:404-900, api/boostpython/expose.h:472-731).

`Optimizer(model, targets, p_min, p_max)` keeps the reference's names: `calculate_goal_function(p)`, the traces, the
parameter scaling helpers.  The optimisation algorithms themselves (BOBYQA, DREAM, SCE-UA, dlib global) stay host-side
in the reference and are not part of the hot path; what the device adds is `calculate_goal_function_batch(P)`, which
evaluates a whole population of parameter vectors in one pass so that a population-based driver can call it once per
generation.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import dptr, f64

NASH_SUTCLIFFE, KLING_GUPTA, ABS_DIFF, RMSE = 0, 1, 2, 3
DISCHARGE, SNOW_COVERED_AREA, SNOW_WATER_EQUIVALENT, ROUTED_DISCHARGE, CELL_CHARGE = 0, 1, 2, 3, 4
USEC = 1000000


def calendar_period_points(start, unit, n):
    """n + 1 boundaries [s] of n calendar periods (unit: 'day', 'week', 'month', 'quarter', 'year') from `start` [s since epoch], UTC calendar:
    a calendar_dt target axis (core/time_axis.h calendar_dt; month / year steps of core/utctime_utilities.cpp:151-228) spelled as the point
    axis `TargetSpecification(time_points=...)` takes.  The day of month is clipped to the target month's length."""
    if unit in ("day", "week"):
        return [int(start) + i * (86400 if unit == "day" else 7 * 86400) for i in range(n + 1)]
    months = {"month": 1, "quarter": 3, "year": 12}[unit]
    t0 = np.datetime64(int(start), "s")
    day0 = t0.astype("datetime64[D]")
    tod = int((t0 - day0) / np.timedelta64(1, "s"))
    m0 = day0.astype("datetime64[M]")
    dom = int((day0 - m0) / np.timedelta64(1, "D"))
    out = []
    for i in range(n + 1):
        m = m0 + np.timedelta64(i * months, "M")
        length = int(((m + np.timedelta64(1, "M")).astype("datetime64[D]") - m.astype("datetime64[D]")) / np.timedelta64(1, "D"))
        d = m.astype("datetime64[D]") + np.timedelta64(min(dom, length - 1), "D")
        out.append(int(d.astype("datetime64[s]").astype(np.int64)) + tod)
    return out


class TargetSpecification:
    """target_specification<PS> (:242-329): observed series on its own fixed axis + what it is compared with."""

    def __init__(self, values, start, delta_t, catchment_indexes=(), scale_factor=1.0, calc_mode=NASH_SUTCLIFFE, s_r=1.0, s_a=1.0, s_b=1.0,
                 catchment_property=DISCHARGE, river_id=0, uid="", time_points=None):
        """`start`, `delta_t` [s]: a fixed_dt target axis; or `time_points` [s]: len(values) + 1 period boundaries of a point axis"""
        self.values = f64(values)
        self.start, self.delta_t = int(start), int(delta_t)
        self.time_points_us = None
        if time_points is not None:
            self.time_points_us = np.ascontiguousarray(np.round(np.asarray(time_points, dtype=np.float64) * USEC), dtype=np.int64)
            if self.time_points_us.size != self.values.size + 1:
                raise RuntimeError("time_points must hold one more entry than values (the end of the last period)")
        self.catchment_indexes = np.ascontiguousarray(catchment_indexes, dtype=np.int64)
        self.scale_factor, self.calc_mode, self.catchment_property = float(scale_factor), int(calc_mode), int(catchment_property)
        self.s_r, self.s_a, self.s_b, self.river_id, self.uid = float(s_r), float(s_a), float(s_b), int(river_id), uid


class Optimizer:
    def __init__(self, model, targets, p_min=None, p_max=None):
        self.model = model
        self.parameter_lower_bound = None if p_min is None else f64(p_min)
        self.parameter_upper_bound = None if p_max is None else f64(p_max)
        self.parameters_trace, self.goal_fn_trace = [], []
        self.set_target_specification(targets)

    def set_target_specification(self, targets):
        self.targets = list(targets)
        arr = (capi.Target * len(self.targets))()
        for a, t in zip(arr, self.targets):
            a.values = dptr(t.values)
            a.t0_us, a.dt_us, a.n = t.start * USEC, t.delta_t * USEC, t.values.size
            a.catchment_ids = t.catchment_indexes.ctypes.data_as(capi.c_i64p)
            a.n_catchments = t.catchment_indexes.size
            a.river_id, a.scale_factor, a.calc_mode, a.property = t.river_id, t.scale_factor, t.calc_mode, t.catchment_property
            a.s_r, a.s_a, a.s_b = t.s_r, t.s_a, t.s_b
            a.period_points_us = t.time_points_us.ctypes.data_as(capi.c_i64p) if t.time_points_us is not None else None
        self.model._ck(self.model._L.sb2_set_targets(self.model._h, C.c_int(len(self.targets)), arr))

    def calculate_goal_function(self, p):
        p = f64(p)
        g = C.c_double(0.0)
        self.model._ck(self.model._L.sb2_calculate_goal_function(self.model._h, dptr(p), C.c_int(p.size), C.byref(g)))
        self.parameters_trace.append(p.copy())
        self.goal_fn_trace.append(g.value)
        return g.value

    def calculate_goal_function_batch(self, P):
        P = f64(P).reshape(-1, self.model.parameter_size)
        goals = np.zeros(P.shape[0])
        self.model._ck(self.model._L.sb2_calculate_goal_function_batch(self.model._h, C.c_int64(P.shape[0]), dptr(P), dptr(goals)))
        self.parameters_trace.extend(P.copy())
        self.goal_fn_trace.extend(goals.tolist())
        return goals

    # reduce/expand/scale helpers of the optimizer (:435-453, 707-739): parameters with p_min == p_max are not optimised
    def active_parameters(self):
        return np.nonzero(np.abs(self.parameter_upper_bound - self.parameter_lower_bound) > 1e-6 * 0 + 0.000001)[0]

    def to_scaled(self, p):
        lo, hi = self.parameter_lower_bound, self.parameter_upper_bound
        act = self.active_parameters()
        return (f64(p)[act] - lo[act]) / (hi[act] - lo[act])

    def from_scaled(self, x, p_full):
        lo, hi = self.parameter_lower_bound, self.parameter_upper_bound
        act = self.active_parameters()
        out = f64(p_full).copy()
        out[act] = lo[act] + f64(x) * (hi[act] - lo[act])
        return out
