"""Deterministic synthetic regions of the shapes BASELINE.json names: cells on a 1 km grid, jittered
stations, sinusoid-plus-noise forcing series.  numpy PCG64, seed 20260101 + config index.  Used by tests/ and bench.py.
"""
import math

import numpy as np

from .region_model import RegionEnvironment, TimeAxis, geo_cell_data_vector

T0_2014_09_01 = 1409529600  # 2014-09-01T00:00:00Z


def make_cells(n_cells, rng, cells_per_catchment=1000, with_routing=False):
    nx = int(math.ceil(math.sqrt(n_cells)))
    i = np.arange(n_cells)
    x = 500.0 + 1000.0 * (i % nx)
    y = 500.0 + 1000.0 * (i // nx)
    L = 1000.0 * nx
    z = 400.0 + 600.0 * (np.sin(x / L) + np.sin(y / L)) / 2 + 200.0 * rng.random(n_cells)
    cid = 1 + i // cells_per_catchment
    rid = cid if with_routing else 0
    dist = 1000.0 * (1 + (i % cells_per_catchment) / 40.0) if with_routing else 0.0
    return geo_cell_data_vector(x, y, z, area=1.0e6, catchment_id=cid, radiation_slope_factor=0.9, glacier=0.01, lake=0.05, reservoir=0.19,
                                forest=0.30, routing_id=rid, routing_distance=dist), L


def surface_z(x, y, L):
    return 400.0 + 600.0 * (np.sin(x / L) + np.sin(y / L)) / 2


def make_stations(n_stations, L, rng):
    k = int(math.ceil(math.sqrt(n_stations)))
    spacing = L / k
    j = np.arange(n_stations)
    x = (0.5 + (j % k)) * spacing + rng.uniform(-0.35, 0.35, n_stations) * spacing
    y = (0.5 + (j // k)) * spacing + rng.uniform(-0.35, 0.35, n_stations) * spacing
    z = surface_z(x, y, L) + 200.0 * rng.random(n_stations)
    return np.stack([x, y, z], axis=1)


def make_station_series(xyz, time_axis, rng, nan_fraction=0.0):
    """-> dict name -> [T][S] on the model axis (POINT_AVERAGE_VALUE)."""
    T, S = time_axis.n, xyz.shape[0]
    t = time_axis.start + time_axis.delta_t * np.arange(T, dtype=np.int64)
    days = t / 86400.0
    # day of year is only used for the seasonal shape of the synthetic signal
    doy = (days - (np.datetime64("2014-01-01") - np.datetime64("1970-01-01")).astype(int)) % 365.25
    h = (t % 86400) / 3600.0
    z = xyz[:, 2][None, :]
    temp = 3 + 11 * np.sin(2 * np.pi * (doy - 110) / 365)[:, None] + 3 * np.sin(2 * np.pi * (h - 9) / 24)[:, None] - 0.006 * z \
        + rng.normal(0, 1.5, (T, S))
    event = rng.random(T) < 0.12
    prec = np.where(event[:, None], rng.exponential(1.8, (T, S)), 0.0)
    rad = np.maximum(0.0, (120 + 180 * np.sin(2 * np.pi * (doy - 80) / 365)) * np.sin(np.pi * (h - 6) / 12))[:, None] * rng.uniform(0.5, 1.0, (T, S))
    wind = np.abs(rng.normal(3, 2, (T, S)))
    rh = np.clip(0.75 + 0.15 * rng.normal(0, 1, (T, S)), 0.3, 1.0)
    out = dict(temperature=temp, precipitation=prec, radiation=rad, wind_speed=wind, rel_hum=rh)
    if nan_fraction > 0:
        for v in out.values():
            v[rng.random((T, S)) < nan_fraction] = np.nan
    return out


def make_region(n_cells, n_steps, n_stations, config_index=0, dt=3600, cells_per_catchment=1000, with_routing=False, nan_fraction=0.0,
                start=T0_2014_09_01):
    """-> (geo record array, TimeAxis, RegionEnvironment)"""
    rng = np.random.Generator(np.random.PCG64(20260101 + config_index))
    geo, L = make_cells(n_cells, rng, cells_per_catchment, with_routing)
    ta = TimeAxis(start, dt, n_steps)
    xyz = make_stations(n_stations, L, rng)
    series = make_station_series(xyz, ta, rng, nan_fraction)
    env = RegionEnvironment(**{k: (xyz, v) for k, v in series.items()})
    return geo, ta, env


def default_state(stack, n_cells, q0=0.8):
    """state_t{} defaults with kirchner.q = q0 mm/h."""
    if stack == 0:
        s = np.tile(np.array([0.4, 0.1, 30000.0, 1.26, 0.0, 0.0, 0.0, 0.0, q0]), (n_cells, 1))
    elif stack == 3:
        s = np.tile(np.array([4.077, 40.77, 0.0, 0.0, 0.0, 0.0, 0.0, q0]), (n_cells, 1))   # skaugen::state() + kirchner.q
    elif stack == 4:   # hbv_physical_snow::state() after distribute() (core/hbv_physical_snow.h:132-189) + kirchner.q
        s = np.tile(np.array([0.0] * 10 + [0.4] * 5 + [0.0] * 5 + [30000.0, 0.0, 0.0, q0]), (n_cells, 1))
    elif stack == 1:
        s = np.zeros((n_cells, 13))
        s[:, 12] = q0
    else:
        s = np.zeros((n_cells, 15))
        s[:, 12:15] = (50.0, 20.0, 10.0)
    return s


def river_chain(n_catchments, depth=8):
    """One river per catchment (id = catchment id); chains of `depth` rivers draining into each other."""
    rivers = []
    for k in range(1, n_catchments + 1):
        down = k + 1 if (k % depth) != 0 and k < n_catchments else 0
        rivers.append([k, down, 3600.0 * (1 + k % 4), 1.0, 7.0, 0.0])
    return np.array(rivers, dtype=np.float64)
