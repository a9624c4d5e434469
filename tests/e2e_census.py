"""End-to-end parity census.

The shipped path interpolates on the FP64 tensor cores (summation order differs from the oracle's, ~1e-13 relative on the forcing) and
feeds that forcing into stacks whose results are defined by DISCRETE decisions -- precipitation phase `T < tx`, the sign tests of the
energy balance, Brent's 12-bit search in `corr_lwc`, the accept / reject sequence of the Runge-Kutta controller.  A last-bit forcing
difference either stays a last-bit difference (then the 1e-9 contract holds) or flips one of those decisions (then the two runs take
different, equally valid paths until the snow pack melts out).  This module counts both:

  * per series: the fraction of (step, cell) values within `rtol` relative (plus an absolute floor of `atol_frac` x the series' largest
    magnitude, as tests/parity.py);
  * per cell: the first step at which any compared series leaves the tolerance, attributed to the decision that flipped there.

It compares two result sets window by window (the device keeps one window of per-cell series); the GPU tests feed device windows against
oracle results, the CPU test feeds two oracle runs whose forcing differs by a synthetic 1e-13 perturbation (which exercises the census
itself, and measures how sensitive the REFERENCE ALGORITHM is to last-bit noise, without a GPU).
"""
import numpy as np

RTOL = 1.0e-9
ATOL_FRAC = 1.0e-13


def outside(got, want, scale, rtol=RTOL, atol_frac=ATOL_FRAC):
    """bool array: |got - want| > rtol * max(|got|, |want|) + atol_frac * scale (NaN on one side only counts as outside)."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    with np.errstate(invalid="ignore"):
        bad = np.abs(got - want) > rtol * np.maximum(np.abs(got), np.abs(want)) + atol_frac * scale
    bad &= ~(nan_g & nan_w)
    bad |= nan_g ^ nan_w
    return bad


class Census:
    """want: dict name -> [T][n] response series and [T+1][n] state series of the checker; forcing_want: dict name -> [T][n].
    `watch`: the state series whose movement attributes a first divergence (see `attribute`)."""

    def __init__(self, want, forcing_want, tx, rtol=RTOL, watch=("gs_lwc", "gs_alpha", "gs_sdc_melt_mean")):
        self.want, self.fw, self.tx, self.rtol = want, forcing_want, tx, rtol
        self.T, self.n = forcing_want["temperature"].shape
        self.names = [k for k, v in want.items() if np.ndim(v) == 2 and v.shape[1] == self.n and v.shape[0] in (self.T, self.T + 1)]
        self.scale = {k: float(np.nanmax(np.abs(want[k]))) if np.any(~np.isnan(want[k])) else 0.0 for k in self.names}
        self.bad = {k: np.zeros((self.T, self.n), dtype=bool) for k in self.names}   # row i: the response of step i / the state at its END
        self.worst = {k: 0.0 for k in self.names}
        self.watch = [w for w in watch if w in self.names]
        self.got_watch = {w: np.full_like(want[w], np.nan) for w in self.watch}
        self.t_got = np.full((self.T, self.n), np.nan)
        self.f_worst = {k: 0.0 for k in forcing_want}
        self.f_outside = {k: 0 for k in forcing_want}
        self.f_scale = {k: float(np.nanmax(np.abs(v))) for k, v in forcing_want.items()}
        self.rows_seen = 0

    def add_window(self, w0, got, forcing_got):
        """got: dict name -> rows [w0, w0+wn) of a response series, or points [w0, w0+wn] (wn+1 rows) of a state series."""
        wn = forcing_got["temperature"].shape[0]
        for k in self.names:
            if k not in got:
                continue
            g = np.asarray(got[k])
            if self.want[k].shape[0] == self.T + 1:
                g, w = g[1:wn + 1], self.want[k][w0 + 1:w0 + wn + 1]   # the point at the end of each step
            else:
                g, w = g[:wn], self.want[k][w0:w0 + wn]
            b = outside(g, w, self.scale[k], self.rtol)
            self.bad[k][w0:w0 + wn] = b
            if b.any():
                with np.errstate(invalid="ignore", divide="ignore"):
                    self.worst[k] = max(self.worst[k], float(np.nanmax(np.where(b, np.abs(g - w) / np.maximum(np.abs(w), 1e-300), 0.0))))
            if k in self.got_watch:
                self.got_watch[k][w0 + 1:w0 + wn + 1] = g
        self.t_got[w0:w0 + wn] = forcing_got["temperature"]
        for k, v in forcing_got.items():
            w = self.fw[k][w0:w0 + wn]
            self.f_outside[k] += int(outside(v, w, self.f_scale[k], 1e-11, 1e-14).sum())
            with np.errstate(invalid="ignore", divide="ignore"):
                self.f_worst[k] = max(self.f_worst[k], float(np.nanmax(np.abs(v - w) / np.maximum(np.abs(w), 1e-300) * (np.abs(w) > 1e-9 * self.f_scale[k]))))
        self.rows_seen += wn

    def _moved(self, name, row, c, rel=1.0e-12):
        if name not in self.got_watch:
            return False
        a, b = self.got_watch[name][row, c], self.want[name][row, c]
        return abs(a - b) > rel * max(abs(a), abs(b))

    def attribute(self, i, c, bad_names):
        """the decision behind the first divergence of cell c at step i"""
        a, b = self.t_got[i, c], self.fw["temperature"][i, c]
        if (a < self.tx) != (b < self.tx):
            return "precipitation_phase_T_lt_tx"
        if (a < 0.0) != (b < 0.0) or (a > 0.0) != (b > 0.0):
            return "sign_of_T"
        if set(bad_names) <= {"pe_output", "ae_output"}:
            # Priestley-Taylor's net radiation is a difference of two large terms clipped at zero: next to zero the relative error of
            # pe (and of ae = pe x ...) is unbounded.  One step, no state involved, nothing persists.
            return "pe_cancellation_next_to_zero"
        if self._moved("gs_lwc", i + 1, c) and not self._moved("gs_alpha", i + 1, c) and not self._moved("gs_sdc_melt_mean", i + 1, c):
            return "brent_corr_lwc"   # liquid water content is the only snow state of the step that moved: the search result
        if any(k.startswith(("gs_", "snow_")) for k in bad_names):
            return "snow_threshold_other"
        if "kirchner_discharge" in bad_names or "avg_discharge" in bad_names:
            return "rk_accept_reject"
        return "unattributed"

    def result(self):
        assert self.rows_seen == self.T, "census: not every window was added"
        out = {"cells": int(self.n), "steps": int(self.T), "rtol": self.rtol, "series": {}}
        bad_any = np.zeros((self.T, self.n), dtype=bool)
        for k in self.names:
            b = self.bad[k]
            out["series"][k] = {"within": float(1.0 - b.mean()), "outside": int(b.sum()), "worst_rel": self.worst[k]}
            bad_any |= b
        out["cell_steps_within"] = float(1.0 - bad_any.mean())
        flipped = bad_any.any(axis=0)
        out["cells_with_a_divergence"] = int(flipped.sum())
        first = np.where(flipped, bad_any.argmax(axis=0), -1)
        causes, examples = {}, []
        persistent = np.zeros(self.n, dtype=bool)
        for c in np.nonzero(flipped)[0]:
            i = int(first[c])
            names = [k for k in self.names if self.bad[k][i, c]]
            cause = self.attribute(i, c, names)
            causes[cause] = causes.get(cause, 0) + 1
            persistent[c] = cause != "pe_cancellation_next_to_zero"
            if len(examples) < 6:
                examples.append({"cell": int(c), "step": i, "cause": cause, "series": names})
        out["first_divergence_by_cause"] = causes
        out["examples"] = examples
        out["cells_with_a_decision_flip"] = int(persistent.sum())
        out["flip_rate_per_cell_year"] = float(persistent.sum() / (self.n * self.T / 8760.0))
        out["mean_steps_outside_per_diverged_cell"] = float(bad_any.sum() / max(1, flipped.sum()))
        out["forcing_outside_1e-11"] = dict(self.f_outside)
        out["forcing_worst_rel"] = dict(self.f_worst)
        return out


def census(got, want, forcing_got, forcing_want, tx, rtol=RTOL, **kw):
    """one-shot form: got / forcing_got cover the whole axis"""
    c = Census(want, forcing_want, tx, rtol, **kw)
    c.add_window(0, got, forcing_got)
    return c.result()
