#!/usr/bin/env python
"""bench.py -- cell-timesteps/s of the pt_gs_k run_cells hot path (BASELINE.json), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--cells C] [--years Y] [--window S]
  python bench.py --impl reference ...      # the CPU path (oracle port; the real reference cannot be built here)

A "step" is one pass of the hot path over the whole workload: BASELINE configs[1], pt_gs_k, 100 000 cells x 10 years
hourly (87 600 steps), BTK temperature + IDW for the other four variables, discharge collector (56 algorithmic bytes
per cell-step), run window by window because the [time][cell] forcing (350 GB) cannot be resident.  At N > 1 every rank
steps its own 100 000-cell shard of an N x 100 000-cell region (weak scaling) and the per-catchment discharge series
are summed across ranks with NCCL.

  value     cell-steps / device time with the station series already in HBM (CUDA events on the launching stream)
  e2e       same metric through the public API with HOST buffers: station series + states H2D (pinned memory),
            catchment discharge + end states D2H, inside the timed region
  roofline  the step kernel alone: 56 B x cell-steps / its CUDA-event time vs the measured HBM copy peak
  cpu_baseline  the CPU oracle's threaded run_cells on a bounded sample, same inputs, rank 0 only
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_CELL_STEP = 56  # 5 forcings read + avg_discharge, charge_m3s written
TRAFFIC_PER_CELL_STEP = 178.0  # ncu dram__bytes_read+write of the three kernels / cell-steps of the captured window (9.12 GB / 51.2 M)
PTGSK_DEFAULT = [-2.439, 0.966, -0.10, 1.5, -0.5, 2.0, 0.1, 1.0, 5.0, 5.0, 30.0, 0.9, 0.6, 5.0, 0.4, 0.4, 1.0, 0.0, 0.0, 0.2, 1.26, 0.04, 100.0, 0.0,
                 6.0, 1.0, 7.0, 0.0, 221.0, 0.0, 1.0]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=100000, help="cells per GPU")
    ap.add_argument("--years", type=float, default=10.0)
    ap.add_argument("--stations", type=int, default=64)
    ap.add_argument("--window", type=int, default=2048, help="time steps per forcing window (measured: 512 -> 7.68, 1024 -> 7.90, 2048 -> 7.96 G cell-steps/s)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_workload(args, rank, world):
    from shyft_b200 import synthetic
    n_steps = int(round(args.years * 8760))
    n_total = args.cells * world
    # the whole region is generated identically on every rank; each rank keeps its shard (cells are independent)
    geo, ta, env = synthetic.make_region(n_total, n_steps, args.stations, config_index=1)
    from shyft_b200.sharding import partition_cells
    b, e = partition_cells(n_total, world, rank)
    return geo, geo[b:e], ta, env, synthetic.default_state(0, e - b)


def pinned_copy(a):
    import torch
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    v = t.numpy()
    v[...] = a
    return v, t


def cpu_baseline(args, geo_local, ta, env, target_seconds):
    """The oracle's threaded run_cells (the reference's work-queue shape, core/region_model.h:991-1021) on a bounded sample."""
    from oracle import oracle as O
    geo_matrix = O.geo_matrix
    cores = O.hardware_concurrency()
    T = min(ta.n, 8760)
    dt_us = ta.delta_t * 10**6

    def forcing(gm):
        f = {}
        for name in ("temperature", "precipitation", "radiation", "wind_speed", "rel_hum"):
            xyz, vals = getattr(env, name)
            vals = O.average_accessor_same_axis(vals[:T], dt_us)
            if name == "temperature":
                f[name] = O.btk_run(xyz, vals, gm[:, :3], ta.start * 10**6, dt_us)
            else:
                f[name] = O.idw_run(name, xyz, vals, gm[:, :3], O.idw_par(max_members=20 if name == "precipitation" else 10), dst_slope=gm[:, 5],
                                    ncore=cores)
        return f

    def timed(n):
        idx = np.linspace(0, geo_local.shape[0] - 1, n).astype(np.int64)  # spread over the shard: all elevations / climates
        gm = geo_matrix(geo_local[idx])
        f = forcing(gm)
        st = np.tile(np.array([0.4, 0.1, 30000.0, 1.26, 0.0, 0.0, 0.0, 0.0, 0.8]), (n, 1))
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            O.ptgsk_run_cells(gm, PTGSK_DEFAULT, f, st, ta.start * 10**6, dt_us, collect_response=False, ncore=cores)
            el = time.perf_counter() - t0
            best = el if best is None else min(best, el)
        return n * T / best, best

    rate, el = timed(max(64, 8 * cores))
    n = int(min(geo_local.shape[0], max(64, rate * target_seconds / 2 / T)))  # two timed repeats
    rate, el = timed(n)
    return {"value": rate, "unit": "cell-timesteps/s", "cores": cores, "kind": "port",
            "sample": f"{n} cells (evenly spread over the shard) x first {T} steps, oracle run_cells with {cores} threads, best of 2 ({el:.2f} s)"}


def workload_name(args, n, T):
    return (f"pt_gs_k {n} cells/GPU x {T} hourly steps ({args.years:g} y), BTK temperature + IDW precipitation/radiation/wind/rel_hum, "
            f"{args.stations} stations, discharge collector, windows of {args.window} steps")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    geo, geo_local, ta, env, st0 = build_workload(args, 0, 1)
    vals = []
    sample = ""
    cores = 0
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(args, geo_local, ta, env, max(2.0, args.cpu_seconds / 2))
        cores, sample = r["cores"], r["sample"]
        if i >= args.warmup:
            vals.append(r["value"])
    v = statistics.mean(vals)
    n_steps = ta.n
    print(json.dumps({
        "impl": "reference", "metric": "cell-timesteps/sec (pt_gs_k run_cells)", "value": v, "unit": "cell-timesteps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * args.cells * n_steps / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, args.cells, n_steps), "cells_per_gpu": args.cells, "n_steps": n_steps,
                   "note": "the CPU oracle's run_cells (interpolated forcing resident in host memory) on a bounded sample of the same cells; "
                           "ms_per_step extrapolated from the sample to the whole workload"},
        "cpu_baseline": {"value": v, "unit": "cell-timesteps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "cell-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import shyft_b200 as sb
    from shyft_b200 import sharding

    geo_all, geo, ta, env, st0 = build_workload(args, rank, world)
    n, T = geo.shape[0], ta.n
    _, global_cids = sharding.global_catchment_index(geo_all["catchment_id"])
    m = sb.PTGSKOptModel(geo, PTGSK_DEFAULT, device=local_rank)
    ip = sb.InterpolationParameter()  # BTK temperature + IDW (defaults of core/region_model.h:65-95)
    # host buffers of the end-to-end leg live in pinned memory
    keep = []
    env_pinned = sb.RegionEnvironment()
    for name in sb.capi.FORCING_NAMES:
        xyz, vals = getattr(env, name)
        v, t = pinned_copy(vals)
        keep.append(t)
        setattr(env_pinned, name, (xyz, v))
    st_pinned, t = pinned_copy(st0)
    keep.append(t)
    q_pinned, t = pinned_copy(np.zeros((T, m.number_of_catchments())))   # results of the end-to-end leg land in pinned memory too
    keep.append(t)
    s_pinned, t = pinned_copy(np.zeros_like(st0))
    keep.append(t)
    h2d = sum(getattr(env, k)[1].nbytes + getattr(env, k)[0].nbytes for k in sb.capi.FORCING_NAMES) + st0.nbytes
    d2h = T * m.number_of_catchments() * 8 + st0.nbytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_catchments():
        if world == 1:
            return None
        local = sharding.device_catchment_discharges(m)
        g = sharding.scatter_local_to_global(local, m.catchment_ids, global_cids, xp=torch)
        return sharding.all_reduce_catchment_series(g)

    def device_pass():
        m.revert_to_initial_state()
        m.run_windowed(ip, window_steps=args.window)
        reduce_catchments()

    e2e_parts = {}

    def e2e_pass():
        t = [time.perf_counter()]

        def lap(name):
            t.append(time.perf_counter())
            e2e_parts.setdefault(name, []).append(1000.0 * (t[-1] - t[-2]))
        m.initialize_cell_environment(ta); lap("initialize_cell_environment")
        m._set_sources(env_pinned); lap("set_sources_h2d")
        m.initial_state = st_pinned
        m.set_states(st_pinned); lap("set_states_h2d")
        m.run_windowed(ip, window_steps=args.window); lap("run_windowed")
        reduce_catchments(); lap("all_reduce")
        out = m.catchment_discharges(out=q_pinned), m.get_states(out=s_pinned); lap("results_d2h")
        return out

    def timed(fn, k):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # resident-input leg
    m.initialize_cell_environment(ta)
    m._set_sources(env)
    m.set_states(st0)
    for _ in range(args.warmup):
        device_pass()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = m.kernel_launches()
    step_ms_acc, interp_ms_acc = [], []

    def device_pass_recorded():
        device_pass()
        s, i = m.last_run_kernel_ms()
        step_ms_acc.append(s)
        interp_ms_acc.append(i)

    total_ms = timed(device_pass_recorded, args.steps)
    launches = m.kernel_launches() - l0
    clocks = sampler.stop() if sampler else None
    cq_device = m.catchment_discharges()
    # end-to-end leg (host buffers)
    e2e_pass()
    wall = []
    barrier()
    for _ in range(max(1, min(args.steps, 3))):
        t0 = time.perf_counter()
        cq_e2e, _ = e2e_pass()
        torch.cuda.synchronize()
        wall.append(time.perf_counter() - t0)
    e2e_s = torch.tensor([statistics.mean(wall)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    if not np.array_equal(cq_e2e, cq_device):
        raise SystemExit("bench.py: the end-to-end pass and the resident pass disagree")
    if not np.all(np.isfinite(cq_device)):
        raise SystemExit("bench.py: non-finite catchment discharge")

    if rank == 0:
        cell_steps = float(n) * T * world
        ms_per_step = total_ms / args.steps
        value = cell_steps / (ms_per_step / 1000.0)
        peak, peak_src = measured_peaks()
        k_ms = statistics.mean(step_ms_acc)  # rank 0's step kernels per pass
        achieved = BYTES_PER_CELL_STEP * float(n) * T / (k_ms / 1000.0) / 1e9
        n_windows = (T + args.window - 1) // args.window
        out = {
            "metric": "cell-timesteps/sec (pt_gs_k run_cells)", "value": value, "unit": "cell-timesteps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args, n, T),
                       "cells_per_gpu": n, "n_steps": T, "window_steps": args.window,
                       "l2": "inputs larger than L2: every window streams %.0f MB of forcing + series per pass" % (n * args.window * 56 / 1e6),
                       "interp_ms_per_step": statistics.mean(interp_ms_acc), "step_kernel_ms_per_step": k_ms},
            "e2e": {"value": cell_steps / float(e2e_s.item()), "unit": "cell-timesteps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "rank0_ms": {k: round(statistics.mean(v[1:]), 2) for k, v in e2e_parts.items()}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            # run_cells of one window = the three kernels of the phase pipeline, launched back to back; "achieved" divides the
            # ALGORITHMIC bytes (56 B per cell-step) by their summed CUDA-event time.  "traffic" is the DRAM
            # traffic ncu measures for one 512-step window of 100 000 cells (profiles/ncu_pipeline_r01_m_winter_window.txt: 3 x the
            # algorithmic bytes, because the phases hand five scratch arrays to each other through HBM -- the stack is fp64-bound).
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": TRAFFIC_PER_CELL_STEP * n * min(args.window, T),
                         "kernel": "run_cells window = ptgsk_forcing_terms_kernel + ptgsk_snow_kernel<0> + ptgsk_response_kernel<1>",
                         "launches_per_step": n_windows, "avg_launch_ms": k_ms / n_windows,
                         "algorithmic_bytes_per_launch": BYTES_PER_CELL_STEP * n * min(args.window, T), "peak_source": peak_src,
                         "binding_roof": {"pipe": "fp64", "pipe_active_pct_ncu": {"forcing_terms": 70.2, "snow": 54.6, "response": 70.1},
                                          "source": "profiles/ncu_pipeline_r01_m_winter_window.txt (sm__pipe_fp64_cycles_active, winter window)"},
                         "note": "fp64-compute bound, not HBM bound: 16-25 exp/log, an adaptive ODE step and incomplete gamma functions "
                                 "per cell-step against 56 bytes (DESIGN.md section 3, SURVEY H3)"},
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args, geo, ta, env, args.cpu_seconds)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
