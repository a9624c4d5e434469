"""Multi-GPU parity on hardware (core/region_model.h:873-900): N NCCL ranks, one per GPU, step the shards of a region whose
catchments straddle the shard boundaries; the all-reduced [T][n_catchments] discharge AND charge series must equal the single-GPU
model's catchment_discharges() / catchment_charges() to 1e-12 (the per-cell results are bit-identical; only the order in which a
straddling catchment's two halves are added differs).  Skipped when fewer than 2 devices are visible."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from fixtures import PTGSK_DEFAULT
from parity import assert_parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_all_reduced_catchment_series_equal_the_single_gpu_run(tmp_path, world):
    if _device_count() < world:
        pytest.skip(f"needs {world} CUDA devices, {_device_count()} visible")
    import shyft_b200 as sb
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from multi_gpu_worker import region
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
           str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    red = np.load(tmp_path / "reduced.npz")
    assert int(red["world"]) == world and int(red["launches"]) > 0
    geo, ta, env, st0 = region()
    m = sb.PTGSKOptModel(geo, PTGSK_DEFAULT)
    m.initialize_cell_environment(ta)
    m._set_sources(env)
    m.set_states(st0)
    m.run_windowed(sb.InterpolationParameter(), window_steps=512)
    assert np.array_equal(red["cids"], m.catchment_ids)       # global catchment index = first appearance over the whole cell vector
    q1, c1 = m.catchment_discharges(), m.catchment_charges()
    assert np.all(np.isfinite(q1)) and q1.max() > 0.0
    assert_parity(red["q"], q1, f"all-reduced catchment discharge, {world} ranks", rtol=1e-12)
    assert_parity(red["c"], c1, f"all-reduced catchment charge, {world} ranks", rtol=1e-12, atol_frac=1e-12)
    # (not bit-identical even for a catchment that lies wholly inside one shard: the deterministic segmented reduction groups a catchment's
    # cells by warp, and a shard's warps start at another cell than the whole region's)
