"""The N > 1 path on CPU: world_size-2 gloo processes exercise the cell partition, the global catchment indexing and the
catchment-series all-reduce of shyft_b200/sharding.py.  The CUDA kernels are not involved: each rank feeds
the per-cell discharge of its shard (taken from the CPU oracle) through the same scatter + all_reduce the GPU path uses."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_cells, T, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from shyft_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)                     # every rank builds the same region description
    cids = 10 + rng.integers(0, 7, n_cells) * 3          # catchments interleaved: they straddle the shard boundary
    q = rng.random((T, n_cells))                         # stand-in for per-cell avg_discharge [T][cell]
    gcix, gcids = sharding.global_catchment_index(cids)
    b, e = sharding.partition_cells(n_cells, world, rank)
    # what a rank's model reports: catchment sums over ITS cells, in ITS first-appearance order
    lcix, lcids = sharding.global_catchment_index(cids[b:e])
    local = np.zeros((T, lcids.size))
    for k in range(lcids.size):
        local[:, k] = q[:, b:e][:, lcix == k].sum(axis=1)
    g = sharding.scatter_local_to_global(torch.from_numpy(local), lcids, gcids, xp=torch)
    sharding.all_reduce_catchment_series(g)
    want = np.zeros((T, gcids.size))
    for k in range(gcids.size):
        want[:, k] = q[:, gcix == k].sum(axis=1)
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([np.allclose(g.numpy(), want, rtol=1e-13, atol=0), b, e]))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_contiguous_balanced_and_complete():
    sys.path.insert(0, ROOT)
    from shyft_b200 import sharding
    for n, w in [(10, 3), (1000000, 8), (7, 8), (100000, 2)]:
        parts = [sharding.partition_cells(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [e - b for b, e in parts]
        assert max(sizes) - min(sizes) <= 1


def test_global_catchment_index_is_first_appearance_order(oracle):
    sys.path.insert(0, ROOT)
    from shyft_b200 import sharding
    cids = np.array([7, 3, 7, 9, 3, 1, 9, 9, 2])
    cix, ids = sharding.global_catchment_index(cids)
    ocix, oids = oracle.catchment_index(cids)
    assert np.array_equal(cix, ocix) and np.array_equal(ids, oids)


def test_world_size_2_gloo_catchment_reduce(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 1001, 48, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        ok, b, e = np.load(tmp_path / f"ok_{r}.npy")
        assert ok == 1.0
    assert np.load(tmp_path / "ok_0.npy")[2] == np.load(tmp_path / "ok_1.npy")[1]


def _ensemble_worker(rank, world, port, n_sets, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from shyft_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    goal = lambda i: 0.25 * i * i - 3.0 * i            # stand-in for one goal-function evaluation of parameter set i
    b, e = sharding.partition_parameter_sets(n_sets, world, rank)
    got = sharding.gather_goal_values([goal(i) for i in range(b, e)], n_sets)
    np.save(os.path.join(out_dir, f"goals_{rank}.npy"), got)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_parameter_set_sharding(tmp_path):
    """BASELINE config 5: the population is split over the ranks, every rank gets all goal values back in population order"""
    import torch.multiprocessing as mp
    n_sets = 37                                          # ragged: 19 + 18
    mp.spawn(_ensemble_worker, args=(2, _free_port(), n_sets, str(tmp_path)), nprocs=2, join=True)
    want = np.array([0.25 * i * i - 3.0 * i for i in range(n_sets)])
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"goals_{r}.npy"), want)


def _oracle_shard_worker(rank, world, port, out_dir):
    """each rank steps ITS cells with the CPU oracle (the stand-in for the device model of a shard), reports catchment sums in its own
    first-appearance order, and the shards' series meet in the all-reduce -- the flow of bench.py --gpus N / tests/test_gpu_multi.py"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from fixtures import PTGSK_DEFAULT
    from oracle import oracle as O
    from shyft_b200 import sharding, synthetic
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, T = 90, 240
    geo, ta, env = synthetic.make_region(n, T, 6, config_index=7, cells_per_catchment=40, start=1417392000)   # catchment 2 straddles cell 45
    gm = O.geo_matrix(geo)
    f = {}
    for name in ("temperature", "precipitation", "radiation", "wind_speed", "rel_hum"):
        xyz, vals = getattr(env, name)
        f[name] = O.idw_run(name, xyz, O.average_accessor_same_axis(vals, 3600 * 10**6), gm[:, :3], O.idw_par(), dst_slope=gm[:, 5])
    st0 = synthetic.default_state(0, n)
    b, e = sharding.partition_cells(n, world, rank)
    run = lambda lo, hi: O.ptgsk_run_cells(gm[lo:hi], PTGSK_DEFAULT, {k: np.ascontiguousarray(v[:, lo:hi]) for k, v in f.items()}, st0[lo:hi],
                                           ta.start * 10**6, 3600 * 10**6)["avg_discharge"]
    q = run(b, e)
    _, gcids = sharding.global_catchment_index(geo["catchment_id"])
    lcix, lcids = sharding.global_catchment_index(geo["catchment_id"][b:e])
    local = np.stack([q[:, lcix == k].sum(axis=1) for k in range(lcids.size)], axis=1)
    g = sharding.all_reduce_catchment_series(sharding.scatter_local_to_global(torch.from_numpy(local), lcids, gcids, xp=torch))
    if rank == 0:
        q_all = run(0, n)
        gcix, _ = sharding.global_catchment_index(geo["catchment_id"])
        want = np.stack([q_all[:, gcix == k].sum(axis=1) for k in range(gcids.size)], axis=1)
        np.save(os.path.join(out_dir, "oracle_shards.npy"), np.array([np.allclose(g.numpy(), want, rtol=1e-12, atol=0), float(want.max() > 0), gcids.size]))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_sharded_oracle_run_equals_whole_region(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_oracle_shard_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    ok, positive, n_catch = np.load(tmp_path / "oracle_shards.npy")
    assert ok == 1.0 and positive == 1.0 and n_catch == 3
