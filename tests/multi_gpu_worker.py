"""Rank program of tests/test_gpu_multi.py (launched by torch.distributed.run, one rank per GPU): each rank steps its shard of a region
whose catchments straddle the shard boundaries through sb2_run_windowed, the per-catchment discharge and charge series are placed at their
global catchment index and summed over the ranks with NCCL (shyft_b200/sharding.py).  Rank 0 writes the reduced series."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def region():
    from shyft_b200 import synthetic
    # 3 001 cells, 700 per catchment: with 2 ranks the boundary (cell 1501) falls inside catchment 3, with 4 ranks inside 2, 3 and 4
    geo, ta, env = synthetic.make_region(3001, 1500, 16, config_index=7, cells_per_catchment=700, start=1417392000)  # 2014-12-01
    return geo, ta, env, synthetic.default_state(0, 3001)


def main():
    import torch
    import torch.distributed as dist
    import shyft_b200 as sb
    from fixtures import PTGSK_DEFAULT
    from shyft_b200 import sharding
    out_dir = sys.argv[1]
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    geo, ta, env, st0 = region()
    b, e = sharding.partition_cells(geo.shape[0], world, rank)
    _, gcids = sharding.global_catchment_index(geo["catchment_id"])
    m = sb.PTGSKOptModel(geo[b:e], PTGSK_DEFAULT, device=local_rank)
    m.initialize_cell_environment(ta)
    m._set_sources(env)
    m.set_states(st0[b:e])
    m.run_windowed(sb.InterpolationParameter(), window_steps=512)
    q = sharding.global_catchment_series(m, gcids, "discharge")
    c = sharding.global_catchment_series(m, gcids, "charge")
    torch.cuda.synchronize()
    if rank == 0:
        np.savez(os.path.join(out_dir, "reduced.npz"), q=q.cpu().numpy(), c=c.cpu().numpy(), cids=gcids, world=world, launches=m.kernel_launches())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
