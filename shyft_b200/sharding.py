"""Cell sharding across the GPUs of one box: cells are independent during run_cells
(core/region_model.h:556-561), so each rank steps a contiguous range of the cell vector with no data-path collective; the
only exchange is the fp64 sum of per-catchment discharge/charge series, because a catchment may straddle a shard
boundary.  torch.distributed is the plumbing (nccl on the GPUs, gloo in the CPU tests).
"""
import numpy as np


def partition_cells(n_cells, world_size, rank):
    """Contiguous, balanced ranges in the caller's cell order -> (begin, end)."""
    base, rem = divmod(int(n_cells), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def global_catchment_index(catchment_ids):
    """cix in order of first appearance over the WHOLE cell vector (core/region_model.h:233-249) -> (cix per cell, cix -> cid)."""
    cid = np.asarray(catchment_ids, dtype=np.int64)
    uniq, first = np.unique(cid, return_index=True)
    order = np.argsort(first, kind="stable")
    cix_to_cid = uniq[order]
    rank_of = np.empty_like(order)
    rank_of[order] = np.arange(order.size)
    return rank_of[np.searchsorted(uniq, cid)], cix_to_cid


def scatter_local_to_global(local_series, local_cids, global_cids, xp=np):
    """local [T][n_local_catchments] -> zero-filled [T][n_global_catchments] with the local columns at their global cix."""
    pos = {int(c): i for i, c in enumerate(np.asarray(global_cids))}
    cols = [pos[int(c)] for c in np.asarray(local_cids)]
    if xp is np:
        out = np.zeros((local_series.shape[0], len(pos)), dtype=np.float64)
        out[:, cols] = local_series
        return out
    import torch
    out = torch.zeros((local_series.shape[0], len(pos)), dtype=torch.float64, device=local_series.device)
    out[:, torch.as_tensor(cols, device=local_series.device)] = local_series
    return out


def all_reduce_catchment_series(global_partial):
    """Sum the per-rank partial catchment series in place (torch tensor, any backend) and return it."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(global_partial, op=dist.ReduceOp.SUM)
    return global_partial


class DeviceArrayView:
    """Wraps a raw device pointer of the library ([rows][cols] fp64) for torch.as_tensor via __cuda_array_interface__."""

    def __init__(self, ptr, rows, cols):
        self.__cuda_array_interface__ = {"shape": (int(rows), int(cols)), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def device_catchment_discharges(model):
    """The model's catchment discharge sums as a torch tensor aliasing library memory [T][n_catchments]."""
    import torch
    ptr, rows, cols = model.device_catchment_discharges()
    return torch.as_tensor(DeviceArrayView(ptr, rows, cols), device="cuda")


def device_catchment_charges(model):
    """The model's catchment charge sums as a torch tensor aliasing library memory [T][n_catchments]."""
    import torch
    ptr, rows, cols = model.device_catchment_charges()
    return torch.as_tensor(DeviceArrayView(ptr, rows, cols), device="cuda")


def global_catchment_series(model, global_cids, what="discharge"):
    """The region's [T][n_global_catchments] catchment series on every rank: this rank's catchment sums placed at their global
    catchment index, then summed over the ranks (NCCL all-reduce; a catchment that straddles a shard boundary gets both halves).
    The semantics of region_model::get_catchment_discharges / charges over the whole cell vector (core/region_model.h:873-900)."""
    local = device_catchment_discharges(model) if what == "discharge" else device_catchment_charges(model)
    g = scatter_local_to_global(local, model.catchment_ids, global_cids, xp=__import__("torch"))
    return all_reduce_catchment_series(g)


# ---- calibration ensembles: the parameter sets are sharded, the cells are not ----------------------------------
def partition_parameter_sets(n_sets, world_size, rank):
    """Contiguous, balanced range of a population's parameter sets for this rank -> (begin, end); every rank holds all cells."""
    return partition_cells(n_sets, world_size, rank)


def gather_goal_values(local_goals, n_sets, device=None):
    """All ranks' goal-function values in population order on every rank: one all_gather of one fp64 per set (ragged shards padded),
    no collective on the data path.  `local_goals` = this rank's values for partition_parameter_sets(n_sets, world, rank)."""
    import torch
    import torch.distributed as dist
    local = torch.as_tensor(np.asarray(local_goals, dtype=np.float64))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local.numpy().copy()
    world, rank = dist.get_world_size(), dist.get_rank()
    width = -(-int(n_sets) // world)
    buf = torch.full((width,), float("nan"), dtype=torch.float64, device=device)
    buf[: local.numel()] = local.to(buf.device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = np.empty(int(n_sets))
    for r in range(world):
        b, e = partition_parameter_sets(n_sets, world, r)
        out[b:e] = parts[r][: e - b].cpu().numpy()
    return out
