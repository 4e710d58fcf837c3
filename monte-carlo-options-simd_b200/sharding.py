"""Multi-GPU sharding of the hot path: one process per GPU (torch.distributed), no data-path collective unless the
trials of the same options are split across GPUs.

The reference's only parallelism is shared-memory threads over the 8-path chunks of one option
(/root/reference/src/mc_simd.rs:49-51).  Across GPUs the work shards at two levels:
  * options-major  -- every rank prices a contiguous block of the batch on its own long_jump substream block; results
                      are independent, the caller just concatenates (no collective).
  * trials-minor   -- every rank simulates a disjoint trial range of the SAME options; the per-option moment sums
                      {sum, sum^2} x {call, put} (4 doubles per option) are summed with one all_reduce (NCCL over
                      NVLink on GPUs; gloo in the CPU tests) and finalised once.
"""
import numpy as np


def partition_options(n_options, world_size, rank):
    """Contiguous block [lo, hi) of the batch owned by `rank` (sizes differ by at most one): the options-major split of
    mcb200_shard_query."""
    n, w, r = int(n_options), int(world_size), int(rank)
    return r * n // w, (r + 1) * n // w


def partition_trials(num_trials, world_size, rank):
    """Trials simulated by `rank`, in whole 8-path chunks like the reference (mc_simd.rs:49).

    Returns (local_num_trials, total_paths): the local count is a multiple of 8; total_paths = 8*floor(num_trials/8)
    is what all ranks simulate together, so the sharded estimator equals the single-GPU one in distribution."""
    chunks = max(int(num_trials) // 8, 0)
    w, r = int(world_size), int(rank)
    mine = (r + 1) * chunks // w - r * chunks // w          # the trials-minor split of mcb200_shard_query
    return 8 * mine, 8 * chunks


def combine_moments(moments, group=None):
    """Sum per-option moment arrays across ranks in place (fp64).  `moments` is a torch tensor on the device the
    process group's backend works with (CUDA for nccl, CPU for gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(moments, op=dist.ReduceOp.SUM, group=group)
    return moments


def finalize_price_moments(moments, rate, years, total_paths, total_num_trials):
    """Host-side statement of the finalisation formula (the GPU path uses mcb200_price_finalize_device):
    price = sum / num_trials * exp(-rT); se = sample std / sqrt(n) * n / num_trials * exp(-rT)."""
    m = np.asarray(moments, np.float64).reshape(-1, 4)
    disc = np.exp(-(np.asarray(rate, np.float32) * np.asarray(years, np.float32)).astype(np.float32)).astype(np.float64)
    n = float(total_paths)
    out = {}
    for j, name in enumerate(("call", "put")):
        s, ss = m[:, 2 * j], m[:, 2 * j + 1]
        mu = s / n
        var = np.maximum((ss - n * mu * mu) / max(n - 1.0, 1.0), 0.0)
        out[name] = (s / total_num_trials * disc).astype(np.float32)
        out[name + "_se"] = (np.sqrt(var / n) * (n / total_num_trials) * disc).astype(np.float32)
    return out
