"""Multi-GPU: configurations are independent, so a batch is range-sharded over ranks (one process per GPU, the scene
replicated on each) and only the per-configuration results travel.

  shard_range(n, rank, world)        contiguous slice [rank*n/world, (rank+1)*n/world)
  gather_results(arrays, n, ...)     all_gather of verdict / first-pair words into full-length arrays on every rank
  check_batch_sharded(fn, q, ...)    run `fn` (e.g. BoundQuery.check) on this rank's slice and gather

The collective is torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests); it carries 12 bytes per
configuration and is not on the data path of the kernels.
"""
import numpy as np


def shard_range(n, rank, world):
    return (rank * n) // world, ((rank + 1) * n) // world


def shard_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_results(local_arrays, n, rank, world, device=None):
    """local_arrays: list of 1-D integer numpy arrays (this rank's slice).  Returns full-length arrays, identical on every rank."""
    if world == 1:
        return [np.asarray(a) for a in local_arrays]
    import torch
    import torch.distributed as dist
    sizes = shard_sizes(n, world)
    width = max(sizes) if sizes else 0
    k = len(local_arrays)
    pack = np.zeros((k, width), dtype=np.int32)
    for i, a in enumerate(local_arrays):
        pack[i, :len(a)] = np.asarray(a).astype(np.int32, copy=False)
    t = torch.from_numpy(pack)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = [np.zeros(n, dtype=np.int32) for _ in range(k)]
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        o = outs[r].cpu().numpy()
        for i in range(k):
            full[i][lo:hi] = o[i, :hi - lo]
    return full


def check_batch_sharded(check_fn, q, rank, world, device=None):
    """check_fn(q_slice) -> (verdict, first_env, first_self[, ...]) for this rank's configurations; returns the three
    full-length arrays (verdict as uint32)."""
    n = len(q)
    lo, hi = shard_range(n, rank, world)
    res = check_fn(q[lo:hi])
    v, fe, fs = res[0], res[1], res[2]
    fv, ffe, ffs = gather_results([v.view(np.int32) if v.dtype == np.uint32 else v, fe, fs], n, rank, world, device)
    return fv.view(np.uint32), ffe, ffs
