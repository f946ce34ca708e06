"""One process per GPU: users are partitioned across ranks, the rating matrix is replicated, and the
only exchange step is the final gather of the per-user top-N triples.

The GPU analogue of the reference's "replicate the cluster, score users with user % nSplits == split"
(M/common/AbstractByClusterAndCountMapper.java:86-102, M/rm/AbstractRM2Reducer.java:203-205): every
rank rebuilds the small per-cluster matrices it needs and scores a contiguous range of the
(cluster, user id)-ordered user list holding ~1/world of the estimated work sum(n_u * I_c).  The
partition itself is computed inside libfilmyou_rm2.so (fy_rm2_params.shard_rank/shard_count);
`plan_shards` restates it for tests and reports.
"""
import numpy as np
import torch
import torch.distributed as dist

FIELDS = (("user", torch.int32), ("item", torch.int32), ("score64", torch.float64),
          ("score32", torch.float32), ("cluster", torch.int32))


def plan_shards(work, world):
    """Boundaries [b_0=0, ..., b_world=len(work)] of contiguous ranges with ~equal total work;
    same rule as run_pipeline() in csrc/rm2_engine.cu (lower_bound on the inclusive prefix sum)."""
    scan = np.cumsum(np.asarray(work, dtype=np.float64))
    tot = scan[-1]
    b = [0]
    for r in range(1, world):
        b.append(int(np.searchsorted(scan, tot * r / world, side="left")))
    b.append(len(scan))
    return b


def gather_results(local, device=None, group=None):
    """All-gather the packed triples of every rank (NCCL on GPUs, gloo in the CPU tests).

    `local` maps field -> 1-D tensor (or numpy array) of this rank's triples, already ordered by
    (cluster, user id); ranks own increasing ranges of that order, so concatenating in rank order
    keeps the global order.  Returns field -> 1-D tensor with every rank's triples."""
    world = dist.get_world_size(group)
    tens = {}
    for name, dt in FIELDS:
        t = local[name]
        if not torch.is_tensor(t):
            t = torch.from_numpy(np.ascontiguousarray(t))
        tens[name] = t.to(device=device, dtype=dt) if device is not None else t.to(dtype=dt)
    dev = tens["user"].device
    n_local = torch.tensor([tens["user"].numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(max(counts), 1)
    out = {}
    for name, dt in FIELDS:
        pad = torch.zeros(n_max, dtype=dt, device=dev)
        pad[:tens[name].numel()] = tens[name]
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out[name] = torch.cat([p[:c] for p, c in zip(parts, counts)])
    return out
