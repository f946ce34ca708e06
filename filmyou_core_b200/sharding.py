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


def _partition_fill(W, beta, world, T, G=None):
    n = len(W)
    c = 0
    while c < n and not W[c] > 0.0:
        c += 1
    rem = W[c] if c < n else 0.0
    pos = 0.0
    if G is not None:
        G[0] = 0.0
    for r in range(world):
        budget = T
        while c < n:
            fee = beta * W[c]
            if not budget > fee:
                break
            budget -= fee
            take = rem if rem < budget else budget
            rem -= take; budget -= take; pos += take
            if rem > 0.0:
                break
            c += 1
            while c < n and not W[c] > 0.0:
                c += 1
            rem = W[c] if c < n else 0.0
        if G is not None:
            G[r + 1] = pos
    return c >= n


def _partition_probe(lo, hi, l):
    w = hi - lo
    x = w * float(l + 1)
    return lo + x / 33.0


def partition_targets(W, beta, world):
    """Cumulative score work at the end of each rank when a rank also pays beta * W[c] once per cluster it touches (the
    H build of a straddled cluster is done by both neighbours): 33-section search (8 rounds of 32 probes) of the minimal
    maximum cost over a greedy fill.  Line for line the `partition_targets` / `k_shard_bounds` of csrc/rm2_kernels.cuh
    (same double arithmetic, every operation rounded separately)."""
    W = [float(x) for x in W]
    total = 0.0
    for x in W:
        total += x
    lo, hi = 0.0, total * (1.0 + beta) + 1.0
    for _ in range(8):
        first = 32
        for l in range(32):
            if _partition_fill(W, beta, world, _partition_probe(lo, hi, l)):
                first = l
                break
        nlo = _partition_probe(lo, hi, first - 1) if first > 0 else lo
        nhi = _partition_probe(lo, hi, first) if first < 32 else hi
        lo, hi = nlo, nhi
    G = [0.0] * (world + 1)
    _partition_fill(W, beta, world, hi, G)
    G[world] = total
    return G


def plan_shards(work, world, cluster_start=None, beta=0.55):
    """Boundaries [b_0=0, ..., b_world=len(work)] of contiguous ranges of the (cluster, user id)-ordered users.
    With `cluster_start` (first rank of every cluster + the end) the rule is the engine's: equal cost where a rank pays
    its users' work plus beta times the work of every cluster it touches (csrc/rm2_kernels.cuh k_shard_bounds); without
    it (or beta = 0) equal work: lower_bound on the inclusive prefix sum."""
    scan = np.cumsum(np.asarray(work, dtype=np.float64))
    tot = scan[-1]
    if cluster_start is not None and beta > 0.0:
        cs = np.asarray(cluster_start, dtype=np.int64)
        W = [float(scan[b - 1] - (scan[a - 1] if a > 0 else 0.0)) if b > a else 0.0 for a, b in zip(cs[:-1], cs[1:])]
        targets = partition_targets(W, beta, world)
    else:
        targets = [tot * r / world for r in range(world + 1)]
    b = [0]
    for r in range(1, world):
        x = int(np.searchsorted(scan, targets[r], side="left"))
        if cluster_start is not None and beta > 0.0 and x < len(scan):
            # a target that is the end of a cluster (up to rounding) falls ON the cluster boundary
            c = int(np.searchsorted(cs, x, side="right") - 1)
            e0, e1 = int(cs[c]), int(cs[c + 1])
            p0, p1 = (scan[e0 - 1] if e0 > 0 else 0.0), scan[e1 - 1]
            if abs(p0 - targets[r]) <= 1e-9 * tot:
                x = e0
            elif abs(p1 - targets[r]) <= 1e-9 * tot:
                x = e1
        b.append(max(b[-1], x))
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
