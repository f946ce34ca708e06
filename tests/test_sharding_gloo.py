"""CPU suite: the N>1 host path (work partition + gather of the packed top-N triples) with
world_size 2 over gloo.  The per-rank triples come from the oracle here (no GPU in this container);
on the GPU box the same gather runs over NCCL in bench.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from filmyou_core_b200 import datagen, sharding
from oracle import rm2_oracle as orc

from conftest import by_user


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = datagen.generate("tiny")
    # the engine's rule: contiguous ranges of the (cluster, user id) order with equal sum n_u * I_c
    order = np.lexsort((r.cl_user, r.cl_cluster))
    n_u = np.bincount(r.user, minlength=r.n_users + 1)
    cl_of = np.zeros(r.n_users + 1, np.int64); cl_of[r.cl_user] = r.cl_cluster
    i_c = np.array([len(np.unique(r.item[cl_of[r.user] == c])) for c in range(r.n_clusters)])
    work = n_u[r.cl_user[order]] * i_c[r.cl_cluster[order]]
    b = sharding.plan_shards(work, world)
    mine = r.cl_user[order][b[rank]:b[rank + 1]]
    local = orc.run(r.user, r.item, r.score, r.cl_user, r.cl_cluster, r.cluster_size, 0.1, r.n_items, 10,
                    only_users=mine) if len(mine) else {k: np.zeros(0) for k, _ in sharding.FIELDS}
    out = sharding.gather_results({k: local[k] for k, _ in sharding.FIELDS})
    if rank == 0:
        np.savez(os.path.join(tmp, "gathered.npz"), **{k: v.numpy() for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gather_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = dict(np.load(os.path.join(str(tmp_path), "gathered.npz")))
    r = datagen.generate("tiny")
    want = orc.run(r.user, r.item, r.score, r.cl_user, r.cl_cluster, r.cluster_size, 0.1, r.n_items, 10)
    for k, _ in sharding.FIELDS:
        assert np.array_equal(got[k], want[k]), k          # same triples, same global order


def test_plan_shards_balances_work_and_covers_everything():
    rng = np.random.default_rng(0)
    work = rng.lognormal(0, 1, 5000)
    for world in (1, 2, 4, 8):
        b = sharding.plan_shards(work, world)
        assert b[0] == 0 and b[-1] == len(work) and all(x <= y for x, y in zip(b, b[1:]))
        parts = [work[b[k]:b[k + 1]].sum() for k in range(world)]
        assert max(parts) <= work.sum() / world + work.max() + 1e-9


def test_build_aware_partition_balances_cost_including_the_duplicated_builds():
    # 50 clusters of roughly equal work over 8 ranks: with beta > 0 the rank that touches one more cluster gets less score
    # work, so the maximum COST (work + beta * touched clusters' work) is lower than with the equal-work rule
    rng = np.random.default_rng(1)
    sizes = rng.integers(2500, 3000, 50)
    cs = np.concatenate([[0], np.cumsum(sizes)])
    work = rng.lognormal(0, 1, cs[-1]) * 1000
    W = np.array([work[a:b].sum() for a, b in zip(cs[:-1], cs[1:])])

    def max_cost(b, beta=0.55):
        out = []
        for k in range(len(b) - 1):
            if b[k + 1] <= b[k]:
                out.append(0.0); continue
            c0 = np.searchsorted(cs, b[k], side="right") - 1
            c1 = np.searchsorted(cs, b[k + 1] - 1, side="right") - 1
            out.append(work[b[k]:b[k + 1]].sum() + beta * W[c0:c1 + 1].sum())
        return max(out)
    for world in (2, 4, 8):
        plain = sharding.plan_shards(work, world)
        aware = sharding.plan_shards(work, world, cluster_start=cs)
        assert aware[0] == 0 and aware[-1] == len(work) and all(x <= y for x, y in zip(aware, aware[1:]))
        assert max_cost(aware) <= max_cost(plain) + 2.0 * work.max()      # boundaries fall on whole users
    assert max_cost(sharding.plan_shards(work, 8, cluster_start=cs)) < 0.98 * max_cost(sharding.plan_shards(work, 8))
    assert sharding.plan_shards(work, 1, cluster_start=cs) == [0, len(work)]
