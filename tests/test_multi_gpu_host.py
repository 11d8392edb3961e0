"""N > 1 host logic on CPU: world-size-2 `gloo` run of the base-range MSM sharding used by
`bench.py --mode msm-sharded` (slice the bases, partial MSM per rank, all_gather of the 96-byte points, fold).
The per-rank partial MSM is computed by the oracle here (no GPU in this container); on the GPU box the same
partition / gather / fold runs with zk_msm_dev + NCCL + zk_g1_sum."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = n // world
    bases = orc.gen_points(7, n, threads=2)
    sc = orc.gen_scalars(8, n)
    part = orc.best_multiexp(sc[rank * per:(rank + 1) * per], bases[rank * per:(rank + 1) * per], threads=2)
    # normalise like the library does (z = 1) so that the gathered representation is canonical
    t = torch.from_numpy(part.astype(np.int64))
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    acc = gathered[0].numpy().astype(np.uint64)
    for g in gathered[1:]:
        acc = orc.g1_add(acc, g.numpy().astype(np.uint64))
    # value timing reduction used by bench.py: max over ranks
    tm = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = orc.best_multiexp(sc, bases, threads=2)
        q.put((orc.g1_to_affine(acc).tolist(), orc.g1_to_affine(full).tolist(), float(tm.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_msm_base_range_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n, world = 1 << 10, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, want, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == want
    assert tmax == 2.0


def test_group_shard_plan_covers_everything_once():
    """zk_group_shard_plan (host logic of the one-proof-over-several-GPUs mode): column shares are disjoint, padded to
    equal size and cover every column; base-range shares tile [0, n)."""
    import zkdcap_b200 as zk
    n = 1 << 19
    for world in (1, 2, 4, 8):
        for count in (1, 2, 3, 4, 6, 7, 11, 22, 28):
            plans = [zk.Group.shard_plan(r, world, count, n) for r in range(world)]
            modes = {p[0] for p in plans}
            assert len(modes) == 1
            mode, per = plans[0][0], plans[0][1]
            total = count if mode == 0 else n
            assert mode == (0 if (count >= world or world == 1) else 1)
            covered = []
            for r, (_, p, lo, hi) in enumerate(plans):
                assert p == per and hi - lo <= per and lo <= hi
                if hi > lo:
                    assert lo == r * per          # rank r's share sits at slot r * per of the exchange buffer
                covered += list(range(lo, hi)) if total < 1000 else [(lo, hi)]
            if total < 1000:
                assert covered == list(range(total))
            else:
                assert covered[0][0] == 0 and covered[-1][1] == n and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


def test_group_balance_plan_properties():
    """zk_group_balance_plan (csrc/prover.cu): every column is placed exactly once on a valid rank, no rank gets more than
    ceil(count / world) columns, the plan is deterministic, and the heaviest rank is no worse than round-robin on skewed costs
    (bit columns next to field-element columns, as in the QE3 shape)."""
    import ctypes as C
    import random
    import zkdcap_b200 as zk
    lib = zk._abi.load()
    rnd = random.Random(5)
    for world in (2, 4, 8):
        for count in (world, world + 1, 17, 22, 28, 3 * world + 2):
            if count < world:
                continue
            cost = [rnd.choice([1, 1, 2, 16, 16, 11]) * 524288 + rnd.randrange(1000) for _ in range(count)]
            arr = (C.c_uint32 * count)(*cost)
            own = (C.c_uint32 * count)()
            lib.zk_group_balance_plan(arr, count, world, own)
            own2 = (C.c_uint32 * count)()
            lib.zk_group_balance_plan(arr, count, world, own2)
            assert list(own) == list(own2)
            assert all(0 <= o < world for o in own)
            cap = -(-count // world)
            loads = [0] * world
            for r in range(world):
                mine = [i for i in range(count) if own[i] == r]
                assert len(mine) <= cap
                loads[r] = sum(cost[i] for i in mine)
            rr = [sum(cost[i] for i in range(count) if i % world == r) for r in range(world)]
            assert max(loads) <= max(rr)
            assert max(loads) <= sum(cost) / world + max(cost)          # the longest-first bound
