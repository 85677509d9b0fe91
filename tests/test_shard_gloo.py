"""CPU, world_size 2 over gloo: the N>1 host path (shard -> local work -> gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_units, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sstem_restoration_b200 import shard
    lo, hi = shard.shard_range(n_units, rank, world)
    # a "restored section" for unit k is a constant plane of value k
    local = torch.stack([torch.full((2, 3), float(k)) for k in range(lo, hi)]) if hi > lo else torch.zeros((0, 2, 3))
    full = shard.gather_sections(local, n_units)
    ok = full.shape == (n_units, 2, 3) and all(bool((full[k] == k).all()) for k in range(n_units))
    # gather to one rank only (the DataParallel pattern): rank 1 receives, rank 0 gets None
    one = shard.gather_sections(local, n_units, dst=1)
    if rank == 1:
        ok = ok and one.shape == (n_units, 2, 3) and all(bool((one[k] == k).all()) for k in range(n_units))
    else:
        ok = ok and one is None
    # uint8 sections (what tools/stack_restore.py gathers) keep their dtype and order
    u8 = shard.gather_sections(local.to(torch.uint8), n_units, dst=0)
    if rank == 0:
        ok = ok and u8.dtype == torch.uint8 and u8.shape == (n_units, 2, 3) and all(bool((u8[k] == k).all()) for k in range(n_units))
    else:
        ok = ok and u8 is None
    # SectionGatherer (bench.py's per-step output gather): on CPU / gloo it must fall back to the collective, twice in a
    # row (its device path alternates two buffers), and return the units in rank order on dst only
    g = shard.SectionGatherer((2, 3), torch.float32, 2, "cpu", dst=0)
    ok = ok and g.mode == "collective"
    for rep in range(2):
        got = g.gather(torch.stack([torch.full((2, 3), float(10 * rank + j + rep)) for j in range(2)]))
        if rank == 0:
            ok = ok and got.shape == (2 * world, 2, 3) and [float(got[i, 0, 0]) for i in range(4)] == [0. + rep, 1. + rep, 10. + rep, 11. + rep]
        else:
            ok = ok and got is None
    # restore_stack's host logic without a GPU is refused loudly (no CPU fallback)
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, ok, float(t)))
    dist.destroy_process_group()


def _run(n_units, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def test_gather_even_split():
    assert _run(6) == [(0, True, 2.0), (1, True, 2.0)]


def test_gather_ragged_split():
    assert _run(7) == [(0, True, 2.0), (1, True, 2.0)]


def test_gather_fewer_units_than_ranks():
    assert _run(1) == [(0, True, 2.0), (1, True, 2.0)]
