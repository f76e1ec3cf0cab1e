"""N>1 host path on CPU: two gloo ranks shard pages, run a stand-in "generate", rank 0 reassembles in page order."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n_pages, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dots_ocr_b200.sharding import shard_round_robin, gather_pages
    mine = shard_round_robin(n_pages, world, rank)
    results = [[i * 10 + k for k in range(3)] for i in mine]          # stand-in for generated ids of page i
    out = gather_pages(mine, results, n_pages)
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((out, float(t)))
    else:
        assert out is None
    dist.destroy_process_group()


def test_two_rank_page_sharding_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    n_pages = 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pages, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, tmax = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert out == [[i * 10 + k for k in range(3)] for i in range(n_pages)]
    assert tmax == 2.0


def test_cost_sharding_balances():
    from dots_ocr_b200.sharding import shard_by_cost, shard_round_robin
    costs = [19600, 5476, 5476, 5476, 19520, 1369, 5476, 5476]
    shards = shard_by_cost(costs, 2)
    assert sorted(i for s in shards for i in s) == list(range(8))
    loads = [sum(costs[i] for i in s) for s in shards]
    assert abs(loads[0] - loads[1]) <= 5476
    assert shard_round_robin(7, 2, 1) == [1, 3, 5]
