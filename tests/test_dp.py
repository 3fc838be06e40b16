"""CPU, world_size 2 over gloo: the data-parallel host logic (graph sharding + bucketed gradient
all-reduce) used by bench.py --gpus N."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphgps_b200.dp import allreduce_gradients, shard_graph_range


def test_shard_graph_range_partitions():
    for n in (256, 7, 1, 0):
        for w in (1, 2, 3, 8):
            spans = [shard_graph_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)),
              torch.nn.Parameter(torch.zeros(2, 2))]
    params[0].grad = torch.full((5, 3), float(rank + 1))
    params[1].grad = torch.arange(7.0) * (rank + 1)
    params[2].grad = None                       # unused parameter: skipped
    bucket = allreduce_gradients(params)
    bucket2 = allreduce_gradients(params, bucket)   # bucket reuse, values already equal -> unchanged
    q.put((rank, params[0].grad.clone(), params[1].grad.clone(), bucket2 is bucket))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g0, g1, reused in res:
        assert torch.allclose(g0, torch.full((5, 3), 1.5))
        assert torch.allclose(g1, torch.arange(7.0) * 1.5)
        assert reused
