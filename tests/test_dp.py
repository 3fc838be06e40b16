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


def _worker(rank, world, port, outdir):
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
    torch.save((rank, params[0].grad.clone(), params[1].grad.clone(), bucket2 is bucket),
               os.path.join(outdir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_world2(outdir):
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, outdir)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    ok = all(p.exitcode == 0 for p in procs)
    for p in procs:
        if p.is_alive():
            p.kill()
    return ok


def test_bucketed_allreduce_world2(tmp_path):
    # results travel through files (no Queue teardown races); one retry covers a port grabbed in between
    ok = _run_world2(str(tmp_path)) or _run_world2(str(tmp_path))
    assert ok
    for rank in range(2):
        r, g0, g1, reused = torch.load(os.path.join(str(tmp_path), f"rank{rank}.pt"))
        assert r == rank
        assert torch.allclose(g0, torch.full((5, 3), 1.5))
        assert torch.allclose(g1, torch.arange(7.0) * 1.5)
        assert reused


def _bucket_worker(rank, world, port, outdir):
    from graphgps_b200.dp import GradBucket
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)

    class Toy(torch.nn.Module):   # parameter names of a GPSLayer: early group = FFN / norms / out-proj
        def __init__(self):
            super().__init__()
            self.local_model = torch.nn.Linear(3, 5)
            self.ff_linear1 = torch.nn.Linear(3, 6)
            self.norm2 = torch.nn.BatchNorm1d(3)

    layers = [Toy(), Toy()]
    bucket = GradBucket(layers)
    ptrs = [p.grad.data_ptr() for l in layers for p in l.parameters()]
    for l in layers:
        for p in l.parameters():
            p.grad.fill_(float(rank + 1))            # writes land in the flat buffer
    assert float(bucket.flat.sum()) == float(rank + 1) * sum(p.numel() for l in layers for p in l.parameters())
    early0 = bucket.segment(0, True)
    late0 = bucket.segment(0, False)
    bucket.allreduce(segments=[early0])              # partial collective first (overlap pattern) ...
    part = float(layers[0].ff_linear1.weight.grad[0, 0]), float(layers[0].local_model.weight.grad[0, 0])
    bucket.allreduce(segments=[late0, bucket.segment(1, True), bucket.segment(1, False)])   # ... then the rest
    same_ptrs = ptrs == [p.grad.data_ptr() for l in layers for p in l.parameters()]
    vals = [float(p.grad.flatten()[0]) for l in layers for p in l.parameters()]
    torch.save((rank, part, vals, same_ptrs, early0.numel(), late0.numel()), os.path.join(outdir, f"b{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_bucket_in_place_allreduce_world2(tmp_path):
    ctx = mp.get_context("spawn")

    def run():
        port = _free_port()
        procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=120)
        ok = all(p.exitcode == 0 for p in procs)
        for p in procs:
            if p.is_alive():
                p.kill()
        return ok

    assert run() or run()
    for rank in range(2):
        r, part, vals, same_ptrs, n_early, n_late = torch.load(os.path.join(str(tmp_path), f"b{rank}.pt"))
        assert part == (1.5, float(rank + 1))        # early segment reduced, late segment still local
        assert all(v == 1.5 for v in vals) and same_ptrs
        assert n_early >= 6 * 3 + 6 + 3 + 3 and n_late >= 5 * 3 + 5
