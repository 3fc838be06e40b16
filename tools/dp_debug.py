"""2-GPU bring-up of the data-parallel step (static bucket, overlapped all-reduce, NCCL inside a CUDA graph).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_debug.py"""
import faulthandler
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import graphgps_b200  # noqa: E402
from graphgps_b200.dp import GradBucket  # noqa: E402
from graphgps_b200.graph import graph_of  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
faulthandler.dump_traceback_later(70, exit=True)
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)


def say(*a):
    sys.stderr.write(f"[rank {rank} {time.strftime('%H:%M:%S')}] " + " ".join(str(x) for x in a) + "\n")
    sys.stderr.flush()


torch.manual_seed(0)
layer = graphgps_b200.GPSLayer(64, "CustomGatedGCN", "Transformer", 4).to(dev).train()
b = graphgps_b200.make_batch("zinc-gatedgcn", seed=rank, dim=64, num_graphs=16).to(dev)
graph_of(b)
ct = (torch.randn_like(b.x), torch.randn_like(b.edge_attr))
bucket = GradBucket([layer]).enable_overlap()
dist.all_reduce(torch.zeros(1, device=dev))
torch.cuda.synchronize()
say("init ok")


def step(mode):
    bb = graphgps_b200.GraphBatch(x=b.x.detach().requires_grad_(True), edge_index=b.edge_index,
                                  edge_attr=b.edge_attr.detach().requires_grad_(True), batch=b.batch, num_graphs=b.num_graphs)
    bb.__dict__["_gps_b200_graph"] = b.__dict__["_gps_b200_graph"]
    bucket.zero_()
    out = layer(bb)
    torch.autograd.backward([out.x, out.edge_attr], list(ct))
    if mode == "plain":
        bucket.allreduce()
    elif mode == "overlap":
        bucket.allreduce_overlapped()


def check(tag):
    torch.cuda.synchronize()
    g = bucket.flat.clone()
    ref = g.clone()
    dist.broadcast(ref, 0)
    say(tag, "bucket identical across ranks:", bool(torch.equal(g, ref)), "norm", float(g.norm()))


for mode in ("plain", "overlap"):
    step(mode)
    check("eager " + mode)

for mode in ("plain", "overlap"):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(mode)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    dist.barrier()
    say("capturing", mode)
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            step(mode)
        say("captured", mode)
        for _ in range(3):
            g.replay()
        check("graph " + mode)
    except Exception as e:  # noqa: BLE001
        say("capture failed", mode, repr(e)[:300])
        torch.cuda.synchronize()
dist.barrier()
say("done")
os._exit(0)   # destroy_process_group() blocks while graphs with captured NCCL kernels are alive
