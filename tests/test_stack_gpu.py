"""GPU: GPSStack - the plane hand-off between consecutive layers, the persistent weight planes and the
captured whole-stack step must not change results: everything is compared with the same layers run one by one with the
hand-off switched off (each layer converting its own inputs and weights), and after an in-place weight update."""
import pytest
import torch

import graphgps_b200
from graphgps_b200.batch import make_batch
from graphgps_b200.graph import graph_of
from util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(stack, b, ct_x, ct_e, handoff):
    for l in stack.layers:
        l.__dict__["plane_handoff"] = handoff
        l.__dict__.pop("_wplanes", None)
    bb = b.clone()
    bb.x.requires_grad_(True)
    bb.edge_attr.requires_grad_(True)
    x_in, e_in = bb.x, bb.edge_attr
    for p in stack.parameters():
        p.grad = None
    out = stack(bb)
    torch.autograd.backward([out.x, out.edge_attr], [ct_x, ct_e])
    return (out.x.detach().clone(), out.edge_attr.detach().clone(), x_in.grad.clone(), e_in.grad.clone(),
            [p.grad.clone() for p in stack.parameters()])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_stack_handoff_and_weight_plane_cache_do_not_change_results(precision):
    torch.manual_seed(0)
    stack = graphgps_b200.GPSStack(3, 64, "CustomGatedGCN", "Transformer", 4, precision=precision).to(DEV).train()
    b = make_batch("zinc-gatedgcn", seed=3, dim=64, num_graphs=16).to(DEV)
    ct_x, ct_e = torch.randn_like(b.x), torch.randn_like(b.edge_attr)
    ref = _run(stack, b, ct_x, ct_e, handoff=False)
    got = _run(stack, b, ct_x, ct_e, handoff=True)
    again = _run_keep(stack, b, ct_x, ct_e)           # second call: weight planes are reused (wplanes_valid)
    for r, g, a in zip(ref[:4], got[:4], again[:4]):      # same kernels on the same plane values; only the order of the
        assert rel_err(g.cpu(), r.cpu()) < 1e-6 and rel_err(a.cpu(), r.cpu()) < 1e-6   # BatchNorm-sum atomics differs
    for r, g in zip(ref[4], got[4]):
        assert rel_err(g.cpu(), r.cpu()) < 2e-5          # split-K atomics: order-dependent rounding only
    # an optimiser-style in-place update must invalidate the cached weight planes
    with torch.no_grad():
        for p in stack.parameters():
            p.add_(0.01 * torch.randn_like(p))
    new_ref = _run(stack, b, ct_x, ct_e, handoff=False)
    new_got = _run_keep(stack, b, ct_x, ct_e)
    assert rel_err(new_ref[0].cpu(), ref[0].cpu()) > 1e-4
    assert rel_err(new_got[0].cpu(), new_ref[0].cpu()) < 1e-6 and rel_err(new_got[2].cpu(), new_ref[2].cpu()) < 1e-6


def _run_keep(stack, b, ct_x, ct_e):
    for l in stack.layers:
        l.__dict__["plane_handoff"] = True
    bb = b.clone()
    bb.x.requires_grad_(True)
    bb.edge_attr.requires_grad_(True)
    x_in, e_in = bb.x, bb.edge_attr
    for p in stack.parameters():
        p.grad = None
    out = stack(bb)
    torch.autograd.backward([out.x, out.edge_attr], [ct_x, ct_e])
    return (out.x.detach().clone(), out.edge_attr.detach().clone(), x_in.grad.clone(), e_in.grad.clone(),
            [p.grad.clone() for p in stack.parameters()])


def test_stack_captured_step_matches_eager_with_bucket():
    torch.manual_seed(1)
    stack = graphgps_b200.GPSStack(2, 64, "CustomGatedGCN", "Transformer", 4).to(DEV).train()
    b = make_batch("zinc-gatedgcn", seed=5, dim=64, num_graphs=12).to(DEV)
    graph_of(b)
    ct_x, ct_e = torch.randn_like(b.x), torch.randn_like(b.edge_attr)
    eager = _run_keep(stack, b, ct_x, ct_e)
    bucket = stack.make_grad_bucket()
    step = stack.capture(b, ct_x, ct_e, bucket=bucket)
    for _ in range(2):
        step.replay()
    torch.cuda.synchronize()
    assert rel_err(step.x_out.cpu(), eager[0].cpu()) < 1e-6 and rel_err(step.grad_x.cpu(), eager[2].cpu()) < 1e-6
    for p, g in zip(stack.parameters(), eager[4]):
        assert rel_err(p.grad.cpu(), g.cpu()) < 1e-5
    # new input data through the static buffers, weights changed in place: the replay follows both
    with torch.no_grad():
        step.x_in.copy_(b.x * 0.5)
        for p in stack.parameters():
            p.mul_(1.01)
    step.replay()
    b2 = b.clone()
    b2.x = step.x_in.detach().clone()      # x_in shares storage with b.x: it already holds the halved values
    for l in stack.layers:
        l.__dict__.pop("_grad_bucket", None)
    eager2 = _run_keep(stack, b2, ct_x, ct_e)
    torch.cuda.synchronize()
    assert rel_err(step.x_out.cpu(), eager2[0].cpu()) < 1e-6


def test_batch_prefetcher_feeds_the_layer_without_changing_results():
    """pinned host batches copied on their own stream `depth` steps ahead, graph structure built on arrival;
    the layer's outputs must equal the plain `.to(device)` path and every batch must arrive with its structure cached."""
    from graphgps_b200.loader import BatchPrefetcher
    torch.manual_seed(3)
    layer = graphgps_b200.GPSLayer(64, "CustomGatedGCN", "Transformer", 4).to(DEV).eval()
    host = [make_batch("zinc-gatedgcn", seed=20 + i, dim=64, num_graphs=8 + i) for i in range(5)]
    with torch.no_grad():
        ref = [layer(b.clone().to(DEV)).x.clone() for b in host]
        got = []
        for b in BatchPrefetcher(host, DEV, depth=2):
            assert "_gps_b200_graph" in b.__dict__ and b.x.is_cuda
            got.append(layer(b).x.clone())
    torch.cuda.synchronize()
    assert len(got) == len(ref)
    for a, r in zip(got, ref):
        assert rel_err(a.cpu(), r.cpu()) < 1e-6
