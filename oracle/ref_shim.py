"""TEST INFRASTRUCTURE — loads the reference's own hot-path layer files *verbatim* on CPU.

The reference (rampasek/GraphGPS) cannot be imported as a package here because
torch_geometric / torch_scatter / performer_pytorch / yacs are not installed (SURVEY.md
section 8c).  Its four hot-path files do run unmodified once the missing third-party *leaf*
symbols are stubbed.  This module installs those stubs into
``sys.modules`` and loads, by path,

    graphgps/layer/performer_layer.py   (in-repo copy of performer_pytorch)
    graphgps/layer/gatedgcn_layer.py
    graphgps/layer/gine_conv_layer.py
    graphgps/layer/gps_layer.py

from ``/root/reference`` (authoring container) or from ``oracle/_ref/`` (git-ignored copies made by
``oracle/build_ref.py`` so they travel to the GPU box).  Nothing from the reference is copied into
tracked files.  Only tests/, bench.py's reference/cpu_baseline legs and __graft_entry__.smoke() may
import this module; the product package never does.

Stub semantics follow SURVEY.md Appendix A:
  * torch_scatter.scatter(src, index, 0, None, dim_size, 'sum') == zeros.index_add_
    (call sites gatedgcn_layer.py:118-123)
  * MessagePassing.propagate: `<name>_j` = kw[name][edge_index[0]], `<name>_i` = kw[name][edge_index[1]],
    aggregation index = edge_index[1], dim_size = N  (PyG default flow source_to_target)
  * GINEConv(nn, eps=0): out = nn((1+eps) x_i + sum_j relu(x_j + e_ij)); eps is a buffer
  * to_dense_batch: zero padded [B, Nmax, d] + bool mask
  * GCNConv(in, out) (PyG 2.2 defaults: improved=False, add_self_loops=True, normalize=True, bias=True):
    x' = lin(x) (Linear without bias, glorot); gcn_norm = add_remaining_self_loops (unit weights; an existing
    self-loop edge is replaced by the single loop of weight 1), deg = scatter_add(w, col), norm = deg^-1/2[row]
    * w * deg^-1/2[col]; out = sum_{j->i} norm_ij x'_j + bias.  PyG's source is NOT under /root/reference
    (third party, README.md:15,25 pins pyg=2.2): this variant is pinned only to this restatement of the
    published algorithm, cross-checked against the independent dense restatement in gps_oracle.OracleGCN.
"""
from __future__ import annotations

import importlib.util
import inspect
import os
import sys
import types

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
_FILES = ["performer_layer.py", "gatedgcn_layer.py", "gine_conv_layer.py", "gps_layer.py"]


def find_reference_layer_dir():
    """Directory holding the four reference layer files, or None."""
    for cand in (os.path.join(_HERE, "_ref"), "/root/reference/graphgps/layer"):
        if all(os.path.isfile(os.path.join(cand, f)) for f in _FILES):
            return cand
    return None


# ----------------------------------------------------------------------------- stubs
def _scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0 and out is None and reduce == "sum"
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add_(0, index, src)


class _PygLinear(nn.Linear):
    def __init__(self, in_channels, out_channels, bias=True, **kw):
        super().__init__(in_channels, out_channels, bias=bias)


class _MessagePassing(nn.Module):
    def __init__(self, aggr="add", **kw):
        super().__init__()
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        src, dst = edge_index[0], edge_index[1]

        def dim_size():
            for v in kwargs.values():
                if torch.is_tensor(v):
                    return v.shape[0]
                if isinstance(v, (tuple, list)) and torch.is_tensor(v[0]):
                    return v[1].shape[0] if v[1] is not None else v[0].shape[0]
            raise RuntimeError("cannot infer number of nodes")

        N = dim_size()

        def collect(fn, extra):
            out = {}
            for name in inspect.signature(fn).parameters:
                if name in extra:
                    out[name] = extra[name]
                elif name.endswith("_j") or name.endswith("_i"):
                    base = kwargs.get(name[:-2])
                    sel = 0 if name.endswith("_j") else 1
                    if base is None:
                        out[name] = None
                    else:
                        if isinstance(base, (tuple, list)):
                            base = base[sel]
                        out[name] = base[src if sel == 0 else dst]
                elif name in kwargs:
                    out[name] = kwargs[name]
            return out

        reserved = {"index": dst, "dim_size": N, "ptr": None, "edge_index": edge_index}
        msg = self.message(**collect(self.message, reserved))
        if type(self).aggregate is not _MessagePassing.aggregate:
            agg_kw = collect(self.aggregate, reserved)
            agg_kw.pop(next(iter(inspect.signature(self.aggregate).parameters)), None)
            aggr_out = self.aggregate(msg, **agg_kw)
        else:
            aggr_out = _scatter(msg, dst, 0, None, N, "sum")
        if type(self).update is not _MessagePassing.update:
            upd_kw = collect(self.update, reserved)
            upd_kw.pop(next(iter(inspect.signature(self.update).parameters)), None)
            return self.update(aggr_out, **upd_kw)
        return aggr_out

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, dim_size=None):  # pragma: no cover - default path above
        return _scatter(inputs, index, 0, None, dim_size, "sum")

    def update(self, inputs):
        return inputs


class _GINEConv(_MessagePassing):
    """PyG 2.2 GINEConv(nn, eps=0., train_eps=False, edge_dim=None)."""

    def __init__(self, nn_module, eps=0.0, train_eps=False, edge_dim=None, **kw):
        super().__init__(aggr="add")
        self.nn = nn_module
        assert edge_dim is None
        if train_eps:
            self.eps = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer("eps", torch.Tensor([eps]))

    def forward(self, x, edge_index, edge_attr=None, size=None):
        out = self.propagate(edge_index, x=(x, x), edge_attr=edge_attr)
        out = out + (1 + self.eps) * x
        return self.nn(out)

    def message(self, x_j, edge_attr):
        return (x_j + edge_attr).relu()


class _GCNConv(_MessagePassing):
    """PyG 2.2 GCNConv with default arguments (call site gps_layer.py:49-51, :186)."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False, add_self_loops=True,
                 normalize=True, bias=True, **kw):
        super().__init__(aggr="add")
        assert not improved and not cached and add_self_loops and normalize and bias
        self.lin = _PygLinear(in_channels, out_channels, bias=False)
        nn.init.xavier_uniform_(self.lin.weight)                 # weight_initializer='glorot'
        self.bias = nn.Parameter(torch.zeros(out_channels))      # inits.zeros

    @staticmethod
    def gcn_norm(edge_index, num_nodes, dtype):
        row, col = edge_index[0], edge_index[1]
        w = torch.ones(edge_index.shape[1], dtype=dtype)
        keep = row != col                                        # add_remaining_self_loops(fill_value=1.)
        loop_w = torch.ones(num_nodes, dtype=dtype)
        loop_w[row[~keep]] = w[~keep]
        loops = torch.arange(num_nodes, dtype=edge_index.dtype)
        edge_index = torch.cat([edge_index[:, keep], torch.stack([loops, loops])], dim=1)
        w = torch.cat([w[keep], loop_w])
        row, col = edge_index[0], edge_index[1]
        deg = _scatter(w, col, 0, None, num_nodes, "sum")
        dis = deg.pow(-0.5)
        dis = dis.masked_fill(dis == float("inf"), 0)
        return edge_index, dis[row] * w * dis[col]

    def forward(self, x, edge_index, edge_weight=None):
        assert edge_weight is None
        edge_index, edge_weight = self.gcn_norm(edge_index, x.shape[0], x.dtype)
        x = self.lin(x)
        out = self.propagate(edge_index, x=x, edge_weight=edge_weight, size=None)
        return out + self.bias

    def message(self, x_j, edge_weight):
        return edge_weight.view(-1, 1) * x_j


def _to_dense_batch(x, batch):
    B = int(batch.max()) + 1 if batch.numel() else 0
    n = torch.bincount(batch, minlength=B)
    ptr = torch.zeros(B + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(n, 0)
    nmax = int(n.max()) if B else 0
    idx = torch.arange(x.shape[0]) - ptr[batch] + batch * nmax
    out = x.new_zeros((B * nmax,) + tuple(x.shape[1:]))
    out[idx] = x
    mask = torch.zeros(B * nmax, dtype=torch.bool)
    mask[idx] = True
    return out.view(B, nmax, *x.shape[1:]), mask.view(B, nmax)


class _Batch:
    def __init__(self, batch=None, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package
    sys.modules[name] = m
    return m


_LOADED = {}


def load_reference(layer_dir=None):
    """Returns a namespace with the reference's GPSLayer, GatedGCNLayer, SelfAttention classes."""
    layer_dir = layer_dir or find_reference_layer_dir()
    if layer_dir is None:
        raise FileNotFoundError("reference layer files not found (neither oracle/_ref nor /root/reference)")
    if layer_dir in _LOADED:
        return _LOADED[layer_dir]

    act_dict = {"relu": nn.ReLU, "gelu": nn.GELU, "swish": nn.SiLU,
                "lrelu_03": lambda: nn.LeakyReLU(0.3)}
    _mod("torch_scatter", scatter=_scatter)
    register = _mod("torch_geometric.graphgym.register", act_dict=act_dict,
                    register_layer=lambda name: (lambda cls: cls))
    layer_m = _mod("torch_geometric.graphgym.models.layer", LayerConfig=object)
    models_m = _mod("torch_geometric.graphgym.models", layer=layer_m)
    graphgym = _mod("torch_geometric.graphgym", register=register, models=models_m)
    conv = _mod("torch_geometric.nn.conv", MessagePassing=_MessagePassing)
    norm = _mod("torch_geometric.nn.norm", LayerNorm=None)
    inits = _mod("torch_geometric.nn.inits", reset=lambda m: None)
    pygnn = _mod("torch_geometric.nn", Linear=_PygLinear, GINEConv=_GINEConv, GCNConv=_GCNConv,
                 GINConv=None, GENConv=None, GATConv=None, PNAConv=None, conv=conv, norm=norm,
                 inits=inits)
    data = _mod("torch_geometric.data", Batch=_Batch)
    utils = _mod("torch_geometric.utils", to_dense_batch=_to_dense_batch)
    _mod("torch_geometric", nn=pygnn, graphgym=graphgym, data=data, utils=utils)
    _mod("performer_pytorch.reversible", ReversibleSequence=None, SequentialSequence=None)
    _mod("local_attention", LocalAttention=None)
    _mod("axial_positional_embedding", AxialPositionalEmbedding=None)
    perf_pkg = _mod("performer_pytorch")
    _mod("graphgps")
    _mod("graphgps.layer")
    _mod("graphgps.layer.bigbird_layer", SingleBigBirdLayer=None)

    def load(fname):
        name = "graphgps.layer." + fname[:-3]
        spec = importlib.util.spec_from_file_location(name, os.path.join(layer_dir, fname))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    perf = load("performer_layer.py")
    perf_pkg.SelfAttention = perf.SelfAttention
    gated = load("gatedgcn_layer.py")
    load("gine_conv_layer.py")
    gps = load("gps_layer.py")
    ns = types.SimpleNamespace(GPSLayer=gps.GPSLayer, GatedGCNLayer=gated.GatedGCNLayer,
                               SelfAttention=perf.SelfAttention, performer=perf,
                               layer_dir=layer_dir)
    _LOADED[layer_dir] = ns
    return ns
