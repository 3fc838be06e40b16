"""Input pipeline for the GPSLayer hot path: pre-collated pinned host batches, copied to the
device on their own stream a configurable number of steps ahead, with the graph structure (CSR by destination, CSC by
source, graph offsets) built on arrival - so the compute stream never waits for a host->device copy and no layer ever
issues the device->host sync that `to_dense_batch` needs in the reference (gps_layer.py:199; the published runs used
`num_workers: 0`, i.e. the loader was on the critical path).

    feeder = BatchPrefetcher(batches, device, depth=2)      # batches: iterable of GraphBatch-like host objects
    for b in feeder:                                         # b lives on the device, structure already built
        out = model(b)

The host objects may be anything exposing x / edge_index / edge_attr / batch (+ num_graphs or ptr): a collated PyG
`Batch`, or `graphgps_b200.GraphBatch`.  Each is pinned once (`pin_memory()` when available) and re-used across epochs.
"""
from __future__ import annotations

import collections

import torch

from .batch import GraphBatch
from .graph import graph_of

_FIELDS = ("x", "edge_index", "edge_attr", "batch")


def collate(graphs, dim=None):
    """Collate per-graph (x [n,d], edge_index [2,e], edge_attr [e,d]) triples into one host GraphBatch the way PyG does:
    node-offset edge indices, sorted `batch` vector, `ptr` offsets, `num_graphs` (no device work, no syncs later)."""
    xs, eis, eas, bs, ptr, off = [], [], [], [], [0], 0
    for g, (x, ei, ea) in enumerate(graphs):
        n = int(x.shape[0])
        xs.append(x)
        eis.append(ei + off)
        if ea is not None:
            eas.append(ea)
        bs.append(torch.full((n,), g, dtype=torch.int64))
        off += n
        ptr.append(off)
    d = dim if dim is not None else (int(xs[0].shape[1]) if xs else 0)
    x = torch.cat(xs) if xs else torch.zeros(0, d)
    ei = torch.cat(eis, dim=1) if eis else torch.zeros(2, 0, dtype=torch.int64)
    ea = torch.cat(eas) if eas else torch.zeros(ei.shape[1], d)
    b = torch.cat(bs) if bs else torch.zeros(0, dtype=torch.int64)
    return GraphBatch(x=x, edge_index=ei, edge_attr=ea, batch=b, num_graphs=len(ptr) - 1,
                      ptr=torch.tensor(ptr, dtype=torch.int64))


def _pin(t):
    return t if (not torch.is_tensor(t) or t.is_pinned() or not torch.cuda.is_available()) else t.pin_memory()


class BatchPrefetcher:
    """Iterates device-resident batches; `depth` batches are in flight on a dedicated copy stream.

    Every yielded batch owns fresh device tensors (safe to keep for the backward pass); its graph structure has been
    built on the copy stream and is cached on the object, and the consumer's current stream is made to wait for the
    copy + build of exactly that batch (event), nothing more."""

    def __init__(self, host_batches, device, depth=2, build_structure=True):
        self.device = torch.device(device)
        self.depth = max(1, int(depth))
        self.build_structure = build_structure
        self.host = []
        for hb in host_batches:
            rec = {k: _pin(getattr(hb, k)) for k in _FIELDS if getattr(hb, k, None) is not None}
            ng = getattr(hb, "num_graphs", None)
            if ng is None and getattr(hb, "ptr", None) is not None:
                ng = int(hb.ptr.shape[0]) - 1
            if ng is None:
                ng = int(rec["batch"][-1]) + 1 if rec["batch"].numel() else 0     # host tensor: no device sync
            rec["num_graphs"] = int(ng)
            self.host.append(rec)
        self.copy_stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def __len__(self):
        return len(self.host)

    def _issue(self, rec):
        cs = self.copy_stream
        with torch.cuda.stream(cs):
            dev = {k: rec[k].to(self.device, non_blocking=True) for k in _FIELDS if k in rec}
            b = GraphBatch(x=dev["x"], edge_index=dev["edge_index"], edge_attr=dev.get("edge_attr"), batch=dev["batch"],
                           num_graphs=rec["num_graphs"])
            if self.build_structure:
                graph_of(b)
            ev = torch.cuda.Event()
            ev.record(cs)
        return b, ev

    def __iter__(self):
        if self.copy_stream is None:
            raise RuntimeError("BatchPrefetcher feeds a CUDA device (graphgps_b200 has no CPU path)")
        queue = collections.deque()
        it = iter(self.host)
        for rec in it:
            queue.append(self._issue(rec))
            if len(queue) >= self.depth:
                break
        while queue:
            b, ev = queue.popleft()
            nxt = next(it, None)
            if nxt is not None:
                queue.append(self._issue(nxt))
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            held = [b.x, b.edge_index, b.edge_attr, b.batch]
            gs = b.__dict__.get("_gps_b200_graph")
            if gs is not None:
                held += [gs.storage, gs.edge_index, gs.batch]
            for t in held:
                if t is not None:
                    t.record_stream(cur)          # allocated on the copy stream, consumed on the compute stream
            yield b
