"""graphgps_b200 — B200 (sm_100a) implementation of the GraphGPS `GPSLayer` hot path.

Public surface (mirrors the reference's module boundary):
    GPSLayer      drop-in for graphgps.layer.gps_layer.GPSLayer
    GraphBatch    duck-typed stand-in for a collated PyG Batch (PyG is optional)
    make_batch    seeded synthetic batches of the BASELINE shapes
"""
from .batch import GraphBatch, SHAPES, make_batch, batch_from_lists  # noqa: F401
from .gps_layer import GPSLayer  # noqa: F401

__all__ = ["GPSLayer", "GraphBatch", "SHAPES", "make_batch", "batch_from_lists"]
