"""gnnfc — B200-native graph-filter hot path of soosiey/gnn-formation-control.

Import as ``import gnnfc`` (the repo-root shim maps that name onto this
directory, whose name is not a Python identifier).

    from gnnfc import GraphFilterBatch          # drop-in for utils/graphUtils/graphML.py:2369
    gf = GraphFilterBatch(G, F, K).cuda()
    gf.addGSO(S)                                 # S [B,E,N,N]
    y = gf(x)                                    # x [B,G,N] -> y [B,F,N]
"""
from . import _cabi
from ._cabi import GfcError, version, last_launch_count
from .gso import build_gso, build_csr, SparseGSO
from .graph_filter import GraphFilterBatch, GraphFilter, GraphFilterBatchGSO, graph_filter
from .recurrent import GraphFilterRNNBatch, GraphFilterMoRNNBatch, GraphFilterL2ShareBatch, torchpermul
from .rollout import Rollout
from .data import (robot_major_to_batch, graphs_from_recording, gso_batch_from_recording, positions_from_recording,
                   RecordingLoader)
from .dp import GradBucket, BucketedReducer, PeerExchange, shard_range, broadcast_parameters

__all__ = ["GraphFilterBatch", "GraphFilter", "GraphFilterBatchGSO", "graph_filter", "build_gso", "build_csr", "SparseGSO",
           "GraphFilterRNNBatch", "GraphFilterMoRNNBatch", "GraphFilterL2ShareBatch", "torchpermul", "Rollout",
           "robot_major_to_batch", "graphs_from_recording", "gso_batch_from_recording", "positions_from_recording",
           "RecordingLoader", "GradBucket", "BucketedReducer", "PeerExchange", "shard_range", "broadcast_parameters", "GfcError", "version",
           "last_launch_count"]
