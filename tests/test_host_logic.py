"""Host-side logic that needs no GPU: the drop-in classes keep the reference's names, constructor
signatures and state_dict layout (checked against the parameter names stored in the golden
fixtures, which come from the unmodified reference), and the product path refuses CPU tensors."""
import inspect

import numpy as np
import pytest
import torch

from conftest import load_golden
from graphneuralnetwork_b200 import _lib, functional as Fn, layers
from graphneuralnetwork_b200.graph import CSRGraph, adj_cache


def _golden_param_shapes(g, prefix=""):
    return {k[len(prefix):]: v.shape for k, v in g.items()
            if k.startswith(prefix) and not k[len(prefix):].startswith("grad.") and ("grad." + k[len(prefix):] in
                                                                                      {x[len(prefix):] for x in g if x.startswith(prefix)})}


@pytest.mark.parametrize("fixture,prefix,build", [
    ("gcn_cora.npz", "", lambda: layers.GCN_Model(1433, 16, 7, 2, 0.5)),
    ("gat_small.npz", "dense.", lambda: layers.GAT(50, 8, 7, 0.0, 0.2, 8)),
    ("gat_small.npz", "sparse.", lambda: layers.SpGAT(50, 8, 7, 0.0, 0.2, 8)),
    ("sage_small.npz", "", lambda: layers.GraphSage(602, [128, 41], [5, 3])),
    ("sage_v2_small.npz", "", lambda: layers.GraphSAGE(2, 64, 32, gcn=False, agg_func='MEAN', Unsupervised=False, class_size=3)),
    ("han_small.npz", "", lambda: layers.HANModel(3, 40, 8, 3, [8], 0.0)),
])
def test_state_dict_matches_reference(fixture, prefix, build):
    g = load_golden(fixture)
    ref = _golden_param_shapes(g, prefix)
    ours = {k: tuple(v.shape) for k, v in build().state_dict().items()}
    assert ours == {k: tuple(v) for k, v in ref.items()}


def test_class_names_and_signatures():
    assert layers.Graph_conv_layer(4, 2)._get_name() == "Graph_conv_layer"  # GCN/GCN.py:23 dispatches on this
    sig = lambda c: list(inspect.signature(c.__init__).parameters)[1:]
    assert sig(layers.Graph_conv_layer) == ["in_features", "out_features", "is_bias", "kwargs"]
    assert sig(layers.GraphAttentionLayer) == ["in_features", "out_features", "dropout", "alpha", "concat", "kwargs"]
    assert sig(layers.SpGraphAttentionLayer) == ["in_features", "out_features", "dropout", "alpha", "concat"]
    assert sig(layers.GATConv) == ["feat_size", "hidden_size", "dropout", "num_heads", "alpha", "num_class", "kwargs"]
    assert sig(layers.NeighborAggregator) == ["input_dim", "output_dim", "use_bias", "aggr_method", "kwargs"]
    assert sig(layers.SageGCN)[:5] == ["input_dim", "hidden_dim", "activation", "aggr_neighbor_method", "aggr_hidden_method"]
    assert sig(layers.GraphSage) == ["input_dim", "hidden_dim", "num_neighbors_list"]
    assert list(inspect.signature(layers.GraphSAGE.forward).parameters)[1:5] == [
        "center_feats_data", "center_nodes_map", "center_neigh_feats_data", "center_neigh_nodes_map"]
    assert list(inspect.signature(layers.Aggregator).parameters) == ["neigh_feat", "agg_func"]


def test_cpu_tensors_are_refused_no_fallback(lib):
    with pytest.raises(_lib.GnnError):
        CSRGraph(torch.zeros(2, dtype=torch.int64), torch.zeros(0, dtype=torch.int32), None, 1, 1)
    with pytest.raises(_lib.GnnError):
        Fn.gather_reduce_raw(torch.zeros(4, 8), None, 2, 2, "mean")
    layer = layers.Graph_conv_layer(4, 2)
    adj = torch.sparse_coo_tensor(torch.tensor([[0, 1], [1, 0]]), torch.ones(2), (2, 2))
    with pytest.raises(_lib.GnnError):
        layer(torch.zeros(2, 4), adj)
    with pytest.raises(ValueError):
        Fn.gather_reduce_raw(torch.zeros(4, 8), None, 2, 2, "median")


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("GNN_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GnnError, match="no CPU fallback"):
        _lib.load()
    monkeypatch.delenv("GNN_B200_LIB")
    monkeypatch.setattr(_lib, "_lib", None)
    assert _lib.load().gnn_version() >= 100


def test_padded_views():
    out = Fn._padded_empty(5, 602, torch.float32, "cpu")
    assert out.shape == (5, 602) and out.stride(0) == 604 and out.stride(1) == 1
    out = Fn._padded_empty(5, 602, torch.bfloat16, "cpu")
    assert out.stride(0) == 608
    assert Fn._ld(out) == 608 and Fn._ld(out[:1]) == 602
