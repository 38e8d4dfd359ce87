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


@pytest.mark.parametrize("tag,build", [
    ("pt_t_sum", lambda f: layers.GraphEncoder(300, 32, 10, 3, 20, None)),
    ("pt_i_sum", lambda f: layers.GraphEncoder(300, 32, 10, 3, 20, f)),
    ("v1_t", lambda f: layers.GATNEModelV1(300, 32, 10, 3, 20, None)),
    ("v1_i", lambda f: layers.GATNEModelV1(300, 32, 10, 3, 20, f)),
])
def test_gatne_state_dict_matches_reference(tag, build):
    g = load_golden("gatne_small.npz")
    ref = _golden_param_shapes(g, tag + ".")
    ours = {k: tuple(v.shape) for k, v in build(torch.from_numpy(g["features"])).state_dict().items()}
    assert ours == {k: tuple(v) for k, v in ref.items()}
    full = layers.GATNEModel(300, 32, 10, 3, 20, None)  # GATNE_Pytorch/models/GATNE.py:117-128
    assert set(full.state_dict()) == {"encoder." + k for k in _golden_param_shapes(g, "pt_t_sum.")} | {"decoder.weights"}


@pytest.mark.parametrize("num_layers", [1, 2, 3, 4])
def test_gcn_model_layout_matches_reference(num_layers):
    """Module names, order and widths of GCN/GCN.py:5-19 for every depth (the fused-ReLU walk of our forward
    relies on the Graph_conv_layer -> ReLU -> Dropout order)."""
    m = layers.GCN_Model(30, 16, 7, num_layers, 0.5)
    names = [n for n, _ in m.gcn_blocks.named_children()]
    expect, widths = [], []
    for i in range(num_layers):
        if i == 0:
            expect += ["gcn0", "relu0", "dropout0"]
            widths.append((30, 16))
        elif i == num_layers - 1:
            expect += [f"gcn{i}"]
            widths.append((16, 7))
        else:
            expect += [f"gcn{i}", f"relu{i}", f"dropout{i}"]
            widths.append((16, 16))
    assert names == expect
    got = [(l.in_features, l.out_features) for l in m.gcn_blocks if l._get_name() == "Graph_conv_layer"]
    assert got == widths
    assert set(m.state_dict()) == {f"gcn_blocks.gcn{i}.{p}" for i in range(num_layers) for p in ("dense.weight", "bias")}
    assert "bias" not in dict(layers.Graph_conv_layer(4, 2, is_bias=False).named_parameters())


def test_runtime_and_new_entry_points_refuse_cpu():
    from graphneuralnetwork_b200 import runtime
    if not torch.cuda.is_available():
        with pytest.raises(_lib.GnnError):
            runtime.CapturedTrainStep(torch.nn.Linear(2, 2), lambda: torch.zeros(()))
        with pytest.raises(_lib.GnnError):
            runtime.CapturedForward(lambda: torch.zeros(()))
    with pytest.raises(_lib.GnnError):
        Fn.typed_gather_reduce(torch.zeros(4, 2, 3), torch.zeros(2, 2, 2, dtype=torch.int64))
    with pytest.raises(ValueError, match="please choice else aggregator"):
        Fn.typed_gather_reduce(torch.zeros(4, 2, 3), torch.zeros(2, 2, 2, dtype=torch.int64), "max")
    with pytest.raises(_lib.GnnError):
        Fn.sddmm(torch.zeros(3, dtype=torch.int64), torch.zeros(3, dtype=torch.int64), torch.zeros(2, 4), torch.zeros(2, 4))
    with pytest.raises(_lib.GnnError):
        layers.SpecialSpmm()(torch.zeros(2, 3, dtype=torch.int64), torch.zeros(3), torch.Size([2, 2]), torch.zeros(2, 4))


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
    assert sig(layers.GraphEncoder) == ["num_nodes", "embedding_size", "embedding_u_size", "edge_type_count",
                                        "attention_size", "features", "agg_func", "kwargs"]
    assert sig(layers.GATNEModelV1) == ["num_nodes", "embedding_size", "embedding_u_size", "edge_type_count", "dim_a",
                                        "features", "kwargs"]
    assert list(inspect.signature(layers.GATNEModel.forward).parameters)[1:] == ["inputs", "node_types", "node_neigh",
                                                                                "context_negative"]
    assert list(inspect.signature(layers.SpecialSpmm.forward).parameters)[1:] == ["indices", "values", "shape", "b"]
    assert list(inspect.signature(layers.SageLayer.__init__).parameters)[1:4] == ["input_size", "output_size", "gcn"]


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


def test_sage_weight_planes_follow_parameter_versions():
    """GraphSage._weight_planes (host logic, device-agnostic): the fp16 hi/lo planes of [W_self ; W_agg] are rebuilt
    only when a parameter's version or storage changes, in place (a captured graph keeps reading the same buffers),
    and hi + lo reproduces the fp32 weights to 2^-21."""
    torch.manual_seed(0)
    model = layers.GraphSage(10, [6, 3], [4, 2])
    ld = 12
    hi, lo = model._weight_planes(ld, "cpu")
    l0 = model.gcn[0]
    Wc = torch.zeros(2 * ld, 6)
    Wc[:10], Wc[ld:ld + 10] = l0.weight.detach(), l0.aggregator.weight.detach()
    assert hi.dtype == lo.dtype == torch.float16 and tuple(hi.shape) == (2 * ld, 6)
    assert float((hi.float() + lo.float() - Wc).abs().max()) <= 2.0 ** -21 * float(Wc.abs().max())
    assert float(hi[10:12].abs().max()) == 0.0 and float(hi[22:].abs().max()) == 0.0  # pad rows meet zero weights
    ver = model._l0_planes["ver"]
    hi2, lo2 = model._weight_planes(ld, "cpu")
    assert hi2 is hi and lo2 is lo and model._l0_planes["ver"] == ver  # unchanged weights: nothing recomputed
    with torch.no_grad():
        l0.weight.mul_(2.0)
    hi3, lo3 = model._weight_planes(ld, "cpu")
    assert hi3 is hi and model._l0_planes["ver"] != ver               # same buffers, new contents
    assert torch.allclose(hi3[:10].float() + lo3[:10].float(), l0.weight.detach(), rtol=0, atol=1e-6)
