"""The torch-only pieces of the drop-in layers against the UNMODIFIED reference classes, on the CPU.
Runs only where /root/reference exists (the build container); the GPU box skips it.  The CUDA-backed
pieces are covered by tests/test_gpu_parity.py against fixtures generated from the same classes."""
import numpy as np
import pytest
import torch

from graphneuralnetwork_b200 import layers
from oracle import ref_loader as R

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not present")
TOL = 1e-5


def _rel(a, b):
    return float(((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).detach())


def _copy_state(dst, src):
    dst.load_state_dict(src.state_dict(), strict=True)


@pytest.mark.parametrize("hidden_method", ["sum", "concat"])
@pytest.mark.parametrize("activation", ["relu", None])
def test_sagegcn_combination_matches_reference(hidden_method, activation):
    """SageGCN.py:23-36: self/neighbour products, sum | concat, optional activation.  Ours receives the
    already-aggregated neighbours (what the kernel returns); the reference aggregates its 3-D input itself."""
    ref = R.sage_pytorch()
    act = torch.nn.functional.relu if activation else None
    torch.manual_seed(0)
    theirs = ref["SageGCN"].SageGCN(24, 16, activation=act, aggr_neighbor_method="mean", aggr_hidden_method=hidden_method)
    ours = layers.SageGCN(24, 16, activation=act, aggr_neighbor_method="mean", aggr_hidden_method=hidden_method)
    _copy_state(ours, theirs)
    src, neigh = torch.randn(50, 24), torch.randn(50, 7, 24)
    assert _rel(ours(src, neigh.mean(dim=1)), theirs(src, neigh)) < TOL


def test_neighbor_aggregator_bias_matches_reference():
    ref = R.sage_pytorch()
    torch.manual_seed(1)
    theirs = ref["Aggregator"].NeighborAggregator(24, 16, use_bias=True, aggr_method="sum")
    theirs.bias.data.normal_()
    ours = layers.NeighborAggregator(24, 16, use_bias=True, aggr_method="sum")
    _copy_state(ours, theirs)
    neigh = torch.randn(40, 5, 24)
    assert _rel(ours(neigh.sum(dim=1)), theirs(neigh)) < TOL


@pytest.mark.parametrize("gcn", [False, True])
def test_sage_v2_layer_split_weight_matches_reference(gcn):
    """GraphSAGE/GraphSAGE.py:7-21: relu(W [self ‖ agg]) computed without the concat copy."""
    ref = R.sage_v2()
    torch.manual_seed(2)
    theirs = ref["GraphSAGE"].SageLayer(20, 12, gcn=gcn)
    ours = layers.SageLayer(20, 12, gcn=gcn)
    _copy_state(ours, theirs)
    a, b = torch.randn(33, 20), torch.randn(33, 20)
    assert _rel(ours(a, b), theirs(a, b)) < TOL


def test_semantic_attention_matches_reference():
    ref = R.han()
    torch.manual_seed(3)
    theirs = ref["SemanticAttention"].SemanticAttention(in_size=64, hidden_size=128)
    ours = layers.SemanticAttention(in_size=64, hidden_size=128)
    _copy_state(ours, theirs)
    z = torch.randn(300, 3, 64)
    assert _rel(ours(z), theirs(z)) < TOL
    z.requires_grad_(True)
    go = torch.randn(300, 64)
    g1, = torch.autograd.grad(ours(z), z, go)
    g2, = torch.autograd.grad(theirs(z), z, go)
    assert _rel(g1, g2) < TOL


def test_gatne_decoder_matches_reference():
    ref_pt, _ = R.gatne()
    torch.manual_seed(4)
    theirs = ref_pt.GraphDecoder(100, 32)
    ours = layers.GraphDecoder(100, 32)
    _copy_state(ours, theirs)
    emb, cn = torch.randn(16, 32), torch.randint(0, 100, (16, 6))
    assert _rel(ours(emb, cn), theirs(emb, cn)) < TOL


def test_model_constructors_consume_the_same_rng_stream():
    """Same seed => same initial weights as the reference constructors (parameter creation order and
    initialisers match), so a run seeded like the reference starts from the reference's weights."""
    gcn_mod, _ = R.gcn()
    for build_ref, build_ours in (
            (lambda: gcn_mod.GCN_Model(30, 16, 7, 3, 0.5), lambda: layers.GCN_Model(30, 16, 7, 3, 0.5)),
            (lambda: R.sage_pytorch()["GraphSage"].GraphSage(30, [16, 5], [4, 3]), lambda: layers.GraphSage(30, [16, 5], [4, 3])),
            (lambda: R.han()["HAN"].HANModel(2, 30, 8, 3, [4, 2], 0.1), lambda: layers.HANModel(2, 30, 8, 3, [4, 2], 0.1)),
            (lambda: R.gat_models()[0].GAT(30, 8, 7, 0.1, 0.2, 4), lambda: layers.GAT(30, 8, 7, 0.1, 0.2, 4)),
            (lambda: R.gat_models()[0].SpGAT(30, 8, 7, 0.1, 0.2, 4), lambda: layers.SpGAT(30, 8, 7, 0.1, 0.2, 4)),
            (lambda: R.sage_v2()["GraphSAGE"].GraphSAGE(2, 30, 16, Unsupervised=False, class_size=4),
             lambda: layers.GraphSAGE(2, 30, 16, Unsupervised=False, class_size=4)),
            (lambda: R.gatne()[0].GATNEModel(50, 16, 6, 2, 8, None), lambda: layers.GATNEModel(50, 16, 6, 2, 8, None)),
            (lambda: R.gatne()[1].GATNEModel(50, 16, 6, 2, 8, None), lambda: layers.GATNEModelV1(50, 16, 6, 2, 8, None)),
    ):
        torch.manual_seed(123)
        a = build_ref().state_dict()
        torch.manual_seed(123)
        b = build_ours().state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
