from .gcn import GCN_Model, Graph_conv_layer
from .gat import (GAT, GATBase, GraphAttentionLayer, SpecialSpmm, SpecialSpmmFunction, SpGAT,
                  SpGraphAttentionLayer)
from .han import GATConv, HANLayer, HANModel, SemanticAttention
from .sage import (Aggregator, CapturedGraphSage, GraphSage, NeighborAggregator, SageGCN, SampledBlock,
                   gather_mean)

from .sage_v2 import GraphSAGE, SageLayer
from .gatne import GATNEModel, GATNEModelV1, GraphDecoder, GraphEncoder
from .gtn import GTConv, GTLayer, GTN_Model

__all__ = [
    "GTConv", "GTLayer", "GTN_Model", "GraphSAGE", "SageLayer", "GATNEModel", "GATNEModelV1", "GraphDecoder", "GraphEncoder",
    "GCN_Model", "Graph_conv_layer", "SpecialSpmm", "SpecialSpmmFunction", "GAT", "GATBase", "GraphAttentionLayer", "SpGAT", "SpGraphAttentionLayer",
    "GATConv", "HANLayer", "HANModel", "SemanticAttention", "Aggregator", "GraphSage", "NeighborAggregator",
    "SageGCN", "SampledBlock", "gather_mean", "CapturedGraphSage",
]
