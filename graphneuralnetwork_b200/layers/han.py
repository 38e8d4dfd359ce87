"""Drop-in HAN node-level attention, semantic attention and model (class names, ctor
signatures, parameter names and forward signatures of /root/reference
HAN/models/NodeAttention.py, SemanticAttention.py and HAN.py).

Each metapath's `GATConv` (HAN/models/NodeAttention.py:44-62) runs its 8 heads in one fused
launch over the CSR of that metapath's float64 0/1 adjacency (`adj > 0`); the second
`F.elu` the reference applies when `num_class is None` (NodeAttention.py:62) is kept and
fused as a double ELU.  `SemanticAttention` is a small dense reduce and stays torch.
"""
import torch
from torch import nn
import torch.nn.functional as F

from .. import _lib
from ..graph import adj_cache
from .gat import GraphAttentionLayer, _fused_heads


class GATConv(nn.Module):
    def __init__(self, feat_size, hidden_size, dropout, num_heads, alpha=0.2, num_class=None, **kwargs):
        super(GATConv, self).__init__(**kwargs)
        self.dropout = dropout
        self.attentions = nn.ModuleList()
        for i in range(num_heads):
            self.attentions.add_module(f'AttentionHead{i}',
                                       GraphAttentionLayer(feat_size, hidden_size, dropout=dropout, alpha=alpha,
                                                           concat=True))
        self.num_class = num_class
        if self.num_class is not None:
            self.out_att = GraphAttentionLayer(hidden_size * num_heads, num_class, dropout=dropout, alpha=alpha,
                                               concat=False)

    def forward(self, x, adj):
        adj = adj_cache.get(adj)
        x = F.dropout(x, self.dropout, training=self.training)
        atts = list(self.attentions)
        halves = [att._halves() for att in atts]
        # heads are concat=True (ELU, NodeAttention.py:35); with num_class None the reference
        # applies F.elu once more to the concatenation (NodeAttention.py:62) => elu=2.  The
        # F.dropout between them is the identity in eval mode and is applied by torch in training.
        fuse_second = self.num_class is None and not (self.training and self.dropout > 0.0)
        x = _fused_heads(x, adj, [att.W for att in atts], [h[0] for h in halves], [h[1] for h in halves],
                         atts[0].alpha, _lib.GAT_SOFTMAX, 2 if fuse_second else 1, self.dropout, self.training)
        if fuse_second:
            return x
        x = F.dropout(x, self.dropout, training=self.training)
        return F.elu(self.out_att(x, adj)) if self.num_class is not None else F.elu(x)


class SemanticAttention(nn.Module):
    """HAN/models/SemanticAttention.py:5-20 (tiny dense reduce; stays torch)."""

    def __init__(self, in_size, hidden_size=128):
        super(SemanticAttention, self).__init__()
        self.project = nn.Sequential(
            nn.Linear(in_size, hidden_size),
            nn.Tanh(),
            nn.Linear(hidden_size, 1, bias=False)
        )

    def forward(self, z):
        w = self.project(z).mean(0)  # (M, 1)
        beta = torch.softmax(w, dim=0)  # (M, 1)
        beta = beta.expand((z.shape[0],) + beta.shape)  # (N, M, 1)
        return (beta * z).sum(1)  # (N, D * K)


class HANLayer(nn.Module):
    def __init__(self, num_meta_paths, in_size, out_size, layer_num_heads, dropout, **kwargs):
        super(HANLayer, self).__init__(**kwargs)
        self.gat_layers = nn.ModuleList()
        for i in range(num_meta_paths):
            self.gat_layers.add_module(f'meta_path_model{i}', GATConv(in_size, out_size, dropout, layer_num_heads))
        self.semantic_attention = SemanticAttention(in_size=out_size * layer_num_heads)

    def forward(self, gs, h):
        semantic_embeddings = []
        for g, gat_layer in zip(gs, self.gat_layers):
            semantic_embeddings.append(gat_layer(h, g).flatten(1))
        semantic_embeddings = torch.stack(semantic_embeddings, dim=1)  # (N, M, D * K)
        return self.semantic_attention(semantic_embeddings)


class HANModel(nn.Module):
    def __init__(self, num_mate_paths, in_size, hidden_size, out_size, num_heads, dropout, **kwargs):
        super(HANModel, self).__init__(**kwargs)
        self.layers = nn.ModuleList()
        self.layers.append(HANLayer(num_mate_paths, in_size, hidden_size, num_heads[0], dropout))
        for l in range(1, len(num_heads)):
            self.layers.append(
                HANLayer(num_mate_paths, hidden_size * num_heads[l - 1], hidden_size, num_heads[l], dropout))
        self.predict = nn.Linear(hidden_size * num_heads[-1], out_size)

    def forward(self, g, h):
        g = [adj_cache.get(a) for a in g]  # dense float64 masks -> CSR once
        for gnn in self.layers:
            h = gnn(g, h)
        return self.predict(h)
