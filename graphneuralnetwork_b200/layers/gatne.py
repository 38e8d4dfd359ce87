"""Drop-in GATNE encoders: the per-edge-type neighbour aggregation of
/root/reference GATNE_Pytorch/models/GATNE.py:6-98 (`GraphEncoder`, `GraphDecoder`, `GATNEModel`)
and GATNE/models/GATNE.py:7-77 (`GATNEModel`, exported here as `GATNEModelV1`).

Fixed by the reference: class names, constructor / forward signatures, parameter names and shapes
(`node_embeddings [N,E]`, `node_type_embeddings [N,T,U]`, `embed_trans [Fd,E]`,
`u_embed_trans [T,Fd,U]`, `trans_weights [T,U,E]`, `trans_weights_s1 [T,U,A]`,
`trans_weights_s2 [T,A,1]`, decoder `weights [N,E]`).

Re-designed for the device:
  * GATNE-T: `node_type_embeddings[node_neigh]` ([B,T,K,T,U] + `torch.diagonal` in GATNE.py:53,57;
    a `torch.cat` of T gathers in GATNE_Pytorch) followed by the sum / mean over K is ONE launch of
    the typed gather-reduce (`functional.typed_gather_reduce`): output row (b,t) reads the type-t
    embedding of each of its K neighbours and adds them in order — no intermediate;
  * GATNE-I: the reference projects every gathered neighbour (`[B,T,K,Fd] x [T,Fd,U]`) and then
    reduces over K; the reduce is linear, so the neighbours' raw features are aggregated first
    (fused gather-sum over the feature table) and the projection runs on `[B,T,Fd]` — K times
    fewer matmul flops and no `[B,T,K,Fd]` gather;
  * the type attention (softmax over T, a few KB) stays torch.
"""
import math

import torch
from torch import nn
import torch.nn.functional as F

from ..functional import gather_reduce, typed_gather_reduce

_AGG = {"SUM": "sum", "MEAN": "mean"}


class _GatneCore(nn.Module):
    """Parameters and forward shared by both reference variants (they differ in initialisation,
    in the name of the attention-width argument and in whether MEAN is offered)."""

    def _declare(self, num_nodes, embedding_size, embedding_u_size, edge_type_count, dim_a, features):
        self.num_nodes, self.embedding_size, self.embedding_u_size = num_nodes, embedding_size, embedding_u_size
        self.edge_type_count, self.dim_a = edge_type_count, dim_a
        self.features = features
        if features is not None:
            self.feature_dim = features.shape[-1]
            self.embed_trans = nn.Parameter(torch.empty(self.feature_dim, embedding_size))
            self.u_embed_trans = nn.Parameter(torch.empty(edge_type_count, self.feature_dim, embedding_u_size))
        else:
            self.node_embeddings = nn.Parameter(torch.empty(num_nodes, embedding_size))
            self.node_type_embeddings = nn.Parameter(torch.empty(num_nodes, edge_type_count, embedding_u_size))
        self.trans_weights = nn.Parameter(torch.empty(edge_type_count, embedding_u_size, embedding_size))
        self.trans_weights_s1 = nn.Parameter(torch.empty(edge_type_count, embedding_u_size, dim_a))
        self.trans_weights_s2 = nn.Parameter(torch.empty(edge_type_count, dim_a, 1))

    def _encode(self, inputs, node_types, node_neigh, agg_func):
        if agg_func not in _AGG:
            raise ValueError("please choice else aggregator!")
        reduce = _AGG[agg_func]
        B, T, K = node_neigh.shape
        if self.features is None:
            base = self.node_embeddings[inputs]
            per_type = typed_gather_reduce(self.node_type_embeddings, node_neigh, reduce)          # [B,T,U]
        else:
            feats = self.features
            base = feats[inputs] @ self.embed_trans
            pooled = gather_reduce(feats, node_neigh.reshape(-1), B * T, K, reduce)                  # [B*T,Fd]
            per_type = torch.einsum('btf,tfu->btu', pooled.view(B, T, -1), self.u_embed_trans)     # [B,T,U]
        # attention over the T edge types with the weights of each sample's own type
        s1, s2 = self.trans_weights_s1[node_types], self.trans_weights_s2[node_types]
        att = torch.softmax(torch.bmm(torch.tanh(torch.bmm(per_type, s1)), s2).squeeze(2), dim=1)  # [B,T]
        mixed = torch.einsum('bt,btu->bu', att, per_type)
        out = base + torch.bmm(mixed.unsqueeze(1), self.trans_weights[node_types]).squeeze(1)
        return F.normalize(out, dim=1)


class GraphEncoder(_GatneCore):
    """GATNE_Pytorch/models/GATNE.py:6-98."""

    def __init__(self, num_nodes, embedding_size, embedding_u_size, edge_type_count, attention_size, features,
                 agg_func='SUM', **kwargs):
        super().__init__(**kwargs)
        self.agg_func = agg_func
        self._declare(num_nodes, embedding_size, embedding_u_size, edge_type_count, attention_size, features)
        self.reset_parameters()

    def reset_parameters(self):
        if self.features is not None:
            nn.init.xavier_uniform_(self.embed_trans.data)
            nn.init.xavier_uniform_(self.u_embed_trans.data)
        else:
            nn.init.uniform_(self.node_embeddings.data)
            nn.init.uniform_(self.node_type_embeddings.data)
        for w in (self.trans_weights, self.trans_weights_s1, self.trans_weights_s2):
            nn.init.xavier_uniform_(w.data)

    def forward(self, inputs, node_types, node_neigh):
        return self._encode(inputs, node_types, node_neigh, self.agg_func)


class GraphDecoder(nn.Module):
    """Scores of every centre against its context / negative nodes (GATNE.py:101-114)."""

    def __init__(self, num_nodes, embedding_size, **kwargs):
        super().__init__(**kwargs)
        self.num_nodes, self.embedding_size = num_nodes, embedding_size
        self.weights = nn.Parameter(nn.init.xavier_uniform_(torch.empty(num_nodes, embedding_size)))

    def forward(self, embed, contest_negative):
        return torch.einsum('be,bce->bc', embed, self.weights[contest_negative])


class GATNEModel(nn.Module):
    """Encoder + decoder (GATNE_Pytorch/models/GATNE.py:117-128)."""

    def __init__(self, num_nodes, embedding_size, embedding_u_size, edge_type_count, attention_size, features,
                 **kwargs):
        super().__init__()
        self.encoder = GraphEncoder(num_nodes, embedding_size, embedding_u_size, edge_type_count, attention_size,
                                    features, **kwargs)
        self.decoder = GraphDecoder(num_nodes, embedding_size)

    def forward(self, inputs, node_types, node_neigh, context_negative):
        return self.decoder(self.encoder(inputs, node_types, node_neigh), context_negative)


class GATNEModelV1(_GatneCore):
    """GATNE/models/GATNE.py:7-77 (class `GATNEModel` there): SUM aggregation only, normal / uniform
    initialisation scaled by 1/sqrt(embedding_size)."""

    def __init__(self, num_nodes, embedding_size, embedding_u_size, edge_type_count, dim_a, features, **kwargs):
        super().__init__(**kwargs)
        self._declare(num_nodes, embedding_size, embedding_u_size, edge_type_count, dim_a, features)
        self.reset_parameters()

    def reset_parameters(self):
        std = 1.0 / math.sqrt(self.embedding_size)
        if self.features is not None:
            self.embed_trans.data.normal_(std=std)
            self.u_embed_trans.data.normal_(std=std)
        else:
            self.node_embeddings.data.uniform_(-1.0, 1.0)
            self.node_type_embeddings.data.uniform_(-1.0, 1.0)
        for w in (self.trans_weights, self.trans_weights_s1, self.trans_weights_s2):
            w.data.normal_(std=std)

    def forward(self, train_inputs, train_types, node_neigh):
        return self._encode(train_inputs, train_types, node_neigh, 'SUM')
