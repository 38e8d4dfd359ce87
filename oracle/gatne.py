"""Oracle: GATNE's per-edge-type neighbour aggregation + type attention, restated on the CPU
(test infrastructure only; reference: GATNE_Pytorch/models/GATNE.py:57-98 `GraphEncoder.forward`
and GATNE/models/GATNE.py:50-77 `GATNEModel.forward`).  Pinned against outputs of the unmodified
reference modules in tests/golden/gatne_small.npz (tests/golden/make_golden.py)."""
import torch
import torch.nn.functional as F


def neighbour_aggregate(node_type_embeddings, node_neigh, agg_func="SUM"):
    """[N,T,U] table, [B,T,K] ids -> [B,T,U]: neighbour k of (b, type t) contributes its TYPE-t
    embedding.  GATNE_Pytorch/models/GATNE.py:61-63 builds it as a cat over t of
    `node_type_embeddings[:, t, :][node_neigh[:, t, :]]`; GATNE/models/GATNE.py:53,57 gathers all T
    embeddings of every neighbour and keeps the diagonal (type of the slot == type of the
    embedding); both then reduce over K (GATNE.py:72-77 / :58)."""
    B, T, K = node_neigh.shape
    per_type = torch.stack([node_type_embeddings[node_neigh[:, t, :], t, :] for t in range(T)], dim=1)  # [B,T,K,U]
    if agg_func == "SUM":
        return per_type.sum(dim=2)
    if agg_func == "MEAN":
        return per_type.mean(dim=2)
    raise ValueError("please choice else aggregator!")


def neighbour_aggregate_features(features, u_embed_trans, node_neigh, agg_func="SUM"):
    """GATNE-I: neighbours are projected per type before the reduce
    (GATNE_Pytorch/models/GATNE.py:66-70 bmm; GATNE/models/GATNE.py:56 einsum + diagonal)."""
    B, T, K = node_neigh.shape
    proj = torch.stack([features[node_neigh[:, t, :]] @ u_embed_trans[t] for t in range(T)], dim=1)  # [B,T,K,U]
    return proj.sum(dim=2) if agg_func == "SUM" else proj.mean(dim=2)


def type_attention(node_embed, node_type_embed, node_types, trans_weights, trans_weights_s1, trans_weights_s2):
    """GATNE_Pytorch/models/GATNE.py:79-98 == GATNE/models/GATNE.py:60-77: softmax over the T edge
    types of tanh(U·s1)·s2, weighted sum of the per-type aggregates, projection by M_r of the
    sample's own type, residual on the base embedding, L2 normalisation."""
    w, s1, s2 = trans_weights[node_types], trans_weights_s1[node_types], trans_weights_s2[node_types]
    att = F.softmax(torch.matmul(torch.tanh(torch.matmul(node_type_embed, s1)), s2).squeeze(2), dim=1).unsqueeze(1)
    mixed = torch.matmul(att, node_type_embed)
    return F.normalize(node_embed + torch.matmul(mixed, w).squeeze(1), dim=1)


def encoder_forward(params, inputs, node_types, node_neigh, features=None, agg_func="SUM"):
    if features is None:
        base = params["node_embeddings"][inputs]
        per_type = neighbour_aggregate(params["node_type_embeddings"], node_neigh, agg_func)
    else:
        base = features[inputs] @ params["embed_trans"]
        per_type = neighbour_aggregate_features(features, params["u_embed_trans"], node_neigh, agg_func)
    return type_attention(base, per_type, node_types, params["trans_weights"], params["trans_weights_s1"],
                          params["trans_weights_s2"])
