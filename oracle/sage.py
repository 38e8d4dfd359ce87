"""Oracle: GraphSAGE sampling blocks and aggregation (reference:
GraphSAGE_Pytorch/sample_utils.py, data_utils.py, models/*.py; GraphSAGE/graph_utils.py,
GraphSAGE.py)."""
import random

import numpy as np
import torch
import torch.nn.functional as F


def sampling(src_nodes, sample_num, neighbor_table):
    """GraphSAGE_Pytorch/sample_utils.py:4-17: without replacement if deg >= k, else with;
    results extended flat, src-major."""
    results = []
    for sid in src_nodes:
        if len(neighbor_table[sid]) < sample_num:
            res = random.choices(list(neighbor_table[sid]), k=sample_num)
        else:
            res = random.sample(list(neighbor_table[sid]), k=sample_num)
        results.extend(res)
    return results


def multihop_sampling(src_nodes, sample_nums, neighbor_table):
    """GraphSAGE_Pytorch/sample_utils.py:20-35."""
    sampling_result = [src_nodes]
    for k, hopk_num in enumerate(sample_nums):
        sampling_result.append(sampling(sampling_result[k], hopk_num, neighbor_table))
    return sampling_result


def gather_features(feat_data, sampling_nodes):
    """GraphSAGE_Pytorch/data_utils.py:64 (vectorised stand-in for the python list gather)."""
    return feat_data[torch.as_tensor(np.asarray(sampling_nodes, dtype=np.int64))]


def aggregate(neighbor_feature, aggr_method="mean"):
    """GraphSAGE_Pytorch/models/Aggregator.py:19-27.  `max`: the reference's `.max(dim=1)`
    returns a tuple and its matmul raises TypeError; the oracle (and the kernel) use `.values`."""
    if aggr_method == "mean":
        return neighbor_feature.mean(dim=1)
    if aggr_method == "sum":
        return neighbor_feature.sum(dim=1)
    if aggr_method == "max":
        return neighbor_feature.max(dim=1).values
    raise ValueError("Unknown aggr type, expected sum, max, or mean, but got {}".format(aggr_method))


def sage_gcn(src, neigh, w_self, w_agg, activation=True, aggr="mean", combine="sum"):
    """GraphSAGE_Pytorch/models/SageGCN.py:23-36 + Aggregator.py:29."""
    neighbor_hidden = torch.matmul(aggregate(neigh, aggr), w_agg)
    self_hidden = torch.matmul(src, w_self)
    hidden = self_hidden + neighbor_hidden if combine == "sum" else torch.cat([self_hidden, neighbor_hidden], dim=1)
    return F.relu(hidden) if activation else hidden


def graphsage_forward(node_features_list, params, num_neighbors_list):
    """GraphSAGE_Pytorch/models/GraphSage.py:18-30."""
    L = len(num_neighbors_list)
    hidden = node_features_list
    for l in range(L):
        nxt = []
        for hop in range(L - l):
            src = hidden[hop]
            neigh = hidden[hop + 1].view((len(src), num_neighbors_list[hop], -1))
            nxt.append(sage_gcn(src, neigh, params[f"gcn.{l}.weight"], params[f"gcn.{l}.aggregator.weight"],
                                activation=(l != L - 1)))
        hidden = nxt
    return hidden[0]


def aggregator_v2(neigh_feat, agg_func="MEAN"):
    """GraphSAGE/graph_utils.py:4-11; 'MAX' there is argmax (a bug) — values used instead."""
    if agg_func == "MEAN":
        return torch.mean(neigh_feat, dim=1)
    if agg_func == "MAX":
        return torch.max(neigh_feat, dim=1).values
    raise ValueError(agg_func)


def embed_mean_v2(feats, index_map):
    """GraphSAGE/GraphSAGE.py:47-49 + graph_utils.py:6: embedding of the rows whose first
    column is not -1, then mean over the k axis."""
    m = index_map[index_map[:, 0] != -1, :]
    return torch.mean(torch.embedding(feats, m), dim=1)
