"""Drop-in for the second GraphSAGE variant (/root/reference GraphSAGE/GraphSAGE.py:7-61): per
layer de-duplicated node sets with `-1`-padded index maps (GraphSAGE/data_utils.py:82-117).

Same class names, constructor and forward signature, parameter names
(`sage_blocks.sage_layer{i}.weight.weight`, `dense.*`).  What changes: `Aggregator`
(graph_utils.py:6) and the `torch.embedding` + mean between layers (GraphSAGE.py:47-49) run as
the fused gather-mean kernel — the `[n,k,F]` neighbour tensor of the inner layers is never
materialised."""
import torch
from torch import nn
import torch.nn.functional as F

from .sage import Aggregator, gather_mean


class SageLayer(nn.Module):
    def __init__(self, input_size, output_size, gcn=False, **kwargs):
        super(SageLayer, self).__init__(**kwargs)
        self.input_size = input_size
        self.output_size = output_size
        self.gcn = gcn
        self.weight = nn.Linear(self.input_size if self.gcn else 2 * self.input_size, self.output_size, bias=False)

    def forward(self, self_feats, aggregate_feats):
        if not self.gcn:
            combined = torch.cat([self_feats, aggregate_feats], dim=1)
        else:
            combined = aggregate_feats
        return F.relu(self.weight(combined))


class GraphSAGE(nn.Module):
    def __init__(self, num_layers, input_size, out_size, gcn=False, agg_func='MEAN', Unsupervised=True, class_size=None,
                 **kwargs):
        super(GraphSAGE, self).__init__(**kwargs)
        self.num_layers = num_layers
        self.gcn = gcn
        self.agg_func = agg_func
        self.sage_blocks = nn.Sequential()
        for index in range(0, num_layers):
            layer_size = out_size if index != 0 else input_size
            self.sage_blocks.add_module('sage_layer' + str(index), SageLayer(layer_size, out_size, gcn=self.gcn))
        self.Unsupervised = Unsupervised
        if not Unsupervised:
            self.dense = nn.Linear(out_size, class_size)

    def forward(self, center_feats_data, center_nodes_map, center_neigh_feats_data, center_neigh_nodes_map,
                contexts_negatives_feats_data, contexts_negatives_nodes_map, contexts_negatives_neigh_feats_data,
                contexts_negatives_neigh_nodes_map, contexts_negatives_shape):
        if contexts_negatives_feats_data is None:  # supervised path (GraphSAGE.py:42-53)
            aggregated = Aggregator(center_neigh_feats_data, self.agg_func)  # pre-gathered outermost layer
            feats_data = None
            for i, block in enumerate(self.sage_blocks):
                feats_data = block(center_feats_data, aggregated)
                if i != self.num_layers - 1:
                    cmap = center_nodes_map[i]
                    center_feats_data = torch.embedding(feats_data, cmap[cmap != -1])
                    nmap = center_neigh_nodes_map[i]
                    valid = nmap[nmap[:, 0] != -1, :]  # -1 padded rows are dropped (GraphSAGE.py:48-49)
                    if self.agg_func == 'MEAN':
                        aggregated = gather_mean(feats_data, valid)  # fused embedding + mean
                    else:
                        aggregated = Aggregator(torch.embedding(feats_data, valid), self.agg_func)
            classes = None
            if not self.Unsupervised:
                classes = self.dense(feats_data)
            return feats_data, classes
        # unsupervised skip-gram head (GraphSAGE.py:54-61): two supervised passes + a small bmm
        center_feats_data, _ = self(center_feats_data, center_nodes_map, center_neigh_feats_data,
                                    center_neigh_nodes_map, None, None, None, None, None)
        contexts_negatives_feats_data, _ = self(contexts_negatives_feats_data, contexts_negatives_nodes_map,
                                                contexts_negatives_neigh_feats_data,
                                                contexts_negatives_neigh_nodes_map, None, None, None, None, None)
        contexts_negatives_feats_data = contexts_negatives_feats_data.reshape(*contexts_negatives_shape, -1)
        return center_feats_data, torch.bmm(center_feats_data.unsqueeze(1),
                                            contexts_negatives_feats_data.permute(0, 2, 1))
