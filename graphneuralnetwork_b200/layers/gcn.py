"""Drop-in GCN layer and model (same class names, ctor signatures, parameter names and
forward signatures as /root/reference GCN/GCN.py:5-52); only `torch.spmm` at GCN/GCN.py:43
is replaced — by the sm_100a row-parallel CSR SpMM behind `functional.spmm`."""
import torch
from torch import nn

from ..functional import spmm
from ..graph import adj_cache


class Graph_conv_layer(nn.Module):
    """Y = Â·(X Wᵀ) + b — transform first, aggregate on the narrow width (GCN/GCN.py:41-47).
    state_dict: dense.weight [out,in], bias [out]."""

    def __init__(self, in_features, out_features, is_bias=True, **kwargs):
        super(Graph_conv_layer, self).__init__(**kwargs)
        self.in_features = in_features
        self.out_features = out_features
        self.dense = nn.Linear(in_features, out_features, bias=False)
        if is_bias:
            self.bias = nn.Parameter(torch.zeros(out_features))
        else:
            self.register_parameter('bias', None)

    def forward(self, X_input, adj):
        support = self.dense(X_input)  # dense X·W stays a torch matmul (north_star)
        output = spmm(adj_cache.get(adj), support)
        if self.bias is not None:
            return output + self.bias
        return output

    def __repr__(self):
        return self.__class__.__name__ + ' (' + str(self.in_features) + ' -> ' + str(self.out_features) + ')'


class GCN_Model(nn.Module):
    """GCN/GCN.py:5-27, unchanged in structure: the adjacency is routed to children whose
    `_get_name()` is 'Graph_conv_layer' (GCN/GCN.py:23)."""

    def __init__(self, num_features, num_hidden, num_classes, num_layers, dropout, **kwargs):
        super(GCN_Model, self).__init__(**kwargs)
        self.gcn_blocks = nn.Sequential()
        for i in range(num_layers):
            if i == 0:
                self.gcn_blocks.add_module(f'gcn{i}', Graph_conv_layer(num_features, num_hidden))
                self.gcn_blocks.add_module(f'relu{i}', nn.ReLU())
                self.gcn_blocks.add_module(f'dropout{i}', nn.Dropout(dropout))
            elif i == num_layers - 1:
                self.gcn_blocks.add_module(f'gcn{i}', Graph_conv_layer(num_hidden, num_classes))
            else:
                self.gcn_blocks.add_module(f'gcn{i}', Graph_conv_layer(num_hidden, num_hidden))
                self.gcn_blocks.add_module(f'relu{i}', nn.ReLU())
                self.gcn_blocks.add_module(f'dropout{i}', nn.Dropout(dropout))

    def forward(self, X, adj):
        adj = adj_cache.get(adj)  # COO -> CSR once per graph
        for gcn_block in self.gcn_blocks:
            if gcn_block._get_name() == 'Graph_conv_layer':
                X = gcn_block(X, adj)
            else:
                X = gcn_block(X)
        return X
