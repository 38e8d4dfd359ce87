"""Drop-in GTN pieces (class / function names, ctor signatures, parameter names and forward signatures of
/root/reference GTN/models/{GTN,GTLayer,GTConv}.py) — SURVEY.md §8f rank 4.

On the message-passing path (re-designed):
  * `gcn_conv(X, H)` (GTN.py:49-52): the GCN over the LEARNED metapath adjacency H.  The reference normalises the
    dense H with two N x N `torch.eye` products (`norm`, GTN.py:7-19: an O(N^3) `torch.mm(deg_inv, H)`) and
    aggregates with a dense `torch.mm(H.t(), X)`.  Here the non-zero pattern of H (off-diagonal) plus the
    self-loops is turned into a CSR of H^T ON THE DEVICE (gnn_dense_mask_count / _fill), the edge values are
    gathered from H (autograd reaches H through the gather), column-normalised in O(nnz), and the aggregation
    is the CSR SpMM with gradients into the edge VALUES (`functional.spmm_values`: SpMM + edge-gradient
    SDDMM + transpose SpMM).
  * `norm(H, add)` keeps the reference's dense-in / dense-out contract for `GTN_Model.normalization`, as
    element-wise work: multiplying by a diagonal matrix is a row scaling, and adds only exact zeros, so the
    result is bit-identical to the reference's `torch.mm(deg_inv, H)`.
Adjacent, kept as plain torch (out of scope: the dense N^3 metapath composition): `GTConv`, `GTLayer`.
"""
import torch
from torch import nn
import torch.nn.functional as F

from ..functional import spmm_values
from ..graph import CSRGraph


def norm(H, add=False):
    """GTN/models/GTN.py:7-19: zero the diagonal (add=True: then put ones on it), divide every COLUMN of H by its
    sum (`deg^-1`, inf -> 0).  Dense in, dense out."""
    Ht = H.t()
    n = Ht.shape[0]
    off = (torch.eye(n, device=H.device) == 0).to(torch.float32)
    Ht = Ht * off
    if add:
        Ht = Ht + torch.eye(n, device=H.device, dtype=torch.float32)
    deg = torch.sum(Ht, dim=1)
    deg_inv = deg.pow(-1)
    deg_inv = torch.where(deg_inv == float('inf'), torch.zeros_like(deg_inv), deg_inv)
    return (deg_inv.unsqueeze(1) * Ht).t()


def _pattern_of_learned_adjacency(H):
    """CSR of (offdiag(H) != 0 | I)^T built on the device, plus the flat positions of its edges in H."""
    n = H.shape[0]
    with torch.no_grad():
        mask = (H.detach().t() != 0).to(torch.float32)
        mask.fill_diagonal_(1.0)  # the +I of norm(add=True); the diagonal VALUE of H itself is dropped
        g = CSRGraph.from_dense_mask(mask.contiguous())  # rows = columns of H (H^T), ascending source ids
        rows = g.edge_rows().to(torch.int64)   # j  (row of H^T)
        cols = g.col.to(torch.int64)           # i  (column of H^T) -> entry H[i, j]
    return g, rows, cols


def gcn_conv(X, H, weight):
    """`GTN_Model.gcn_conv` (GTN.py:49-52): D^-1 (offdiag(H) + I)^T (X W) with D the column sums of
    offdiag(H) + I.  X [N, w_in], H [N, N] dense (learned, autograd flows into it), weight [w_in, w_out].
    The gradient w.r.t. H is produced on the support of H (its non-zero entries): the reference's dense autograd
    also fills the structural zeros, but in GTN a zero of the composed adjacency is a zero of every factor
    product, so no parameter receives anything from them (the model-level gradients are identical)."""
    XW = torch.mm(X, weight)
    n = H.shape[0]
    g, rows, cols = _pattern_of_learned_adjacency(H)
    diag = rows == cols
    vals = torch.where(diag, torch.ones((), device=H.device, dtype=H.dtype), H[cols, rows])  # (offdiag(H) + I)[i, j]
    deg = torch.zeros(n, device=H.device, dtype=vals.dtype).index_add(0, rows, vals)        # column sums, by row of H^T
    deg_inv = deg.pow(-1)
    deg_inv = torch.where(deg_inv == float('inf'), torch.zeros_like(deg_inv), deg_inv)
    return spmm_values(g, (vals * deg_inv[rows]).float(), XW)


class GTConv(nn.Module):
    """GTN/models/GTConv.py: softmax-weighted sum of the edge-type adjacencies (1x1 convolution)."""

    def __init__(self, in_channels, out_channels, **kwargs):
        super(GTConv, self).__init__(**kwargs)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels, 1, 1))
        self.bias = None
        self.scale = nn.Parameter(torch.Tensor([0.1]), requires_grad=False)

    def forward(self, A):
        return torch.sum(A * F.softmax(self.weight, dim=1), dim=1)


class GTLayer(nn.Module):
    """GTN/models/GTLayer.py: one metapath-composition step (dense bmm; out of scope, kept as the reference has it)."""

    def __init__(self, in_channels, out_channels, first=True):
        super(GTLayer, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.first = first
        self.conv1 = GTConv(in_channels, out_channels)
        if self.first:
            self.conv2 = GTConv(in_channels, out_channels)

    def forward(self, A, H_=None):
        if self.first:
            a = self.conv1(A)
            b = self.conv2(A)
            H = torch.bmm(a, b)
            W = [(F.softmax(self.conv1.weight, dim=1)).detach(), (F.softmax(self.conv2.weight, dim=1)).detach()]
        else:
            a = self.conv1(A)
            H = torch.bmm(H_, a)
            W = [(F.softmax(self.conv1.weight, dim=1)).detach()]
        return H, W


class GTN_Model(nn.Module):
    """GTN/models/GTN.py:22-88.  state_dict: `layers.{l}.conv{1,2}.{weight,scale}`, `weight`, `bias`,
    `linear1.*`, `linear2.*`."""

    def __init__(self, num_edge, num_channels, w_in, w_out, num_class, num_layers, is_norm, **kwargs):
        super(GTN_Model, self).__init__(**kwargs)
        self.num_edge = num_edge
        self.num_channels = num_channels
        self.w_in = w_in
        self.w_out = w_out
        self.num_class = num_class
        self.num_layers = num_layers
        self.is_norm = is_norm
        self.layers = nn.ModuleList([GTLayer(num_edge, num_channels, first=(i == 0)) for i in range(num_layers)])
        self.weight = nn.Parameter(torch.Tensor(w_in, w_out))
        self.bias = nn.Parameter(torch.Tensor(w_out))
        self.linear1 = nn.Linear(self.w_out * self.num_channels, self.w_out)
        self.linear2 = nn.Linear(self.w_out, self.num_class)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.weight)
        nn.init.zeros_(self.bias)

    def gcn_conv(self, X, H):
        return gcn_conv(X, H, self.weight)

    def normalization(self, H):
        return torch.stack([norm(H[i, :, :]) for i in range(self.num_channels)], dim=0)

    def forward(self, A, X, target_x):
        A = A.unsqueeze(0).permute(0, 3, 1, 2)
        Ws = []
        for i in range(self.num_layers):
            if i == 0:
                H, W = self.layers[i](A)
            else:
                H = self.normalization(H)
                H, W = self.layers[i](A, H)
            Ws.append(W)
        X_ = torch.cat([F.relu(self.gcn_conv(X, H[i])) for i in range(self.num_channels)], dim=1)
        X_ = F.relu(self.linear1(X_))
        y = self.linear2(X_[target_x])
        return y, Ws
