"""Oracle: GCN graph construction and layer (reference: GCN/data_utils.py, GCN/GCN.py)."""
import numpy as np
import scipy.sparse as sp
import torch


def symmetrise(edges: np.ndarray, n: int):
    """GCN/data_utils.py:32-35: COO of ones from the directed edge list, then
    adj + adj.T*(adj.T > adj) - adj*(adj.T > adj)  (max-symmetrisation)."""
    adj = sp.coo_matrix((np.ones(edges.shape[0]), (edges[:, 0], edges[:, 1])), shape=(n, n), dtype=np.float32)
    adj = adj + adj.T.multiply(adj.T > adj) - adj.multiply(adj.T > adj)
    return adj


def normalize_adj(mx):
    """GCN/data_utils.py:54-60: D^-1/2 (A) D^-1/2 via (A·D).T·D, inf -> 0; float64 when the
    caller added sp.eye (GCN/data_utils.py:78)."""
    mx = sp.coo_matrix(mx)
    rowsum = np.array(mx.sum(1))
    with np.errstate(divide="ignore"):
        d_inv_sqrt = np.power(rowsum, -0.5).flatten()
    d_inv_sqrt[np.isinf(d_inv_sqrt)] = 0.
    d_mat_inv_sqrt = sp.diags(d_inv_sqrt)
    return mx.dot(d_mat_inv_sqrt).transpose().dot(d_mat_inv_sqrt)


def normalize_features(mx):
    """GCN/data_utils.py:39-51: row-normalise."""
    rowsum = np.array(mx.sum(1))
    with np.errstate(divide="ignore"):
        r_inv = np.power(rowsum.astype(float), -1).flatten()
    r_inv[np.isinf(r_inv)] = 0
    return sp.diags(r_inv).dot(mx)


def to_coo_arrays(sparse_mx):
    """GCN/data_utils.py:63-70 up to the tensor constructor: .tocoo().astype(float32);
    returns (row int64, col int64, val float32) in scipy's order."""
    m = sparse_mx.tocoo().astype(np.float32)
    return m.row.astype(np.int64), m.col.astype(np.int64), m.data


def build_adjacency(edges: np.ndarray, n: int):
    """load_cora's adjacency pipeline (GCN/data_utils.py:76-85) on a directed edge list."""
    adj = symmetrise(edges, n)
    adj = normalize_adj(adj + sp.eye(adj.shape[0]))
    return to_coo_arrays(adj)


def coo_to_csr(row, col, val, n_rows):
    """Stable row sort -> (rowptr int64, col int32, val float32): the layout the kernels walk."""
    order = np.argsort(row, kind="stable")
    rowptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(rowptr, row + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, col[order].astype(np.int32), None if val is None else val[order].astype(np.float32)


def csr_transpose(rowptr, col, val, n_rows, n_cols):
    """CSR of the transpose, stable in the original row order; also returns the permutation."""
    rows = np.repeat(np.arange(n_rows, dtype=np.int32), np.diff(rowptr))
    order = np.argsort(col, kind="stable")
    rowptr_t = np.zeros(n_cols + 1, dtype=np.int64)
    np.add.at(rowptr_t, col.astype(np.int64) + 1, 1)
    rowptr_t = np.cumsum(rowptr_t)
    return rowptr_t, rows[order], None if val is None else val[order], order.astype(np.int64)


def dense_mask_to_csr(adj: np.ndarray):
    """`adj > 0` -> CSR pattern in adj.nonzero() order (GAT/models/layers.py:29,98)."""
    mask = adj > 0
    r, c = np.nonzero(mask)
    rowptr = np.zeros(adj.shape[0] + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    return np.cumsum(rowptr), c.astype(np.int32)


def spmm(row, col, val, X: torch.Tensor, n_rows: int) -> torch.Tensor:
    """GCN/GCN.py:43: torch.spmm on the COO tensor sparse_mx_to_torch_sparse_tensor builds."""
    adj = torch.sparse_coo_tensor(torch.from_numpy(np.vstack((row, col))), torch.from_numpy(np.asarray(val)),
                                  (n_rows, X.shape[0]))
    return torch.spmm(adj, X)


def spmm_f64(rowptr, col, val, X: np.ndarray) -> np.ndarray:
    """Float64 CSR product (ground truth for tolerance checks at sizes torch.spmm handles slowly)."""
    n = len(rowptr) - 1
    v = np.ones(len(col)) if val is None else val.astype(np.float64)
    m = sp.csr_matrix((v, col.astype(np.int64), rowptr), shape=(n, X.shape[0]))
    return m @ X.astype(np.float64)


def graph_conv_layer(X, weight, bias, coo, n):
    """GCN/GCN.py:41-47: support = X·Wᵀ; out = spmm(adj, support) (+ bias)."""
    support = torch.nn.functional.linear(X, weight)
    out = spmm(coo[0], coo[1], coo[2], support, n)
    return out + bias if bias is not None else out


def gcn_model(X, params, coo, n, num_layers=2):
    """GCN/GCN.py:5-27 in eval mode (dropout = identity): gcn -> relu -> ... -> gcn."""
    for i in range(num_layers):
        X = graph_conv_layer(X, params[f"gcn_blocks.gcn{i}.dense.weight"], params.get(f"gcn_blocks.gcn{i}.bias"), coo, n)
        if i != num_layers - 1:
            X = torch.relu(X)
    return X
