"""Oracle: GAT / HAN attention (reference: GAT/models/layers.py, GAT/models/GAT.py,
HAN/models/NodeAttention.py, HAN/models/SemanticAttention.py, HAN/models/HAN.py)."""
import numpy as np
import torch
import torch.nn.functional as F


class replay_dropout:
    """Deterministic stand-in for F.dropout (same signature): call k draws its keep mask from torch's CPU generator
    seeded with base + k.  tests/golden/make_golden.py patches it into the reference while a train-mode fixture
    is made; the oracle and the GPU tests replay the same masks (`mask(k, shape, p)`)."""

    def __init__(self, base_seed):
        self.base, self.k = int(base_seed), 0

    @staticmethod
    def mask(seed, shape, p):
        g = torch.Generator().manual_seed(int(seed))
        return torch.bernoulli(torch.full(tuple(shape), 1.0 - p), generator=g) / (1.0 - p)

    def __call__(self, x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        keep = self.mask(self.base + self.k, x.shape, p).to(x.device)
        self.k += 1
        return x * keep


def dense_head(h, W, a, adj, alpha, concat=True, materialise_pairs=False, dropout=None, p=0.0):
    """GAT/models/layers.py:22-37 (== HAN/models/NodeAttention.py:22-38); eval mode unless `dropout`
    (an F.dropout-like callable, e.g. replay_dropout) is given: then layers.py:31 is applied to the attention.
    materialise_pairs=True follows the reference literally (the [N,N,2F'] tensor, :25-26);
    False uses the exact decomposition a·[Wh_i||Wh_j] = a[:F']·Wh_i + a[F':]·Wh_j so larger N
    fit in memory — same masked softmax and dense product after that."""
    Wh = torch.mm(h, W)
    N, Fp = Wh.shape
    if materialise_pairs:
        a_input = torch.cat([Wh.repeat(1, N).view(N * N, -1), Wh.repeat(N, 1)], dim=1).view(N, -1, 2 * Fp)
        e = F.leaky_relu(torch.matmul(a_input, a).squeeze(2), alpha)
    else:
        e = F.leaky_relu(Wh @ a[:Fp] + (Wh @ a[Fp:]).T, alpha)
    zero_vec = -9e15 * torch.ones_like(e)
    attention = torch.where(adj > 0, e, zero_vec)
    attention = F.softmax(attention, dim=1)
    if dropout is not None:
        attention = dropout(attention, p, True)
    h_prime = torch.matmul(attention, Wh)
    return F.elu(h_prime) if concat else h_prime


def sparse_head(x, W, a, adj, alpha, concat=True):
    """GAT/models/layers.py:94-131 eval mode: edge_e = exp(-LeakyReLU(a·[h_i||h_j])), rowsum,
    h' = (E·h)/rowsum.  `a` is [1, 2F']."""
    N = x.shape[0]
    edge = adj.nonzero().t()
    h = torch.mm(x, W)
    edge_h = torch.cat((h[edge[0, :], :], h[edge[1, :], :]), dim=1).t()
    edge_e = torch.exp(-F.leaky_relu(a.mm(edge_h).squeeze(), alpha))
    A = torch.sparse_coo_tensor(edge, edge_e, (N, N))
    e_rowsum = torch.matmul(A, torch.ones(N, 1, dtype=h.dtype))
    h_prime = torch.matmul(A, h).div(e_rowsum)
    return F.elu(h_prime) if concat else h_prime


def gat_model(x, params, adj, alpha, nheads, sparse=False, dropout=None, p=0.0):
    """GAT/models/GAT.py:14-18: cat of heads -> elu(out_att).  Eval mode unless `dropout` is given (dense heads
    only): then the four dropout sites of the train-mode forward (GAT.py:15,17 and layers.py:31) are applied in
    the reference's call order."""
    head = sparse_head if sparse else dense_head
    kw = {} if (dropout is None or sparse) else {"dropout": dropout, "p": p}
    if dropout is not None:
        x = dropout(x, p, True)
    xs = [head(x, params[f"attentions.AttentionHead{k}.W"], params[f"attentions.AttentionHead{k}.a"], adj, alpha, True, **kw)
          for k in range(nheads)]
    x = torch.cat(xs, dim=1)
    if dropout is not None:
        x = dropout(x, p, True)
    return F.elu(head(x, params["out_att.W"], params["out_att.a"], adj, alpha, False, **kw))


def gatconv(x, params, prefix, adj, alpha, nheads, num_class=None):
    """HAN/models/NodeAttention.py:58-62 eval mode, incl. the second ELU when num_class is None."""
    xs = [dense_head(x, params[f"{prefix}attentions.AttentionHead{k}.W"], params[f"{prefix}attentions.AttentionHead{k}.a"],
                     adj, alpha, True) for k in range(nheads)]
    x = torch.cat(xs, dim=1)
    if num_class is not None:
        return F.elu(dense_head(x, params[f"{prefix}out_att.W"], params[f"{prefix}out_att.a"], adj, alpha, False))
    return F.elu(x)


def semantic_attention(z, params, prefix):
    """HAN/models/SemanticAttention.py:15-20."""
    w = torch.tanh(F.linear(z, params[f"{prefix}project.0.weight"], params[f"{prefix}project.0.bias"]))
    w = F.linear(w, params[f"{prefix}project.2.weight"]).mean(0)
    beta = torch.softmax(w, dim=0)
    beta = beta.expand((z.shape[0],) + beta.shape)
    return (beta * z).sum(1)


def han_model(gs, h, params, nheads_list, alpha=0.2):
    """HAN/models/HAN.py:16-40 eval mode."""
    for l, nheads in enumerate(nheads_list):
        embs = [gatconv(h, params, f"layers.{l}.gat_layers.meta_path_model{m}.", g, alpha, nheads).flatten(1)
                for m, g in enumerate(gs)]
        z = torch.stack(embs, dim=1)
        h = semantic_attention(z, params, f"layers.{l}.semantic_attention.")
    return F.linear(h, params["predict.weight"], params["predict.bias"])


def edge_attention_f64(rowptr, col, Wh, s, t, alpha, mode=0, keep=None):
    """O(E) float64 restatement of one fused call (all heads): used at sizes where the dense
    N×N oracle does not fit.  Wh [N,H,Fp], s,t [N,H].  mode 0 softmax / 1 exp(-lrelu)."""
    Wh = np.asarray(Wh, np.float64); s = np.asarray(s, np.float64); t = np.asarray(t, np.float64)
    N, H, Fp = Wh.shape
    out = np.zeros((N, H, Fp))
    for i in range(N):
        js = col[rowptr[i]:rowptr[i + 1]]
        if len(js) == 0:
            out[i] = Wh.mean(0)
            continue
        z = s[i][None, :] + t[js]
        e = np.where(z > 0, z, alpha * z)
        if mode == 1:
            p = np.exp(-e)
        else:
            p = np.exp(e - e.max(0, keepdims=True))
        att = p / p.sum(0, keepdims=True)
        if keep is not None:
            att = att * keep[rowptr[i]:rowptr[i + 1]]
        out[i] = np.einsum("eh,ehf->hf", att, Wh[js])
    return out


def special_spmm(indices, values, shape, b, grad_output=None):
    """GAT/models/layers.py:43-64 restated in float64 numpy: out = S·b for the COO matrix
    (indices, values) — duplicates add — and, given grad_output, the two gradients of
    `SpecialSpmmFunction.backward`: grad_values[e] = <grad_output[row_e], b[col_e]> (layers.py:59-61
    picks these out of a dense N x N product) and grad_b = Sᵀ·grad_output (layers.py:63)."""
    import numpy as np
    rows, cols = np.asarray(indices[0]), np.asarray(indices[1])
    v, bb = np.asarray(values, dtype=np.float64), np.asarray(b, dtype=np.float64)
    out = np.zeros((int(shape[0]), bb.shape[1]))
    np.add.at(out, rows, v[:, None] * bb[cols])
    if grad_output is None:
        return out
    g = np.asarray(grad_output, dtype=np.float64)
    grad_values = np.einsum("ef,ef->e", g[rows], bb[cols])
    grad_b = np.zeros_like(bb)
    np.add.at(grad_b, cols, v[:, None] * g[rows])
    return out, grad_values, grad_b
