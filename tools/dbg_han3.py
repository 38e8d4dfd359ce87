import sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from graphneuralnetwork_b200 import layers, synthetic as S, _lib
from graphneuralnetwork_b200.functional import gat_aggregate
from graphneuralnetwork_b200.graph import CSRGraph, adj_cache
DEV = "cuda"
n = S.ACM["n"]; H, Fp, M = 8, 8, 3
masks = [torch.from_numpy(S.symmetric_mask(n, t, seed=11 + i)).to(DEV) for i, t in enumerate(S.ACM["metapath_nnz"])]
graphs = [adj_cache.get(m) for m in masks]
big = CSRGraph.block_diagonal(graphs)
for scale in (1.0, 2.0, 4.0):
    torch.manual_seed(0)
    Wh = (torch.randn(n, M * H * Fp, device=DEV) * scale)
    s = torch.randn(M * n, H, device=DEV) * scale
    t = torch.randn(M * n, H, device=DEV) * scale
    dO = torch.randn(n, M * H * Fp, device=DEV)
    def run_b():
        a, b, c = (x.clone().requires_grad_(True) for x in (Wh, s, t))
        o = gat_aggregate(big, a, b, c, H, Fp, 0.2, elu=1, batch=M)
        o.backward(dO)
        return o.detach(), a.grad, b.grad, c.grad
    def run_u():
        outs = []
        for m in range(M):
            a = Wh[:, m * 64:(m + 1) * 64].contiguous().requires_grad_(True)
            b = s[m * n:(m + 1) * n].contiguous().requires_grad_(True)
            c = t[m * n:(m + 1) * n].contiguous().requires_grad_(True)
            o = gat_aggregate(graphs[m], a, b, c, H, Fp, 0.2, elu=1)
            o.backward(dO[:, m * 64:(m + 1) * 64].contiguous())
            outs.append((o.detach(), a.grad, b.grad, c.grad))
        return (torch.cat([x[0] for x in outs], 1), torch.cat([x[1] for x in outs], 1),
                torch.cat([x[2] for x in outs], 0), torch.cat([x[3] for x in outs], 0))
    B1, B2, U = run_b(), run_b(), run_u()
    for name, i in (("out", 0), ("dWh", 1), ("ds", 2), ("dt", 3)):
        d = (B1[i] - U[i]).abs()
        print(scale, name, "b-vs-b", float((B1[i] - B2[i]).abs().max()), "b-vs-u max", float(d.max()), "ref max", float(U[i].abs().max()),
              "argmax", np.unravel_index(int(d.argmax()), d.shape))
    # per metapath/head dWh
    d = (B1[1] - U[1]).abs().view(n, M, H, Fp).amax(dim=(0, 3)) / U[1].abs().view(n, M, H, Fp).amax(dim=(0, 3))
    print(d.cpu().numpy())
