"""Diagnostic (GPU): the batched block-diagonal HAN attention launch against M separate launches on the ACM-sized
golden's parameters — out / dWh / ds / dt of both, and both against a float64 dense autograd of metapath 1.
    gpurun -- python tools/diag_han_batched.py"""
import sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from graphneuralnetwork_b200 import layers, synthetic as S, _lib
from graphneuralnetwork_b200.functional import gat_aggregate
from graphneuralnetwork_b200.graph import CSRGraph, adj_cache
from conftest import load_golden
DEV = "cuda"
g = load_golden("han_acm.npz")
n = S.ACM["n"]; H, Fp, M = 8, 8, 3
masks = [torch.from_numpy(S.symmetric_mask(n, t, seed=11 + i)).to(DEV) for i, t in enumerate(S.ACM["metapath_nnz"])]
X = torch.from_numpy(np.random.default_rng(14).standard_normal((n, S.ACM["feats"]), dtype=np.float32)).to(DEV)
model = layers.HANModel(3, S.ACM["feats"], 8, S.ACM["classes"], [8], 0.0)
model.load_state_dict({k: torch.from_numpy(g[k]) for k in model.state_dict().keys()})
model = model.to(DEV).train()
lay = model.layers[0]
graphs = [adj_cache.get(m) for m in masks]
big = CSRGraph.block_diagonal(graphs)
convs = list(lay.gat_layers)
Ws = torch.cat([torch.cat([hd.W for hd in c.attentions], 1) for c in convs], 1).detach()
a = torch.stack([torch.stack([hd.a[:, 0] for hd in c.attentions]) for c in convs]).detach()
Wh0 = X @ Ws
Wh4 = Wh0.view(n, M, H, Fp)
s0 = (Wh4 * a[None, :, :, :Fp]).sum(-1).permute(1, 0, 2).reshape(M * n, H).contiguous()
t0 = (Wh4 * a[None, :, :, Fp:]).sum(-1).permute(1, 0, 2).reshape(M * n, H).contiguous()
y = torch.from_numpy(g["labels"]).to(DEV)
def tail(z):
    return torch.nn.functional.cross_entropy(model.predict(lay.semantic_attention(z.view(n, M, H * Fp))), y)
def run(batched):
    Wh, s, t = (x.clone().requires_grad_(True) for x in (Wh0, s0, t0))
    if batched:
        z = gat_aggregate(big, Wh, s, t, H, Fp, 0.2, elu=2, batch=M)
    else:
        z = torch.cat([gat_aggregate(graphs[m], Wh[:, m*64:(m+1)*64].contiguous(), s[m*n:(m+1)*n], t[m*n:(m+1)*n], H, Fp, 0.2, elu=2)
                       for m in range(M)], 1)
    zz = z.detach().clone().requires_grad_(True)
    tail(zz).backward()
    z.backward(zz.grad)
    return z.detach(), Wh.grad, s.grad, t.grad, zz.grad
B, U = run(True), run(False)
for name, i in (("z", 0), ("dWh", 1), ("ds", 2), ("dt", 3), ("dz", 4)):
    d = (B[i] - U[i]).abs()
    am = np.unravel_index(int(d.argmax()), d.shape)
    print(name, "max diff", float(d.max()), "ref max", float(U[i].abs().max()), "argmax", am, "vals", float(B[i][am]), float(U[i][am]))
for name, i in (("ds", 2), ("dt", 3)):
    d = (B[i] - U[i]).abs().view(M, n, H).amax(1) / U[i].abs().view(M, n, H).amax(1)
    print(name, d.cpu().numpy())
rp = big.rowptr.cpu()
i = int(np.unravel_index(int((B[3] - U[3]).abs().argmax()), B[3].shape)[0])
print("row", i, "deg", int(rp[i + 1] - rp[i]))

# float64 dense truth for metapath 1 given the same dz
m = 1
mask = masks[m] > 0
Whd = Wh0[:, m*64:(m+1)*64].double().requires_grad_(True)
sd = s0[m*n:(m+1)*n].double().requires_grad_(True)
td = t0[m*n:(m+1)*n].double().requires_grad_(True)
outs = []
for h in range(H):
    e = torch.nn.functional.leaky_relu(sd[:, h:h+1] + td[:, h:h+1].T, 0.2)
    att = torch.softmax(torch.where(mask, e, torch.full_like(e, -9e15)), dim=1)
    outs.append(att @ Whd[:, h*Fp:(h+1)*Fp])
zt = torch.nn.functional.elu(torch.nn.functional.elu(torch.cat(outs, 1)))
for tag, R in (("batched", B), ("unbatched", U)):
    for x in (Whd, sd, td):
        x.grad = None
    zt.backward(R[4][:, m*64:(m+1)*64].double(), retain_graph=True)
    print(tag, "z err", float((R[0][:, m*64:(m+1)*64] - zt).abs().max() / zt.abs().max()))
    for name, got, ref in (("dWh", R[1][:, m*64:(m+1)*64], Whd.grad), ("ds", R[2][m*n:(m+1)*n], sd.grad), ("dt", R[3][m*n:(m+1)*n], td.grad)):
        d = (got - ref).abs()
        per_head = (d.view(n, H, -1).amax((0, 2)) / ref.abs().view(n, H, -1).amax((0, 2))).cpu().numpy()
        print("   ", name, "max rel", float(d.max() / ref.abs().max()), "per head", np.array2string(per_head, precision=2))
