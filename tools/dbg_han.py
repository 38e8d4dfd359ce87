import sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from graphneuralnetwork_b200 import layers, synthetic as S
DEV = "cuda"
n = S.ACM["n"]
masks = [torch.from_numpy(S.symmetric_mask(n, t, seed=11 + i)).to(DEV) for i, t in enumerate(S.ACM["metapath_nnz"])]
X = torch.from_numpy(np.random.default_rng(14).standard_normal((n, 256), dtype=np.float32)).to(DEV)
torch.manual_seed(0)
layer = layers.HANLayer(3, 256, 8, 8, 0.0).to(DEV)
res = {}
for batched in (True, False):
    layer.batched = batched
    layer.zero_grad()
    x = X.clone().requires_grad_(True)
    out = layer(masks, x)
    out.square().sum().backward()
    res[batched] = (out.detach(), x.grad.clone(), {k: p.grad.clone() for k, p in layer.named_parameters()})
def rel(a, b): return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
print("out", rel(res[True][0], res[False][0]))
print("dx", rel(res[True][1], res[False][1]))
for k in res[True][2]:
    print(k, rel(res[True][2][k], res[False][2][k]))
