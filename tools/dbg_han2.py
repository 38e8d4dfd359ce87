import sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from graphneuralnetwork_b200 import layers, synthetic as S
from conftest import rel_err, load_golden
DEV = "cuda"
g = load_golden("han_acm.npz")
n = S.ACM["n"]
masks = [torch.from_numpy(S.symmetric_mask(n, t, seed=11 + i)).to(DEV) for i, t in enumerate(S.ACM["metapath_nnz"])]
X = torch.from_numpy(np.random.default_rng(14).standard_normal((n, S.ACM["feats"]), dtype=np.float32)).to(DEV)
for batched in (True, False):
    model = layers.HANModel(3, S.ACM["feats"], 8, S.ACM["classes"], [8], 0.0)
    model.load_state_dict({k: torch.from_numpy(g[k]) for k in model.state_dict().keys()})
    model = model.to(DEV).train()
    model.layers[0].batched = batched
    out = model(masks, X)
    print(batched, "out", rel_err(out.detach().cpu().numpy(), g["out"]))
    torch.nn.functional.cross_entropy(out, torch.from_numpy(g["labels"]).to(DEV)).backward()
    worst = sorted(((rel_err(p.grad.cpu().numpy(), g["grad." + k]), k) for k, p in model.named_parameters()), reverse=True)[:6]
    for w in worst: print("   ", w)
