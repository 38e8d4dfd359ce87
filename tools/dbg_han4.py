import torch
torch.manual_seed(0)
X = torch.randn(3025, 1870, device="cuda")
for w in (64, 192):
    D = torch.randn(3025, w, device="cuda") * 1e-4
    ref = X.double().T @ D.double()
    got = X.T @ D
    print("X^T D", w, float((got - ref).abs().max() / ref.abs().max()))
    W = torch.randn(1870, w, device="cuda") * 0.05
    ref = X.double() @ W.double(); got = X @ W
    print("X W", w, float((got - ref).abs().max() / ref.abs().max()))
    ref = D.double() @ W.double().T; got = D @ W.T
    print("D W^T", w, float((got - ref).abs().max() / ref.abs().max()))
print(torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision())
