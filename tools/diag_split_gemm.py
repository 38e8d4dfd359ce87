"""Accuracy of the layer-0 product: fp32 SIMT path and the fp16x3 tensor-core path against float64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphneuralnetwork_b200 import functional as Fn, layers
dev = "cuda"
if os.environ.get("NO_REDUCED"):
    torch.backends.cuda.matmul.allow_fp16_reduced_precision_reduction = False
    torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = False
    print("reduced-precision reductions OFF")
# raw GEMM check: fp16 hi/lo operands, three products, vs float64
A = torch.randn(26624, 1208, device=dev); B = torch.randn(1208, 128, device=dev) * 0.05
Ah = A.half(); Al = (A - Ah.float()).half(); Bh = B.half(); Bl = (B - Bh.float()).half()
ref = A.double() @ B.double()
rel0 = lambda a: float((a.double() - ref).abs().max() / ref.abs().max())
print("raw GEMM: fp32 SIMT", f"{rel0(A @ B):.2e}",
      "| fp16x3", f"{rel0(torch.mm(Ah, Bh, out_dtype=torch.float32) + torch.mm(Al, Bh, out_dtype=torch.float32) + torch.mm(Ah, Bl, out_dtype=torch.float32)):.2e}",
      "| fp16 hi only", f"{rel0(torch.mm(Ah, Bh, out_dtype=torch.float32)):.2e}",
      "| operands exact? ", f"{rel0((Ah.double() + Al.double()) @ (Bh.double() + Bl.double())):.2e}")
g = torch.Generator().manual_seed(3)
for n, F_in, B, fan in ((5000, 602, 64, [7, 4]), (232965, 602, 1024, [25, 10])):
    table = Fn.pad_table(torch.randn(n, F_in, generator=g).to(dev))
    torch.manual_seed(1)
    model = layers.GraphSage(F_in, [128, 41], fan).to(dev).eval()
    sizes = [B, B * fan[0], B * fan[0] * fan[1]]
    blocks = [torch.randint(0, n, (s,), generator=g, dtype=torch.int32).to(dev) for s in sizes]
    with torch.no_grad():
        model.tensor_core_gemm = True
        fast = model.forward_sampled(table, blocks).double()
        h0_fast = [h.double() for h in model._layer0_one_launch_one_gemm(table, blocks)]
        model.tensor_core_gemm = False
        exact = model.forward_sampled(table, blocks).double()
        h0_exact = [h.double() for h in model._layer0_one_launch_one_gemm(table, blocks)]
        P = {k: v.double() for k, v in model.state_dict().items()}
        t64 = table.double()
        feats = [t64[b.long()] for b in blocks]

        def layer(l, src, neigh, act):  # SageGCN.py:23-36 in float64
            h = src @ P[f"gcn.{l}.weight"] + neigh.mean(1) @ P[f"gcn.{l}.aggregator.weight"]
            return torch.relu(h) if act else h

        h0_ref = [layer(0, feats[h], feats[h + 1].view(feats[h].shape[0], fan[h], -1), True) for h in range(2)]
        ref = layer(1, h0_ref[0], h0_ref[1].view(h0_ref[0].shape[0], fan[0], -1), False)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print(f"n={n}: logits  fp32-path vs f64 {rel(exact, ref):.2e} | fp16x3 vs f64 {rel(fast, ref):.2e} | fp16x3 vs fp32-path {rel(fast, exact):.2e}")
    for h in range(2):
        print(f"   layer0 hop{h}: fp32-path vs f64 {rel(h0_exact[h], h0_ref[h]):.2e} | fp16x3 vs f64 {rel(h0_fast[h], h0_ref[h]):.2e}")
