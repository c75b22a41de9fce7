"""CPU probe: end-to-end error of candidate split-precision schemes for the E-row GEMMs (C, O) and the node
linears, against the exact fp32 oracle.  Emulates operand rounding only (fp32 accumulate)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.nn.functional as F
from difusco_b200 import synthetic as syn
from oracle import difusco_oracle as orc

torch.set_grad_enabled(False)
def bf16(x): return x.to(torch.bfloat16).to(torch.float32)
def fp16(x): return x.to(torch.float16).to(torch.float32)

SCHEMES = {
  "exact":        lambda x, w: (x, w),
  "bf16x3":       None,   # a_hi*b_hi + a_lo*b_hi + a_hi*b_lo
  "fp16hi+bf16lo x W_fp16 (2 MMA)": None,
  "bf16hi+bf16lo x W_bf16 (2 MMA)": None,
  "fp16hi+fp16lo x W_fp16 (2 MMA)": None,
  "tf32x1": None,
}

def lin_emul(scheme, x, w, b):
  if scheme == "exact":
    return F.linear(x, w, b)
  if scheme == "bf16x3":
    xh = bf16(x); xl = bf16(x - xh); wh = bf16(w); wl = bf16(w - wh)
    return F.linear(xh, wh) + F.linear(xl, wh) + F.linear(xh, wl) + b
  if scheme.startswith("fp16hi+bf16lo"):
    xh = fp16(x); xl = bf16(x - xh); wh = fp16(w)
    return F.linear(xh, wh) + F.linear(xl, wh) + b
  if scheme.startswith("bf16hi+bf16lo"):
    xh = bf16(x); xl = bf16(x - xh); wh = bf16(w)
    return F.linear(xh, wh) + F.linear(xl, wh) + b
  if scheme.startswith("fp16hi+fp16lo"):
    xh = fp16(x); xl = fp16(x - xh); wh = fp16(w)
    return F.linear(xh, wh) + F.linear(xl, wh) + b
  if scheme == "tf32x1":
    def tf32(t): return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    return F.linear(tf32(x), tf32(w)) + b
  raise ValueError(scheme)

class W2(orc.Weights):
  scheme = "exact"
  def lin(self, name, x):
    parts = name.split(".")
    edge_or_node = (parts[0] == "layers" and parts[2] in "UVABC") or (parts[0] == "per_layer_out" and parts[2] == "2")
    if edge_or_node and self.scheme != "exact":
      return lin_emul(self.scheme, x, self.t[name + ".weight"], self.t[name + ".bias"])
    return super().lin(name, x)

N, K, B = int(os.environ.get("N", 200)), int(os.environ.get("K", 20)), int(os.environ.get("B", 2))
w = syn.make_encoder_weights(0, out_channels=2)
pts, ei = syn.tsp_sparse_batch(N, K, B, seed=5)
xt = (syn.initial_noise(ei.shape[1], 3) > 0).astype(np.float32)
ref64 = orc.encoder_forward_sparse_tsp(orc.Weights(w, torch.float64), pts, xt, np.array([500.0]), ei, gather_then_gemm=False)
pr64 = ref64.softmax(-1)
for sch in SCHEMES:
  ww = W2(w); ww.scheme = sch
  out = orc.encoder_forward_sparse_tsp(ww, pts, xt, np.array([500.0]), ei, gather_then_gemm=False).double()
  p = out.softmax(-1)
  print(f"{sch:40s} logits rel-Linf {float((out - ref64).abs().max() / ref64.abs().max()):.2e}   prob max-rel {float((p / pr64 - 1).abs().max()):.2e}")
