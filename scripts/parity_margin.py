"""How much of the 1e-4 contract each implementation uses: forward errors against the CPU oracle (A/B of kernel variants)."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from difusco_b200 import synthetic as syn
from oracle import difusco_oracle as orc
import gpu_util as G

w2 = syn.make_encoder_weights(0, out_channels=2)
W = orc.Weights(w2)


def errs(out, ref):
  out, ref = np.asarray(out, np.float64), np.asarray(ref, np.float64)
  p = torch.softmax(torch.as_tensor(out), -1).numpy()
  pr = torch.softmax(torch.as_tensor(ref), -1).numpy()
  return np.abs(out - ref).max() / np.abs(ref).max(), np.abs(p / pr - 1).max()


for (N, K, B, t) in ((500, 50, 2, 969.0), (200, 20, 3, 500.0), (1000, 100, 1, 31.0)):
  pts, ei = syn.tsp_sparse_batch(N, K, B, seed=N + B)
  xt = (syn.initial_noise(ei.shape[1], N) > 0).astype(np.float32)
  ref = orc.encoder_forward_sparse_tsp(W, pts, xt, np.array([t]), ei, gather_then_gemm=False).numpy()
  ref64 = orc.encoder_forward_sparse_tsp(orc.Weights(w2, dtype=torch.float64), pts, xt, np.array([t]), ei,
                                         gather_then_gemm=False).numpy()
  for impl in ("fp32", "tc1", "tc"):
    out = G.encoder(w2, 2, impl=impl)(G.cu(pts), torch.tensor([t]), G.cu(xt), G.cu(ei)).cpu().numpy()
    a, b = errs(out, ref)
    c, d = errs(out, ref64)
    print(f"TSP-{N} k={K} B={B} t={t:.0f} impl={impl:4s}: vs fp32 oracle logits {a:.2e} probs {b:.2e} | vs fp64 oracle logits {c:.2e} probs {d:.2e}",
          flush=True)
  a, b = errs(ref, ref64)
  print(f"   fp32 oracle vs fp64 oracle: logits {a:.2e} probs {b:.2e}", flush=True)
