"""Bring-up diagnostics of the CTA-pair kernel's GEMM1 (structured inputs -> where does each element land)."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from difusco_b200 import synthetic as syn
import gpu_util as G

w = syn.make_encoder_weights(0, out_channels=2)
C = w["layers.1.C.weight"].astype(np.float64)   # [out n][in k]
impl = sys.argv[1] if len(sys.argv) > 1 else "tc"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 256
enc = G.encoder(w, 2, impl=impl)
V = 64
rng = np.random.default_rng(0)
ei = np.stack([np.sort(rng.integers(0, V, E)), rng.integers(0, V, E)]).astype(np.int64)
ctx = enc.set_graph(G.cu(ei), V, 1)
st = torch.cuda.current_stream().cuda_stream


def run(x):
  xin = G.cu(x.astype(np.float32))
  acc = torch.full((E, 256), float("nan"), device="cuda")
  try:
    ctx.debug_edge_gemm(1, xin.data_ptr(), acc.data_ptr(), st)
    torch.cuda.synchronize()
  except Exception as ex:
    print("LAUNCH FAILED:", ex, "watchdog", ctx.debug_watchdog(), flush=True)
    raise
  return acc.cpu().numpy().astype(np.float64)


for k0 in (0, 1, 7, 8, 16, 31, 32, 33, 100, 255):
  x = np.zeros((E, 256)); x[:, k0] = 1.0
  got = run(x)
  ref = np.tile(C[:, k0][None, :], (E, 1))
  err = np.abs(got - ref)
  bad_rows = np.where(err.max(1) > 1e-3)[0]
  bad_cols = np.where(err.max(0) > 1e-3)[0]
  print(f"one-hot k={k0}: max err {err.max():.3e}; bad rows {len(bad_rows)} (first {bad_rows[:6]}), bad cols {len(bad_cols)} (first {bad_cols[:6]})")
  if err.max() > 1e-3 and k0 in (0, 33):
    r = bad_rows[0]
    # does the output row match some OTHER k column of C, or another N permutation?
    cand = np.abs(C.T[None, :, :] - got[r][None, None, :]).max(-1).ravel()   # per k: distance to C[:,k]
    print("   row", r, "got[:8]", got[r, :8], "ref[:8]", ref[r, :8], "closest k", int(cand.argmin()), "dist", cand.min())
x = np.zeros((E, 256)); x[:, 5] = np.arange(E) % 97 + 1
got = run(x)
ref = x @ C.T
print("row-scaled one-hot: max rel err", np.abs(got - ref).max() / np.abs(ref).max())
x = rng.standard_normal((E, 256)) * 3
got = run(x)
ref = x.astype(np.float32).astype(np.float64) @ C.T
print("random: max rel err", np.abs(got - ref).max() / np.abs(ref).max(), "finite", np.isfinite(got).all())
