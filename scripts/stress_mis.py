"""Stress the MIS path (variable degree, unsorted edges, last layer without GEMM2) to catch rare sync bugs."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from difusco_b200 import synthetic as syn
import gpu_util as G
B = int(os.environ.get("B", 32)); reps = int(os.environ.get("REPS", 6))
w = syn.make_encoder_weights(0, out_channels=2)
m = G.mis_model(w, "tc", inference_diffusion_steps=50)
ei, sizes = syn.mis_batch(700, 800, 0.15, B, seed=2)
V = sum(sizes)
xt = G.cu((syn.initial_noise(V, 2) > 0).astype(np.float32)); d_ei = G.cu(ei)
ctx = m.model.engine()
for r in range(reps):
  try:
    m.denoise_labels(d_ei, xt, seed=r)
    torch.cuda.synchronize()
    print("rep", r, "ok", ctx.debug_watchdog(), flush=True)
  except Exception as e:
    print("rep", r, "FAILED", type(e).__name__, str(e)[:120], "watchdog", ctx.debug_watchdog(), flush=True)
    break
