"""Small forwards + a 3-step denoise for compute-sanitizer (memcheck): TSP categorical, MIS, Gaussian, both impls."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from difusco_b200 import synthetic as syn
import gpu_util as G
w2 = syn.make_encoder_weights(0, out_channels=2); w1 = syn.make_encoder_weights(1, out_channels=1)
for impl in ("tc", "fp32"):
  pts, ei = syn.tsp_sparse_batch(70, 9, 2, seed=1)          # E = 1260: partial last tile
  xt = (syn.initial_noise(ei.shape[1], 1) > 0).astype(np.float32)
  m = G.tsp_model(w2, impl, sparse_factor=9, inference_diffusion_steps=3)
  m.model(G.cu(pts), torch.tensor([500.0]), G.cu(xt), G.cu(ei))
  m.denoise_heatmap(G.cu(pts), G.cu(ei), G.cu(xt), seed=1)
  g = G.tsp_model(w1, impl, diffusion_type="gaussian", sparse_factor=9, inference_diffusion_steps=3)
  g.denoise_heatmap(G.cu(pts), G.cu(ei), G.cu(syn.initial_noise(ei.shape[1], 2)), seed=1)
  eim, sizes = syn.mis_batch(40, 60, 0.15, 2, seed=3)
  mm = G.mis_model(w2, impl, inference_diffusion_steps=3)
  mm.denoise_labels(G.cu(eim), G.cu((syn.initial_noise(sum(sizes), 3) > 0).astype(np.float32)), seed=1)
  torch.cuda.synchronize()
  print(impl, "done", flush=True)
