"""Secondary timings of the other BASELINE.json configs (not the headline bench): full fused denoise loop on
synthetic instances of each shape, device-resident inputs, CUDA events.  Prints one JSON line per config."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from difusco_b200 import synthetic as syn
import gpu_util as G

def timed(fn, reps=2):
  fn(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps

def tsp(name, N, K, B, diffusion, P=1):
  w = syn.make_encoder_weights(0 if diffusion == "categorical" else 1, out_channels=2 if diffusion == "categorical" else 1)
  m = G.tsp_model(w, "tc", diffusion_type=diffusion, sparse_factor=K, parallel_sampling=P, inference_diffusion_steps=50)
  if P > 1:
    pts1 = syn.tsp_points(N, 1, 0); ei1 = torch.from_numpy(syn.knn_edge_index(pts1, K))
    ei = m.duplicate_edge_index(ei1, N, torch.device("cpu")).numpy(); pts = np.tile(pts1, (P, 1))
  else:
    pts, ei = syn.tsp_sparse_batch(N, K, B, seed=3)
  z = syn.initial_noise(ei.shape[1], 1)
  xt = (z > 0).astype(np.float32) if diffusion == "categorical" else z
  d_pts, d_ei, d_xt = G.cu(pts), G.cu(ei), G.cu(xt)
  ms = timed(lambda: m.denoise_heatmap(d_pts, d_ei, d_xt, seed=1))
  n = B * P
  print(json.dumps({"config": name, "V": int(pts.shape[0]), "E": int(ei.shape[1]), "instances": n, "ms_per_50step_batch": ms,
                    "graphs_per_s": n / (ms / 1e3)}), flush=True)

def mis(name, B):
  w = syn.make_encoder_weights(0, out_channels=2)
  m = G.mis_model(w, "tc", inference_diffusion_steps=50)
  ei, sizes = syn.mis_batch(700, 800, 0.15, B, seed=2)
  V = sum(sizes)
  xt = (syn.initial_noise(V, 2) > 0).astype(np.float32)
  d_ei, d_xt = G.cu(ei), G.cu(xt)
  ms = timed(lambda: m.denoise_labels(d_ei, d_xt, seed=1))
  print(json.dumps({"config": name, "V": V, "E": int(ei.shape[1]), "instances": B, "ms_per_50step_batch": ms,
                    "graphs_per_s": B / (ms / 1e3)}), flush=True)

tsp("C2 TSP-500 k=50 categorical B=16", 500, 50, 16, "categorical")
tsp("C2' TSP-500 k=50 categorical B=1 (reference test loader batch)", 500, 50, 1, "categorical")
tsp("C3 TSP-1000 k=100 gaussian B=8", 1000, 100, 8, "gaussian")
mis("C4 MIS ER-[700,800] p=0.15 categorical B=32", 32)
tsp("C5 TSP-10000 k=50 categorical P=4 (per GPU)", 10000, 50, 1, "categorical", P=4)
