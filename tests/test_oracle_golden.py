"""Pin the CPU oracle (oracle/difusco_oracle.py) against outputs of the reference itself
(tests/golden/*.npz, made by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_linf
from oracle import difusco_oracle as orc
from difusco_b200 import synthetic as syn

TOL = 1e-5   # fp32 restatement vs fp32 reference: same library kernels, reassociation only


def test_schedule_tables_exact():
  g = golden("schedules")
  for sch in ("linear", "cosine"):
    Qs, Qbar = orc.categorical_tables(1000, sch)
    assert np.array_equal(Qs, g[f"cat_{sch}_Qs"])
    assert np.array_equal(Qbar, g[f"cat_{sch}_Qbar"])
    beta, alpha, alphabar = orc.gaussian_tables(1000, sch)
    assert np.array_equal(beta, g[f"gau_{sch}_beta"])
    assert np.array_equal(alpha, g[f"gau_{sch}_alpha"])
    assert np.array_equal(alphabar, g[f"gau_{sch}_alphabar"])


@pytest.mark.parametrize("kind", ["linear", "cosine"])
@pytest.mark.parametrize("steps", [50, 10, 1000])
def test_inference_schedule_exact(kind, steps):
  g = golden("schedules")
  assert np.array_equal(np.array(orc.inference_schedule(kind, 1000, steps)), g[f"infer_{kind}_{steps}"])


def test_forward_tsp_categorical(weights2):
  g = golden("fwd_tsp_cat")
  w = orc.Weights(weights2)
  taps = []
  out = orc.encoder_forward_sparse_tsp(w, g["points"], g["xt"], g["t"], g["edge_index"], taps=taps)
  assert rel_linf(out.numpy(), g["logits"]) < TOL
  # per_layer_out output = e_new - e_in; compare the residual branch at layer 0 and 11
  e_prev = orc.Weights(weights2).lin("edge_embed", orc.scalar_embed(torch.from_numpy(g["xt"]), 256))
  assert rel_linf((taps[0][1] - e_prev).numpy()[::7], g["out_layer0_rows"]) < 1e-5
  assert rel_linf((taps[11][1] - taps[10][1]).numpy()[::7], g["out_layer11_rows"]) < 1e-4
  # fp64 arbiter agrees with the reference to fp32 accuracy
  out64 = orc.encoder_forward_sparse_tsp(orc.Weights(weights2, torch.float64), g["points"], g["xt"],
                                         g["t"], g["edge_index"])
  assert rel_linf(out64.numpy(), g["logits"]) < 2e-5


def test_forward_tsp_gaussian(weights1):
  g = golden("fwd_tsp_gauss")
  out = orc.encoder_forward_sparse_tsp(orc.Weights(weights1), g["points"], g["xt"], g["t"], g["edge_index"])
  assert out.shape == g["pred"].shape
  assert rel_linf(out.numpy(), g["pred"]) < TOL


@pytest.mark.parametrize("agg", ["sum", "mean", "max"])
def test_forward_mis(weights2, agg):
  g = golden("fwd_mis_cat")
  ref = g["logits"] if agg == "sum" else golden(f"fwd_mis_cat_{agg}")["logits"]
  out = orc.encoder_forward_mis(orc.Weights(weights2), g["xt"], g["t"], g["edge_index"], aggregation=agg)
  assert rel_linf(out.numpy(), ref) < TOL


def test_forward_dense(weights2):
  g = golden("fwd_dense_cat")
  out = orc.encoder_forward_dense(orc.Weights(weights2), g["points"], g["xt"], g["t"])
  assert out.shape == g["out"].shape
  assert rel_linf(out.numpy(), g["out"]) < TOL


def test_dense_equals_sparse_complete_graph(weights2):
  """Dense (B=1) == sparse on the row-major complete graph incl. self pairs (SURVEY 3.3 probe):
  the mapping the CUDA path uses for the dense API."""
  g = golden("fwd_dense_cat")
  w = orc.Weights(weights2)
  V = g["points"].shape[1]
  ei = syn.complete_edge_index(V)
  for b in range(2):
    d = orc.encoder_forward_dense(w, g["points"][b:b + 1], g["xt"][b:b + 1], g["t"][b:b + 1])
    s = orc.encoder_forward_sparse_tsp(w, g["points"][b], g["xt"][b].reshape(-1), g["t"][b:b + 1], ei)
    assert rel_linf(s.t().reshape(1, 2, V, V).numpy(), d.numpy()) < 5e-6


def _check_traj(name, task, dtype_w, diffusion, useed, weights, points_key="points", sched="cosine"):
  g = golden(name)
  w = orc.Weights(weights)
  V, K, P, steps = [int(x) for x in g["meta"]]
  points = None
  if task == "tsp":
    points = np.tile(g[points_key], (P, 1))
  ei = g["edge_index"]
  xt0 = g["xt0"].astype(np.float32)
  # teacher-forced: every step gets the reference's own xt_in
  forced = [xt0] + [g["xt_out"][i] for i in range(steps - 1)]
  us = [syn.uniforms(ei.shape[1] if task == "tsp" else xt0.shape[0], useed, i) for i in range(steps)]
  rec = []
  final = orc.denoise(w, task, diffusion, ei, xt0, points=points, steps=steps, uniforms=us,
                      inference_schedule_kind=sched, forced_xt=forced, record=rec)
  for i, r in enumerate(rec):
    assert rel_linf(r["net_out"].numpy(), g["net_out"][i]) < 5e-6, (name, i)
    if diffusion == "categorical" and i < steps - 1:
      assert np.abs(r["p"].numpy().clip(0, 1) - g["p"][i]).max() < 2e-6
    # sampled state: identical except where |p-u| is within fp32 noise
    if diffusion == "categorical" and i < steps - 1:
      flips = (r["xt_out"].numpy() != g["xt_out"][i])
      near = np.abs(g["p"][i] - us[i]) < 1e-5
      assert not np.any(flips & ~near)
  assert rel_linf(final.numpy(), g["xt_out"][-1]) < 5e-6
  # free-running from xt0 reproduces the reference trajectory too (tiny graph: no near-ties expected)
  free = orc.denoise(w, task, diffusion, ei, xt0, points=points, steps=steps, uniforms=us,
                     inference_schedule_kind=sched)
  assert rel_linf(free.numpy(), g["xt_out"][-1]) < 1e-4


def test_traj_tsp_categorical(weights2):
  _check_traj("traj_tsp_cat", "tsp", None, "categorical", 100, weights2)


def test_traj_tsp_gaussian(weights1):
  _check_traj("traj_tsp_gauss", "tsp", None, "gaussian", 101, weights1)


def test_traj_mis_categorical(weights2):
  _check_traj("traj_mis_cat", "mis", None, "categorical", 102, weights2)


def test_traj_mis_gaussian(weights1):
  _check_traj("traj_mis_gauss", "mis", None, "gaussian", 103, weights1, sched="linear")


def test_traj_dense_categorical(weights2):
  """Dense reference trajectory reproduced through the sparse complete-graph oracle (B=1)."""
  g = golden("traj_dense_cat")
  V, _, _, steps = [int(x) for x in g["meta"]]
  w = orc.Weights(weights2)
  ei = syn.complete_edge_index(V)
  xt0 = g["xt0"].reshape(-1).astype(np.float32)
  us = [syn.uniforms(V * V, 104, i) for i in range(steps)]
  forced = [xt0] + [g["xt_out"][i].reshape(-1) for i in range(steps - 1)]
  rec = []
  final = orc.denoise(w, "tsp", "categorical", ei, xt0, points=g["points"], steps=steps, uniforms=us,
                      forced_xt=forced, record=rec)
  for i, r in enumerate(rec):
    ref = g["net_out"][i]            # (1, 2, V, V)
    assert rel_linf(r["net_out"].t().reshape(1, 2, V, V).numpy(), ref) < 1e-5
  assert rel_linf(final.numpy(), g["xt_out"][-1].reshape(-1)) < 1e-5


def test_posterior_consts_closed_form():
  """c[x][k] reproduces the one-hot matmul chain of pl_meta_model.py:113-137 on random p0."""
  _, Qbar = orc.categorical_tables(1000, "linear")
  rng = np.random.default_rng(0)
  for (t, tt) in [(1000, 969), (500, 469), (31, 1), (1, 0)]:
    p0 = torch.from_numpy(rng.random((64, 2)).astype(np.float32))
    p0 = p0 / p0.sum(-1, keepdim=True)
    xt = torch.from_numpy((rng.random(64) > 0.5).astype(np.float32))
    p, nxt = orc.categorical_posterior(Qbar, t, tt, p0, xt, u=rng.random(64).astype(np.float32))
    # literal evaluation (float64) of the textbook posterior q(x_{t'}=1 | x_t, x_0) summed over x_0
    Q = np.linalg.inv(Qbar[tt]) @ Qbar[t]
    lit = np.zeros(64)
    for n in range(64):
      x = int(xt[n])
      lit[n] = sum(Q[1, x] * Qbar[tt][k, 1] / Qbar[t][k, x] * float(p0[n, k]) for k in (0, 1))
    assert np.abs(p.numpy() - lit).max() < 1e-6
    if tt == 0:
      assert np.array_equal(nxt.numpy(), p.clamp(min=0).numpy())
