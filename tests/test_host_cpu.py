"""CPU-only checks of the host side: C-ABI surface, host arithmetic, interface mirror.  No GPU."""
import os
import re
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from conftest import ROOT, golden
from difusco_b200 import _cabi, build, synthetic as syn
from difusco_b200.models.gnn_encoder import GNNEncoder, reference_frequency_tables
from difusco_b200.pl_mis_model import MISModel
from difusco_b200.pl_tsp_model import TSPModel
from difusco_b200.utils.diffusion_schedulers import (CategoricalDiffusion, GaussianDiffusion,
                                                    InferenceSchedule)
from oracle import difusco_oracle as orc


def _args(**kw):
  a = dict(diffusion_type="categorical", diffusion_schedule="linear", diffusion_steps=1000, sparse_factor=50,
           n_layers=12, hidden_dim=256, aggregation="sum", parallel_sampling=1, sequential_sampling=1,
           inference_schedule="cosine", inference_diffusion_steps=50, inference_trick="ddim")
  a.update(kw)
  return NS(**a)


def test_library_builds_and_exports_every_declared_symbol():
  build.build()
  assert os.path.exists(_cabi.LIB_PATH)
  header = open(os.path.join(ROOT, "include", "difusco_b200.h")).read()
  declared = sorted(set(re.findall(r"\b(dfb_[a-z_0-9]+)\s*\(", header)))
  assert declared, "no declarations parsed"
  L = _cabi.lib()
  for name in declared:
    assert hasattr(L, name), f"{name} declared in include/difusco_b200.h but not exported"
  assert sorted(_cabi.SYMBOLS) == declared
  assert L.dfb_abi_version() == 2


def test_no_cpu_fallback_context_fails_loudly():
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  with pytest.raises(_cabi.DfbError, match="no CUDA device"):
    _cabi.Context(0)
  enc = GNNEncoder(12, 256, 2, sparse=True)
  with pytest.raises(RuntimeError, match="CUDA device only"):
    enc(torch.zeros(4, 2), torch.tensor([1.0]), torch.zeros(8), torch.zeros(2, 8, dtype=torch.long))


def test_scheduler_mirror_bit_exact():
  g = golden("schedules")
  for sch in ("linear", "cosine"):
    c, ga = CategoricalDiffusion(1000, sch), GaussianDiffusion(1000, sch)
    assert np.array_equal(c.Qs, g[f"cat_{sch}_Qs"]) and np.array_equal(c.Q_bar, g[f"cat_{sch}_Qbar"])
    assert np.array_equal(ga.beta, g[f"gau_{sch}_beta"]) and np.array_equal(ga.alpha, g[f"gau_{sch}_alpha"])
    assert np.array_equal(ga.alphabar, g[f"gau_{sch}_alphabar"])
    assert np.array_equal(ga.betabar, np.cumprod(ga.beta))
  for kind in ("linear", "cosine"):
    for steps in (50, 10, 1000):
      s = InferenceSchedule(kind, 1000, steps)
      got = np.array([[int(a), int(b)] for a, b in map(s, range(steps))])
      assert np.array_equal(got, g[f"infer_{kind}_{steps}"])
  with pytest.raises(ValueError):
    InferenceSchedule("quadratic", 1000, 10)(0)
  with pytest.raises(AssertionError):
    InferenceSchedule("linear", 1000, 10)(10)


def test_state_dict_keys_match_reference_order():
  for out, nfo in ((2, False), (1, False), (2, True)):
    enc = GNNEncoder(12, 256, out, sparse=True, node_feature_only=nfo)
    shapes = syn.encoder_param_shapes(12, 256, out)
    sd = enc.state_dict()
    assert list(sd.keys()) == list(shapes.keys())
    for k, v in sd.items():
      assert tuple(v.shape) == tuple(shapes[k]), k
    assert float(enc.per_layer_out[3][2].weight.detach().abs().max()) == 0.0   # zero_module
    w = syn.make_encoder_weights(0, out_channels=out)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    # Lightning checkpoints carry a "model." prefix on the task module
  m = TSPModel(_args())
  assert all(k.startswith("model.") for k in m.state_dict().keys())


def test_posterior_consts_match_oracle():
  m = TSPModel(_args())
  _, Qbar = orc.categorical_tables(1000, "linear")
  for (t, tt) in orc.inference_schedule("cosine", 1000, 50):
    c, last = m.posterior_consts(t, tt)
    assert np.array_equal(c.reshape(2, 2), orc.categorical_posterior_consts(Qbar, t, tt))
    assert last == int(tt == 0)
  g = TSPModel(_args(diffusion_type="gaussian"))
  beta, alpha, ab = orc.gaussian_tables(1000, "linear")
  for (t, tt) in orc.inference_schedule("cosine", 1000, 50):
    c, _ = g.posterior_consts(t, tt)
    kind, a, b1, x = orc.gaussian_posterior_consts(beta, alpha, ab, t, tt, "ddim")
    if kind == "ddim":
      assert np.array_equal(c, np.array([a, b1, x, 0.0], np.float32))
    else:
      assert np.array_equal(c, np.array([a, b1, 0.0, x], np.float32)) and c[3] == 0.0   # t = 1: noise coeff is 0


def test_reference_signature_posteriors_cpu():
  """categorical_posterior / gaussian_posterior keep the reference call convention (numpy (1,) ints)."""
  m = TSPModel(_args())
  _, Qbar = orc.categorical_tables(1000, "linear")
  rng = np.random.default_rng(1)
  p0 = torch.from_numpy(rng.random((1, 5, 4, 2)).astype(np.float32))
  p0 = p0 / p0.sum(-1, keepdim=True)
  xt = torch.from_numpy((rng.random(20) > 0.5).astype(np.int64))
  out = m.categorical_posterior(np.array([0]), torch.tensor([31]), p0, xt)   # last step: deterministic
  ref_p, ref_next = orc.categorical_posterior(Qbar, 31, 0, p0.reshape(20, 2), xt.float())
  assert torch.allclose(out, ref_next, atol=1e-7)
  g = TSPModel(_args(diffusion_type="gaussian"))
  beta, alpha, ab = orc.gaussian_tables(1000, "linear")
  pred, x = torch.randn(20), torch.randn(20)
  out = g.gaussian_posterior(np.array([469]), torch.tensor([500]), pred, x)
  assert torch.allclose(out, orc.gaussian_posterior(beta, alpha, ab, 500, 469, pred, x), atol=1e-6)
  with pytest.raises(ValueError):
    TSPModel(_args(diffusion_type="poisson"))


def test_duplicate_edge_index():
  m = MISModel(_args(parallel_sampling=3, sparse_factor=-1))
  ei = torch.tensor([[0, 1, 2], [1, 2, 0]])
  out = m.duplicate_edge_index(ei, 3, torch.device("cpu"))
  assert out.tolist() == [[0, 1, 2, 3, 4, 5, 6, 7, 8], [1, 2, 0, 4, 5, 3, 7, 8, 6]]


def test_frequency_tables_are_the_reference_expressions():
  t = reference_frequency_tables(256)
  x = torch.tensor([0.3, 0.7])
  assert torch.equal(orc._dim_t(128, torch.float32), torch.from_numpy(t["__const.dimt_pos"]))
  assert torch.equal(orc._dim_t(256, torch.float32), torch.from_numpy(t["__const.dimt_scalar"]))
  te = orc.timestep_embedding(torch.tensor([517.0]), 256, torch.float32)
  f = torch.from_numpy(t["__const.time_freqs"])
  assert torch.equal(te[0, :128], torch.cos(517.0 * f))


def test_mis_decode_matches_reference_function():
  """utils/mis_utils.py drop-in vs outputs of the reference's mis_decode_np (tests/golden/mis_decode.npz)."""
  import scipy.sparse
  from difusco_b200.utils.mis_utils import mis_decode_np
  g = golden("mis_decode")
  for case in range(3):
    ei, pred = g[f"ei{case}"], g[f"pred{case}"]
    adj = scipy.sparse.coo_matrix((np.ones_like(ei[0]), (ei[0], ei[1])))
    sol = mis_decode_np(pred, adj)
    assert np.array_equal(sol, g[f"sol{case}"])
    sel = np.flatnonzero(sol)                       # independence: no edge between two selected nodes (self loops aside)
    a = adj.tocsr()
    for i in sel:
      nb = a.indices[a.indptr[i]:a.indptr[i + 1]]
      assert not np.any(sol[nb[nb != i]])


def test_mcts_text_export_matches_reference_converter(tmp_path):
  """tsp_mcts/convert_numpy_to_txt.py drop-in: byte-identical files to the reference script's (tests/golden/mcts_txt.npz)."""
  from difusco_b200.tsp_mcts import convert_numpy_to_txt as conv
  g = golden("mcts_txt")
  for case in range(2):
    heat, pts, prob = g[f"heat{case}"], g[f"points{case}"], float(g[f"prob{case}"])
    n = len(pts)
    src = tmp_path / f"in{case}" / "numpy_heatmap"
    src.mkdir(parents=True)
    np.save(src / "test-heatmap-0.npy", heat)
    np.save(src / "test-points-0.npy", pts)
    (path,) = conv.main(str(tmp_path / f"in{case}"), str(tmp_path / f"out{case}"), num_nodes=n, num_files=1,
                        expected_valid_prob=prob)
    assert path.endswith(f"heatmap/tsp{n}/heatmaptsp{n}_0.txt")
    with open(path, "rb") as fh:
      assert fh.read() == g[f"txt{case}"].tobytes()
  with pytest.raises(ValueError):
    from difusco_b200 import _cabi
    _cabi.write_heatmap_txt(str(tmp_path / "missing_dir" / "x.txt"), np.zeros((2, 2)))


def test_mis_decode_yields_maximal_independent_sets():
  """Properties of the greedy decode on random graphs (isolated nodes, ties): independent, maximal, and the first
  node in score order is always selected."""
  import scipy.sparse
  from difusco_b200.utils.mis_utils import mis_decode_np
  rng = np.random.default_rng(4)
  for case in range(25):
    n = int(rng.integers(2, 80))
    m = int(rng.integers(0, 4 * n))
    r, c = rng.integers(0, n, m), rng.integers(0, n, m)
    keep = r != c
    r, c = np.concatenate([r[keep], c[keep]]), np.concatenate([c[keep], r[keep]])     # undirected, as the datasets store it
    adj = scipy.sparse.coo_matrix((np.ones(len(r)), (r, c)), shape=(n, n)).tocsr()
    score = rng.random(n) if case % 2 else np.round(rng.random(n) * 3) / 3
    sol = mis_decode_np(score, adj)
    chosen = np.flatnonzero(sol)
    assert sol.shape == (n,) and set(np.unique(sol)) <= {0, 1}
    assert sol[np.argsort(-score)[0]] == 1
    assert adj[chosen][:, chosen].nnz == 0                                             # independent
    covered = np.asarray(adj[:, chosen].sum(axis=1)).reshape(-1) > 0
    assert np.all(covered | (sol == 1))                                                # maximal


def test_load_from_checkpoint_reads_lightning_checkpoints(tmp_path):
  """train.py:127 `model_class.load_from_checkpoint(ckpt_path, param_args=args)`: a Lightning checkpoint written for
  the reference module (keys 'model.<GNNEncoder key>') loads into the drop-in unchanged."""
  from types import SimpleNamespace as NS
  from difusco_b200 import synthetic as syn
  from difusco_b200.pl_tsp_model import TSPModel
  args = NS(diffusion_type="categorical", diffusion_schedule="linear", diffusion_steps=1000, sparse_factor=50, n_layers=12,
            hidden_dim=256, aggregation="sum", parallel_sampling=1, sequential_sampling=1, inference_schedule="cosine",
            inference_diffusion_steps=50, inference_trick="ddim")
  w = syn.make_encoder_weights(seed=0, out_channels=2)
  ckpt = {"state_dict": {"model." + k: torch.from_numpy(v.copy()) for k, v in w.items()}, "epoch": 3,
          "hyper_parameters": {"param_args": None}}
  path = tmp_path / "last.ckpt"
  torch.save(ckpt, path)
  m = TSPModel.load_from_checkpoint(str(path), param_args=args)
  got = m.state_dict()
  assert set(got) == set(ckpt["state_dict"])
  for k, v in ckpt["state_dict"].items():
    assert torch.equal(got[k], v), k
