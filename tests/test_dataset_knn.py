"""TSP line parsing + k-NN graph construction.  The oracle for the graph is the reference's own
algorithm, sklearn's KDTree(leaf_size=30, euclidean).query on float64 points (co_datasets/tsp_graph_dataset.py:56-57)."""
import os
import tempfile

import numpy as np
import pytest
import torch

from difusco_b200.co_datasets.tsp_graph_dataset import TSPGraphDataset, knn_edge_index_gpu


def _write(tmp, pts, tours):
  f = os.path.join(tmp, "tsp.txt")
  with open(f, "w") as fh:
    for p, t in zip(pts, tours):
      fh.write(" ".join(f"{float(x)!r} {float(y)!r}" for x, y in p) + " output " + " ".join(str(i + 1) for i in t) + "\n")
  return f


def test_line_parser_and_dense_item_cpu():
  rng = np.random.default_rng(0)
  pts = [rng.random((7, 2)) for _ in range(3)]
  tours = [np.r_[rng.permutation(7), 0] for _ in range(3)]
  for t in tours:
    t[-1] = t[0]
  with tempfile.TemporaryDirectory() as tmp:
    ds = TSPGraphDataset(_write(tmp, pts, tours), sparse_factor=-1)
    assert len(ds) == 3
    for i in range(3):
      p, t = ds.get_example(i)
      assert np.array_equal(p, pts[i]) and np.array_equal(t, tours[i])     # repr() round-trips float64 exactly
      idx, pt, adj, tour = ds[i]
      assert idx.tolist() == [i] and pt.dtype == torch.float32 and adj.shape == (7, 7)
      assert adj.sum() == 7 and all(adj[tours[i][j], tours[i][j + 1]] == 1 for j in range(7))


@pytest.mark.gpu
@pytest.mark.parametrize("n,k", [(50, 5), (500, 50), (1000, 100), (2000, 50), (10000, 50)])
def test_knn_graph_matches_kdtree(n, k):
  from sklearn.neighbors import KDTree
  pts = np.random.default_rng(n + k).random((n, 2))
  _, ref = KDTree(pts, leaf_size=30, metric="euclidean").query(pts, k=k, return_distance=True)
  ei = knn_edge_index_gpu(pts, k).cpu().numpy()
  assert ei.shape == (2, n * k)
  assert np.array_equal(ei[0], np.repeat(np.arange(n), k))
  assert np.array_equal(ei[1].reshape(n, k), ref), "neighbour indices differ from the reference's KDTree query"
  off = knn_edge_index_gpu(pts, k, node_offset=7 * n).cpu().numpy()
  assert np.array_equal(off, ei + 7 * n)


@pytest.mark.gpu
def test_sparse_item_layout_matches_reference_semantics():
  from sklearn.neighbors import KDTree
  rng = np.random.default_rng(3)
  n, k = 40, 6
  pts = [rng.random((n, 2))]
  tours = [np.r_[rng.permutation(n), 0]]
  tours[0][-1] = tours[0][0]
  with tempfile.TemporaryDirectory() as tmp:
    ds = TSPGraphDataset(_write(tmp, pts, tours), sparse_factor=k)
    idx, graph, pind, eind, tour = ds[0]
  assert pind.tolist() == [n] and eind.tolist() == [n * k] and graph.x.dtype == torch.float32
  # the item is made of CPU tensors like the reference's (DataLoader pin_memory / the model's own .to(device) work)
  assert graph.edge_index.device.type == "cpu" and graph.edge_attr.device.type == "cpu" and graph.x.device.type == "cpu"
  _, ref = KDTree(pts[0], leaf_size=30, metric="euclidean").query(pts[0], k=k, return_distance=True)
  assert np.array_equal(graph.edge_index[1].cpu().numpy().reshape(n, k), ref)
  succ = np.zeros(n, dtype=np.int64)
  succ[tours[0][:-1]] = tours[0][1:]
  want = (ref == succ[:, None]).reshape(-1, 1)
  assert np.array_equal(graph.edge_attr.cpu().numpy(), want)
