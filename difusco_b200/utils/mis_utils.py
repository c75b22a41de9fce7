"""MIS greedy decode - drop-in for the reference's utils/mis_utils.py (the step right after the
MIS denoise path: pl_mis_model.py:194-196).

mis_decode_np(predictions (V,) float, adj_matrix scipy.sparse (V,V)) -> (V,) int {0,1}:
visit nodes by descending score; a node not yet excluded joins the set and excludes all its neighbours
(its own self-loop entry too, then it is re-marked as selected) - mis_utils.py:3-18.
Implemented on the CSR arrays directly (no per-row sparse slicing); the visiting order is the same
`np.argsort(-predictions)` call, so ties break identically.
"""
import numpy as np


def mis_decode_np(predictions, adj_matrix):
  csr = adj_matrix.tocsr()
  indptr, indices = csr.indptr, csr.indices
  state = np.zeros(predictions.shape[0], dtype=np.int64)     # 0 undecided, 1 selected, -1 excluded
  for node in np.argsort(-predictions):
    if state[node] == -1:
      continue
    state[indices[indptr[node]:indptr[node + 1]]] = -1
    state[node] = 1
  return (state == 1).astype(int)
