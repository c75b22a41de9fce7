"""CPU oracle of the TSP decode rows.  TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing
in difusco_b200/).  numpy restatement of

  greedy_merge       utils/tsp_utils.py:89-145 + utils/cython_merge/cython_merge.pyx:19-120
  two_opt            utils/tsp_utils.py:12-49
  tour_length        utils/tsp_utils.py:148-156

Pinned: tests/test_tsp_decode.py checks every function against tests/golden/tsp_decode.npz, which
tests/golden/make_golden.py produced by running the reference's own functions (Cython merge compiled from the
reference's .pyx, batched_two_opt_torch on the CPU device).
"""
import numpy as np
import scipy.sparse


def symmetric_heat(n, heat, edge_index=None):
  """tsp_utils.py:99-110: float32 matrix A + A^T (dense heat) or coo(h,(r,c)) + coo(h,(c,r)) (sparse)."""
  if edge_index is None:
    return heat + heat.T
  r, c = edge_index
  return (scipy.sparse.coo_matrix((heat, (r, c)), shape=(n, n)).toarray() +
          scipy.sparse.coo_matrix((heat, (c, r)), shape=(n, n)).toarray())


def greedy_merge(points, sym):
  """cython_merge.pyx:19-104 -> (tour of n+1 nodes per tsp_utils.py:133-141, merge_iterations)."""
  n = points.shape[0]
  p = points.astype("double")
  with np.errstate(divide="ignore", invalid="ignore"):
    order = np.argsort((-sym.astype("double") / np.linalg.norm(p[:, None] - p, axis=-1)).flatten())
  frag = list(range(n))          # fragment id of each node
  nbrs = [[] for _ in range(n)]
  members = {i: [i] for i in range(n)}
  merged = iterations = 0
  for flat in order:
    iterations += 1
    i, j = int(flat) // n, int(flat) % n
    if frag[i] == frag[j] or len(nbrs[i]) == 2 or len(nbrs[j]) == 2:
      continue
    nbrs[i].append(j)
    nbrs[j].append(i)
    fi, fj = frag[i], frag[j]
    for v in members[fi]:
      frag[v] = fj
    members[fj] += members.pop(fi)
    merged += 1
    if merged == n - 1:
      break
  a, b = [v for v in range(n) if len(nbrs[v]) < 2]
  nbrs[a].append(b)
  nbrs[b].append(a)
  tour = [0]
  while len(tour) < n + 1:
    cand = [v for v in nbrs[tour[-1]] if len(tour) == 1 or v != tour[-2]]
    tour.append(max(cand))
  return tour, iterations


def two_opt(points, tours, max_iterations):
  """tsp_utils.py:12-49 in float64 numpy: every tour applies its own best move while the batch-wide best move
  improves by more than 1e-6."""
  pts = np.asarray(points, dtype=np.float64)
  tours = np.array(tours, dtype=np.int64)
  n = pts.shape[0]
  iterations = 0
  while True:
    head, nxt = pts[tours[:, :-1]], pts[tours[:, 1:]]                     # (B, n, 2)
    d = lambda u, v: np.sqrt(np.sum((u - v) ** 2, axis=-1))
    change = (d(head[:, :, None], head[:, None, :]) + d(nxt[:, :, None], nxt[:, None, :])
              - d(head, nxt)[:, :, None] - d(head, nxt)[:, None, :])
    change = np.triu(change, k=2)
    flat = change.reshape(len(tours), -1)
    pick = flat.argmin(axis=1)
    if not flat.min() < -1e-6:
      break
    for b, idx in enumerate(pick):
      i, j = idx // n, idx % n
      tours[b, i + 1:j + 1] = tours[b, i + 1:j + 1][::-1].copy()
    iterations += 1
    if iterations >= max_iterations:
      break
  return tours, iterations


def tour_length(points, route):
  import warnings
  import scipy.spatial
  with warnings.catch_warnings():
    warnings.simplefilter("ignore", DeprecationWarning)
    dm = scipy.spatial.distance_matrix(points, points)
  total = 0
  for a, b in zip(route[:-1], route[1:]):
    total += dm[a, b]
  return total
