"""Heat-map export for the MCTS solver: drop-in for the reference's tsp_mcts/convert_numpy_to_txt.py (SURVEY 8f row
f4).  Reads `<heatmap_dir>/numpy_heatmap/test-<prefix>-<i>.npy` + `test-points-<i>.npy` (what
TSPModel.run_save_numpy_heatmap writes, pl_tsp_model.py:258-267) and writes
`<output_dir>/<prefix>/tsp<N>/heatmaptsp<N>_<i>.txt`: first line N, then N rows of N values "%.6f" separated by
blanks - the format tsp_mcts/code/TSP_IO.h:461-492 parses.

    python -m difusco_b200.tsp_mcts.convert_numpy_to_txt --heatmap_dir D --output_dir O --num_nodes 10000 \\
        --num_files 16 --expected_valid_prob 0.02 --heatmap_prefix heatmap

Same arithmetic as the reference (float32 throughout for float32 inputs; convert_numpy_to_txt.py:21-45); the N*N text
rows are written by the C library (dfb_write_heatmap_txt) instead of one Python f-string per entry.
"""
import argparse
import os

import numpy as np

from .. import _cabi


def sparsify_heatmap(heat, points, num_nodes, expected_valid_prob):
  """(N, N) heat map -> symmetric, row-normalised sparse prior (convert_numpy_to_txt.py:21-45):
  distance bonus, keep the top `expected_valid_prob` fraction of positive entries plus every row's 3 largest,
  +0.01 on kept entries, symmetrise, normalise rows."""
  gaps = np.linalg.norm(points[:, None, :] - points[None, :, :], axis=-1)
  prior = heat + 0.01 * (1.0 - gaps)
  prior[prior == np.inf] = 0.0
  keep = int(num_nodes * num_nodes * expected_valid_prob)
  positive = np.sort(prior[prior > 0.0])
  threshold = positive[-keep]
  strongest = np.argsort(prior, axis=1)[:, -3:]
  kept = prior > threshold
  kept[np.arange(num_nodes)[:, None], strongest] = True
  prior = prior * kept
  prior[prior != 0.0] += 1e-2
  prior = prior + prior.T
  return prior / prior.sum(axis=1, keepdims=True)


def write_heatmap_txt(path, matrix):
  _cabi.write_heatmap_txt(path, matrix)


def main(heatmap_dir, output_dir, num_nodes=10000, num_files=16, expected_valid_prob=0.02, heatmap_prefix="heatmap"):
  written = []
  for i in range(num_files):
    heat = np.load(f"{heatmap_dir}/numpy_heatmap/test-{heatmap_prefix}-{i}.npy")
    points = np.load(f"{heatmap_dir}/numpy_heatmap/test-points-{i}.npy")
    prior = sparsify_heatmap(heat, points, num_nodes, expected_valid_prob)
    folder = f"{output_dir}/{heatmap_prefix}/tsp{num_nodes}"
    os.makedirs(folder, exist_ok=True)
    path = f"{folder}/heatmaptsp{num_nodes}_{i}.txt"
    write_heatmap_txt(path, prior)
    nz = prior > 0.0
    print(f"{path}: valid_prob {nz.mean():.6f}  per-node min/max {nz.sum(axis=1).min()}/{nz.sum(axis=1).max()}")
    written.append(path)
  return written


if __name__ == "__main__":
  ap = argparse.ArgumentParser()
  ap.add_argument("--heatmap_dir", required=True)
  ap.add_argument("--output_dir", required=True)
  ap.add_argument("--num_nodes", type=int, default=10000)
  ap.add_argument("--num_files", type=int, default=16)
  ap.add_argument("--expected_valid_prob", type=float, default=0.02)
  ap.add_argument("--heatmap_prefix", default="heatmap")
  main(**vars(ap.parse_args()))
