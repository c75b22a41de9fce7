"""TSPModel: inference-side drop-in for the reference's difusco/pl_tsp_model.py.

Same method names and signatures on the denoise path:
  forward(x, adj, t, edge_index)                                              :38-39
  categorical_denoise_step(points, xt, t, device, edge_index=None, target_t=None)   :122-138
  gaussian_denoise_step(points, xt, t, device, edge_index=None, target_t=None)      :140-151
  test_step(batch, batch_idx, split='test')                                   :153-256
test_step runs the reference's loop (:185-222) as ONE fused device loop, then the reference's decode
(:227-256): merge_tours (host C++), batched 2-opt (CUDA), TSPEvaluator - and returns the
reference's metrics dict.  `--save_numpy_heatmap` (:224-225, :258-267) is honoured.
"""
import os

import numpy as np
import torch

from .pl_meta_model import COMetaModel
from .utils.tsp_utils import TSPEvaluator, batched_two_opt_torch, merge_tours


class TSPModel(COMetaModel):
  def __init__(self, param_args=None):
    super().__init__(param_args=param_args, node_feature_only=False)

  def forward(self, x, adj, t, edge_index):
    return self.model(x, t, adj, edge_index)

  # ------------------------------------------------------------------------------------
  def _prepare(self, points, edge_index, device):
    """Make self.model's engine hold this call's graph + coordinates; returns dense batch B or 0."""
    if self.sparse:
      V = points.shape[0]
      self.model.set_graph(edge_index.long().to(device), V, 1)
      self.model.set_points(points.float().to(device))
      return 0
    B, V, _ = points.shape
    self.model.set_graph(self.model._complete_graph(B, V, device), B * V, B)
    self.model.set_points(points.reshape(B * V, 2).float().to(device))
    return B

  def _denoise_step(self, points, xt, t, device, edge_index, target_t):
    with torch.no_grad():
      self._prepare(points, edge_index, device)
      shape = xt.shape
      out = self._fused_step(xt.float().to(device), t, target_t)
      return out.reshape(-1) if self.sparse else out.reshape(shape)

  def categorical_denoise_step(self, points, xt, t, device, edge_index=None, target_t=None):
    return self._denoise_step(points, xt, t, device, edge_index, target_t)

  def gaussian_denoise_step(self, points, xt, t, device, edge_index=None, target_t=None):
    return self._denoise_step(points, xt, t, device, edge_index, target_t)

  # ------------------------------------------------------------------------------------
  def denoise_heatmap(self, points, edge_index, xt, steps=None, seed=None):
    """xt0 -> raw final xt on device, the whole loop fused (no host sync inside)."""
    steps = steps or self.args.inference_diffusion_steps
    with torch.no_grad():
      dev = self.model._device()
      self._prepare(points.to(dev), edge_index.to(dev) if edge_index is not None else None, dev)
      x = xt.reshape(-1).float().contiguous().to(dev).clone()
      self._fused_loop(x, steps, seed)
      return x.reshape(xt.shape)

  def test_step(self, batch, batch_idx, split="test"):
    edge_index = None
    np_edge_index = None
    device = batch[-1].device
    if not self.sparse:
      real_batch_idx, points, adj_matrix, gt_tour = batch
      np_points = points.cpu().numpy()[0]
    else:
      real_batch_idx, graph_data, point_indicator, edge_indicator, gt_tour = batch
      points = graph_data.x.reshape((-1, 2))
      edge_index = graph_data.edge_index.reshape((2, -1))
      num_edges = edge_index.shape[1]
      batch_size = point_indicator.shape[0]
      adj_matrix = graph_data.edge_attr.reshape((batch_size, num_edges // batch_size))
      np_points = points.cpu().numpy()
      np_edge_index = edge_index.cpu().numpy()
    P = self.args.parallel_sampling
    if P > 1:
      if not self.sparse:
        points = points.repeat(P, 1, 1)
      else:
        points = points.repeat(P, 1)
        edge_index = self.duplicate_edge_index(edge_index, np_points.shape[0], device)
    np_gt_tour = gt_tour.cpu().numpy().reshape(-1)
    stacked_tours, heatmaps = [], []
    ns, merge_iterations = 0, 0
    for _ in range(self.args.sequential_sampling):
      xt = torch.randn_like(adj_matrix.float())
      if P > 1:
        xt = xt.repeat(P, 1, 1) if not self.sparse else xt.repeat(P, 1)
        xt = torch.randn_like(xt)
      if self.diffusion_type != "gaussian":
        xt = (xt > 0).long()
      if self.sparse:
        xt = xt.reshape(-1)
      xt = self.denoise_heatmap(points, edge_index, xt.float())
      if self.diffusion_type == "gaussian":
        adj_mat = xt.cpu().detach().numpy() * 0.5 + 0.5
      else:
        adj_mat = xt.float().cpu().detach().numpy() + 1e-6
      heatmaps.append(adj_mat)
      if getattr(self.args, "save_numpy_heatmap", False):
        self.run_save_numpy_heatmap(adj_mat, np_points, real_batch_idx, split)
      tours, merge_iterations = merge_tours(adj_mat, np_points, np_edge_index, sparse_graph=self.sparse,
                                            parallel_sampling=P,
                                            exact=getattr(self.args, "exact_merge", True))
      solved_tours, ns = batched_two_opt_torch(np_points.astype("float64"), np.array(tours).astype("int64"),
                                               max_iterations=getattr(self.args, "two_opt_iterations", 1000),
                                               device=device)
      stacked_tours.append(solved_tours)
    solved_tours = np.concatenate(stacked_tours, axis=0)
    tsp_solver = TSPEvaluator(np_points)
    gt_cost = tsp_solver.evaluate(np_gt_tour)
    all_solved_costs = [tsp_solver.evaluate(solved_tours[i]) for i in range(P * self.args.sequential_sampling)]
    best_solved_cost = np.min(all_solved_costs)
    metrics = {f"{split}/gt_cost": gt_cost, f"{split}/2opt_iterations": ns,
               f"{split}/merge_iterations": merge_iterations}
    for k, v in metrics.items():
      self.log(k, v, on_epoch=True, sync_dist=True)
    self.log(f"{split}/solved_cost", best_solved_cost, prog_bar=True, on_epoch=True, sync_dist=True)
    # not part of the reference's return value: kept for callers that want the artefacts of the last call
    self.last_heatmap = heatmaps[-1] if len(heatmaps) == 1 else np.stack(heatmaps)
    self.last_solved_tours, self.last_solved_cost = solved_tours, best_solved_cost
    return metrics

  def run_save_numpy_heatmap(self, adj_mat, np_points, real_batch_idx, split):
    """--save_numpy_heatmap (pl_tsp_model.py:258-267): <save_dir>/<name>/<version>/numpy_heatmap/{split}-heatmap-<idx>.npy
    and {split}-points-<idx>.npy, the input format of tsp_mcts/convert_numpy_to_txt.py."""
    if self.args.parallel_sampling > 1 or self.args.sequential_sampling > 1:
      raise NotImplementedError("Save numpy heatmap only support single sampling")
    logger = getattr(self, "logger", None)
    if logger is not None:
      exp_save_dir = os.path.join(logger.save_dir, logger.name, logger.version)
    else:
      exp_save_dir = getattr(self.args, "storage_path", ".")
    heatmap_path = os.path.join(exp_save_dir, "numpy_heatmap")
    os.makedirs(heatmap_path, exist_ok=True)
    real_batch_idx = real_batch_idx.cpu().numpy().reshape(-1)[0]
    np.save(os.path.join(heatmap_path, f"{split}-heatmap-{real_batch_idx}.npy"), adj_mat)
    np.save(os.path.join(heatmap_path, f"{split}-points-{real_batch_idx}.npy"), np_points)

  def validation_step(self, batch, batch_idx):
    return self.test_step(batch, batch_idx, split="val")
