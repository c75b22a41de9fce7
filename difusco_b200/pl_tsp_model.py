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

  # ------------------------------------------------------------------------------------
  # test_step = unpack -> (sample, fused denoise loop, decode) x sequential_sampling -> metrics
  # ------------------------------------------------------------------------------------
  def _unpack(self, batch):
    """The two batch layouts of the reference's datasets (pl_tsp_model.py:158-171)."""
    if self.sparse:
      index, graph, node_counts, _, gt_tour = batch
      coords = graph.x.reshape((-1, 2))
      edges = graph.edge_index.reshape((2, -1))
      n_graphs = node_counts.shape[0]
      labels = graph.edge_attr.reshape((n_graphs, edges.shape[1] // n_graphs))
      return index, coords, edges, labels, gt_tour, coords.cpu().numpy(), edges.cpu().numpy()
    index, coords, labels, gt_tour = batch
    return index, coords, None, labels, gt_tour, coords.cpu().numpy()[0], None

  def _initial_noise(self, like, copies):
    """pl_tsp_model.py:186-199: the noise tensor is drawn twice when parallel sampling is on (only the second
    draw is used); kept so torch's generator advances exactly as in the reference."""
    noise = torch.randn_like(like.float())
    if copies > 1:
      noise = noise.repeat(copies, 1) if self.sparse else noise.repeat(copies, 1, 1)
      noise = torch.randn_like(noise)
    if self.diffusion_type != "gaussian":
      noise = (noise > 0).long()
    return (noise.reshape(-1) if self.sparse else noise).float()

  def _heatmap_to_numpy(self, xt):
    if self.diffusion_type == "gaussian":
      return xt.cpu().detach().numpy() * 0.5 + 0.5          # :219-220
    return xt.float().cpu().detach().numpy() + 1e-6         # :221-222

  def test_step(self, batch, batch_idx, split="test"):
    device = batch[-1].device
    index, coords, edges, labels, gt_tour, np_points, np_edge_index = self._unpack(batch)
    copies, rounds = self.args.parallel_sampling, self.args.sequential_sampling
    if copies > 1:
      coords = coords.repeat(copies, 1) if self.sparse else coords.repeat(copies, 1, 1)
      if self.sparse:
        edges = self.duplicate_edge_index(edges, np_points.shape[0], device)
    two_opt_cap = getattr(self.args, "two_opt_iterations", 1000)
    ns, merge_iterations = 0, 0
    heatmaps, refined = [], []
    for _ in range(rounds):
      heat = self._heatmap_to_numpy(self.denoise_heatmap(coords, edges, self._initial_noise(labels, copies)))
      heatmaps.append(heat)
      if getattr(self.args, "save_numpy_heatmap", False):
        self.run_save_numpy_heatmap(heat, np_points, index, split)
      tours, merge_iterations = merge_tours(heat, np_points, np_edge_index, sparse_graph=self.sparse,
                                            parallel_sampling=copies, exact=getattr(self.args, "exact_merge", True))
      better, ns = batched_two_opt_torch(np_points.astype("float64"), np.array(tours).astype("int64"),
                                         max_iterations=two_opt_cap, device=device)
      refined.append(better)
    refined = np.concatenate(refined, axis=0)
    scorer = TSPEvaluator(np_points)
    best = np.min([scorer.evaluate(refined[i]) for i in range(copies * rounds)])
    metrics = {f"{split}/gt_cost": scorer.evaluate(gt_tour.cpu().numpy().reshape(-1)),
               f"{split}/2opt_iterations": ns, f"{split}/merge_iterations": merge_iterations}
    for name, value in metrics.items():
      self.log(name, value, on_epoch=True, sync_dist=True)
    self.log(f"{split}/solved_cost", best, prog_bar=True, on_epoch=True, sync_dist=True)
    # not part of the reference's return value: artefacts of the last call for callers that want them
    self.last_heatmap = heatmaps[0] if rounds == 1 else np.stack(heatmaps)
    self.last_solved_tours, self.last_solved_cost = refined, best
    return metrics

  def run_save_numpy_heatmap(self, adj_mat, np_points, real_batch_idx, split):
    """--save_numpy_heatmap (pl_tsp_model.py:258-267): <save_dir>/<name>/<version>/numpy_heatmap/{split}-heatmap-<idx>.npy
    and {split}-points-<idx>.npy, the input format of tsp_mcts/convert_numpy_to_txt.py."""
    if self.args.parallel_sampling > 1 or self.args.sequential_sampling > 1:
      raise NotImplementedError("Save numpy heatmap only support single sampling")
    logger = getattr(self, "logger", None)
    root = (os.path.join(logger.save_dir, logger.name, logger.version) if logger is not None
            else getattr(self.args, "storage_path", "."))
    target = os.path.join(root, "numpy_heatmap")
    os.makedirs(target, exist_ok=True)
    tag = real_batch_idx.cpu().numpy().reshape(-1)[0]
    np.save(os.path.join(target, f"{split}-heatmap-{tag}.npy"), adj_mat)
    np.save(os.path.join(target, f"{split}-points-{tag}.npy"), np_points)

  def validation_step(self, batch, batch_idx):
    return self.test_step(batch, batch_idx, split="val")
