"""TSPModel: inference-side drop-in for the reference's difusco/pl_tsp_model.py.

Same method names and signatures on the denoise path:
  forward(x, adj, t, edge_index)                                              :38-39
  categorical_denoise_step(points, xt, t, device, edge_index=None, target_t=None)   :122-138
  gaussian_denoise_step(points, xt, t, device, edge_index=None, target_t=None)      :140-151
  test_step(batch, batch_idx, split='test')                                   :153-256
test_step runs the reference's loop (:185-222) as ONE fused device loop and returns the heatmap;
tour decoding (merge_tours / 2-opt, :227-237) is the next row outside this path: plug a callable
into `self.decoder(adj_mat, np_points, np_edge_index) -> dict` to get solved-cost metrics.
"""
import numpy as np
import torch

from .pl_meta_model import COMetaModel


class TSPModel(COMetaModel):
  def __init__(self, param_args=None):
    super().__init__(param_args=param_args, node_feature_only=False)
    self.decoder = None

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
    heatmaps = []
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
    metrics = {f"{split}/heatmap": heatmaps[-1] if len(heatmaps) == 1 else np.stack(heatmaps)}
    if self.decoder is not None:
      metrics.update(self.decoder(heatmaps, np_points, np_edge_index, gt_tour))
    return metrics

  def validation_step(self, batch, batch_idx):
    return self.test_step(batch, batch_idx, split="val")
