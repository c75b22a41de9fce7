"""GNNEncoder: drop-in for the reference's `models.gnn_encoder.GNNEncoder` at inference.

Same constructor signature, same parameter tree (so `state_dict()` keys and shapes are identical
and Lightning checkpoints load with `load_state_dict`), same `forward(x, timesteps, graph,
edge_index)` contract (difusco/models/gnn_encoder.py:290-462) - but `forward` does no PyTorch
math: it hands raw device pointers to the sm_100a CUDA library through the C-ABI
(include/difusco_b200.h).  There is no eager fallback: without the library or without a B200
the call raises.

Interface notes (reference behaviour kept):
  * sparse TSP  : forward(x (V,2), t (1,), graph=xt (E,), edge_index (2,E)) -> (E, out)    :383-402
  * MIS         : forward(xt (V,), t (1,), edge_index=(2,E))              -> (V, out)      :404-414
  * dense TSP   : forward(x (B,V,2), t (B,), graph=xt (B,V,V))            -> (B,out,V,V)   :350-381
    evaluated as the row-major complete graph incl. self pairs (gnn_encoder.py:365 makes the
    graph all-ones), GroupNorm per sample.
  * dense + node_feature_only raises NotImplementedError (:457), as in the reference.
Only inference is in scope: per-edge timestep vectors (training, pl_tsp_model.py:66-68) raise
NotImplementedError; autograd is not supported - outputs carry no grad_fn, and a call with grad mode
enabled on parameters that require grad warns once.
"""
import math

import numpy as np
import torch
from torch import nn

from .. import _cabi


def reference_frequency_tables(hidden_dim):
  """The three tiny frequency tables, evaluated with the reference's own torch expressions so
  the device uses bit-identical values (nn.py:113-116; gnn_encoder.py:216-217, :243-244)."""
  half = hidden_dim // 2
  freqs = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
  i = torch.arange(half, dtype=torch.float32)
  dimt_pos = 10000 ** (2.0 * (torch.div(i, 2, rounding_mode="trunc")) / half)
  i = torch.arange(hidden_dim, dtype=torch.float32)
  dimt_scalar = 10000 ** (2 * torch.div(i, 2, rounding_mode="trunc") / hidden_dim)
  return {"__const.time_freqs": freqs.numpy(), "__const.dimt_pos": dimt_pos.numpy(),
          "__const.dimt_scalar": dimt_scalar.numpy()}


class GNNLayer(nn.Module):
  """Parameter holder of one gated-GCN layer (gnn_encoder.py:20-65).  Compute is fused in CUDA."""

  def __init__(self, hidden_dim, aggregation="sum", norm="layer", learn_norm=True, track_norm=False,
               gated=True):
    super().__init__()
    if not gated:
      raise AssertionError("Use gating with GCN, pass the `--gated` flag")   # gnn_encoder.py:49
    if norm != "layer" or not learn_norm:
      raise NotImplementedError("difusco_b200 implements the reference default norm='layer' with affine")
    self.hidden_dim, self.aggregation = hidden_dim, aggregation
    for name in "UVABC":
      setattr(self, name, nn.Linear(hidden_dim, hidden_dim, bias=True))
    self.norm_h = nn.LayerNorm(hidden_dim, elementwise_affine=True)
    self.norm_e = nn.LayerNorm(hidden_dim, elementwise_affine=True)

  def forward(self, *a, **k):
    raise NotImplementedError("GNNLayer is fused into the edge-layer CUDA kernel; call GNNEncoder.forward")


class GNNEncoder(nn.Module):
  _warned_no_grad = False

  def __init__(self, n_layers, hidden_dim, out_channels=1, aggregation="sum", norm="layer",
               learn_norm=True, track_norm=False, gated=True,
               sparse=False, use_activation_checkpoint=False, node_feature_only=False,
               *args, **kwargs):
    super().__init__()
    self.sparse = sparse
    self.node_feature_only = node_feature_only
    self.hidden_dim = hidden_dim
    self.n_layers = n_layers
    self.out_channels = out_channels
    self.aggregation = aggregation
    self.use_activation_checkpoint = use_activation_checkpoint   # training-only memory knob: ignored
    ted = hidden_dim // 2
    self.node_embed = nn.Linear(hidden_dim, hidden_dim)
    self.edge_embed = nn.Linear(hidden_dim, hidden_dim)
    self.time_embed = nn.Sequential(nn.Linear(hidden_dim, ted), nn.ReLU(), nn.Linear(ted, ted))
    self.out = nn.Sequential(nn.GroupNorm(32, hidden_dim), nn.ReLU(),
                             nn.Conv2d(hidden_dim, out_channels, kernel_size=1, bias=True))
    self.layers = nn.ModuleList([GNNLayer(hidden_dim, aggregation, norm, learn_norm, track_norm, gated)
                                 for _ in range(n_layers)])
    self.time_embed_layers = nn.ModuleList([nn.Sequential(nn.ReLU(), nn.Linear(ted, hidden_dim))
                                            for _ in range(n_layers)])
    self.per_layer_out = nn.ModuleList([
        nn.Sequential(nn.LayerNorm(hidden_dim, elementwise_affine=learn_norm), nn.SiLU(),
                      nn.Linear(hidden_dim, hidden_dim)) for _ in range(n_layers)])
    for seq in self.per_layer_out:            # zero_module (gnn_encoder.py:343-345, nn.py:68-74)
      for p in seq[2].parameters():
        p.detach().zero_()
    self._ctx = None
    self._weights_key = None
    self._graph_key = None
    self._points_key = None
    self._complete_cache = {}

  # ------------------------------------------------------------------------------------------
  # engine plumbing
  # ------------------------------------------------------------------------------------------
  def _device(self):
    dev = self.node_embed.weight.device
    if dev.type != "cuda":
      raise RuntimeError("difusco_b200.GNNEncoder runs on a CUDA device only (no CPU fallback); "
                         "move the module with .cuda() first")
    return dev

  def engine(self):
    dev = self._device()
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if self._ctx is None or self._ctx.device != idx:
      self._ctx = _cabi.Context(idx)
      _cabi.device_context(idx, prefer=self._ctx)   # k-NN / 2-opt helpers share the first model's context
      self._ctx.set_aggregation(self.aggregation)
      self._weights_key = self._graph_key = self._points_key = None
    self._sync_weights()
    return self._ctx

  def _sync_weights(self):
    key = tuple((p.data_ptr(), p._version) for p in self.parameters())
    if key == self._weights_key:
      return
    sd = {k: v.detach().float().cpu().numpy() for k, v in self.state_dict().items()}
    self._ctx.load_weights(sd, self.n_layers, self.hidden_dim, self.out_channels, self.node_feature_only,
                           consts=reference_frequency_tables(self.hidden_dim))
    self._weights_key = key
    self._graph_key = self._points_key = None

  @staticmethod
  def _stream():
    return torch.cuda.current_stream().cuda_stream

  def set_graph(self, edge_index, num_nodes, gn_segments=1):
    """Prepare (and cache) the graph of subsequent calls.  edge_index (2,E) int64, any device."""
    ctx = self.engine()
    ei = edge_index.long().contiguous()
    key = (ei.data_ptr(), tuple(ei.shape), ei._version, int(num_nodes), int(gn_segments), str(ei.device))
    if key != self._graph_key:
      if ei.dim() != 2 or ei.shape[0] != 2:
        raise ValueError("edge_index must have shape (2, E)")
      ctx.prepare_graph(ei.data_ptr(), int(num_nodes), int(ei.shape[1]), int(gn_segments), self._stream())
      self._graph_key = key
      self._graph_hold = ei
      self._points_key = None
    return ctx

  def set_points(self, points):
    ctx = self.engine()
    p = points.float().contiguous()
    key = (p.data_ptr(), tuple(p.shape), p._version, self._graph_key)
    if key != self._points_key:
      ctx.set_points(p.data_ptr(), self._stream())
      self._points_key = key
      self._points_hold = p
    return ctx

  def _complete_graph(self, B, V, device):
    key = (B, V, str(device))
    if key not in self._complete_cache:
      r = torch.arange(V, device=device).repeat_interleave(V)
      c = torch.arange(V, device=device).repeat(V)
      off = (torch.arange(B, device=device) * V).repeat_interleave(V * V)
      self._complete_cache[key] = torch.stack([r.repeat(B) + off, c.repeat(B) + off]).long().contiguous()
    return self._complete_cache[key]

  @staticmethod
  def _single_t(timesteps):
    t = timesteps.reshape(-1).float()
    if t.numel() != 1 and not bool((t == t[0]).all()):
      raise NotImplementedError("per-element timesteps (the training path) are outside difusco_b200's scope")
    return float(t[0])

  # ------------------------------------------------------------------------------------------
  # forward variants (gnn_encoder.py:350-462)
  # ------------------------------------------------------------------------------------------
  def sparse_forward(self, x, graph, timesteps, edge_index):
    V, E = x.shape[0], edge_index.shape[1]
    ctx = self.set_graph(edge_index, V, 1)
    self.set_points(x.to(self._device()))
    xt = graph.reshape(-1).float().contiguous().to(self._device())
    out = torch.empty((E, self.out_channels), device=self._device(), dtype=torch.float32)
    ctx.encoder_forward(xt.data_ptr(), self._single_t(timesteps), out.data_ptr(), self._stream())
    return out

  def sparse_forward_node_feature_only(self, x, timesteps, edge_index):
    V = x.shape[0]
    ctx = self.set_graph(edge_index, V, 1)
    xt = x.reshape(-1).float().contiguous().to(self._device())
    out = torch.empty((V, self.out_channels), device=self._device(), dtype=torch.float32)
    ctx.encoder_forward(xt.data_ptr(), self._single_t(timesteps), out.data_ptr(), self._stream())
    return out

  def dense_forward(self, x, graph, timesteps, edge_index=None):
    del edge_index
    B, V, _ = x.shape
    dev = self._device()
    t = timesteps.reshape(-1).float()
    same_t = t.numel() == 1 or bool((t == t[0]).all())
    out = torch.empty((B, self.out_channels, V, V), device=dev, dtype=torch.float32)
    if same_t:          # one block-diagonal call, GroupNorm per sample (gn_segments = B)
      ctx = self.set_graph(self._complete_graph(B, V, dev), B * V, B)
      self.set_points(x.reshape(B * V, 2).to(dev))
      xt = graph.reshape(-1).float().contiguous().to(dev)
      flat = torch.empty((B * V * V, self.out_channels), device=dev, dtype=torch.float32)
      ctx.encoder_forward(xt.data_ptr(), float(t[0]), flat.data_ptr(), self._stream())
      out.copy_(flat.reshape(B, V, V, self.out_channels).permute(0, 3, 1, 2))
    else:               # different timestep per sample (dense_forward allows it, :375)
      for b in range(B):
        out[b:b + 1] = self.dense_forward(x[b:b + 1], graph[b:b + 1], t[b:b + 1])
    return out

  def forward(self, x, timesteps, graph=None, edge_index=None):
    if torch.is_grad_enabled() and not GNNEncoder._warned_no_grad and any(p.requires_grad for p in self.parameters()):
      # inference engine: never builds an autograd graph; make misuse visible (once)
      import warnings
      warnings.warn("difusco_b200.GNNEncoder is an inference engine: outputs carry no grad_fn "
                    "(call it under torch.no_grad(); training is outside this package's scope)")
      GNNEncoder._warned_no_grad = True
    if self.node_feature_only:
      if self.sparse:
        return self.sparse_forward_node_feature_only(x, timesteps, edge_index)
      raise NotImplementedError
    if self.sparse:
      return self.sparse_forward(x, graph, timesteps, edge_index)
    return self.dense_forward(x, graph, timesteps, edge_index)
