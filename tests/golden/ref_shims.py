"""Import shims that let the UNMODIFIED reference (/root/reference/difusco) run in this container.

Test infrastructure only (used by tests/golden/make_golden.py, which generates the committed
fixtures).  The reference needs four third-party packages that are not installed here and
cannot be installed (no network):

  torch_sparse        -> SparseTensor(row, col, value, sparse_sizes), .size(d), .to(), sum(st, dim=1)
                         used at gnn_encoder.py:13-16,177-191,417-423.  torch-sparse==0.6.15
                         (environment.yml:131) defines sum(dim=1) as the exact row-wise segmented
                         sum of `value`; here: zeros.index_add_(0, row, value).
  pytorch_lightning   -> LightningModule (a plain nn.Module with a no-op .log), rank_zero_info
  torch_geometric     -> data.Data / data.DataLoader as names only
  pickle5             -> stdlib pickle

Nothing from the reference is copied: these are stand-ins for its *dependencies*.
"""
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference/difusco"


def install(root=None):
  """root: directory holding the reference's `difusco/` sources (default /root/reference/difusco; bench.py passes
  oracle/_ref/difusco, the byte-for-byte copy that travels to the GPU box)."""
  global REFERENCE_ROOT
  if root is not None:
    REFERENCE_ROOT = root
  if "torch_sparse" in sys.modules and getattr(sys.modules["torch_sparse"], "_dfb_shim", False):
    if REFERENCE_ROOT not in sys.path:
      sys.path.insert(0, REFERENCE_ROOT)
    return

  ts = types.ModuleType("torch_sparse")
  ts._dfb_shim = True

  class SparseTensor(object):
    def __init__(self, row=None, col=None, value=None, sparse_sizes=None):
      self.row, self.col, self.value, self.sparse_sizes = row, col, value, sparse_sizes

    def size(self, d):
      return self.sparse_sizes[d]

    def to(self, *a, **k):
      return self

  def _sum(st, dim=1):
    assert dim == 1
    out = torch.zeros((st.sparse_sizes[0],) + tuple(st.value.shape[1:]), dtype=st.value.dtype,
                      device=st.value.device)
    return out.index_add_(0, st.row, st.value)

  def _mean(st, dim=1):
    s = _sum(st, dim)
    cnt = torch.zeros(st.sparse_sizes[0], dtype=st.value.dtype).index_add_(
        0, st.row, torch.ones_like(st.row, dtype=st.value.dtype))
    return s / cnt.clamp(min=1).unsqueeze(-1)

  def _max(st, dim=1):
    out = torch.full((st.sparse_sizes[0],) + tuple(st.value.shape[1:]), float("-inf"),
                     dtype=st.value.dtype)
    idx = st.row.unsqueeze(-1).expand_as(st.value)
    out = out.scatter_reduce(0, idx, st.value, reduce="amax", include_self=True)
    return torch.where(torch.isinf(out), torch.zeros_like(out), out)

  ts.SparseTensor, ts.sum, ts.mean, ts.max = SparseTensor, _sum, _mean, _max
  sys.modules["torch_sparse"] = ts

  pl = types.ModuleType("pytorch_lightning")
  pl._dfb_shim = True

  class LightningModule(torch.nn.Module):
    def log(self, *a, **k):
      pass

  pl.LightningModule = LightningModule
  plu = types.ModuleType("pytorch_lightning.utilities")
  plu.rank_zero_info = lambda *a, **k: None
  pl.utilities = plu
  sys.modules["pytorch_lightning"] = pl
  sys.modules["pytorch_lightning.utilities"] = plu

  tg = types.ModuleType("torch_geometric")
  tgd = types.ModuleType("torch_geometric.data")

  class Data(object):
    def __init__(self, **kw):
      self.__dict__.update(kw)

  tgd.Data = Data
  tgd.DataLoader = object
  tg.data = tgd
  sys.modules["torch_geometric"] = tg
  sys.modules["torch_geometric.data"] = tgd

  import pickle
  sys.modules["pickle5"] = pickle

  # utils/tsp_utils.py:9 imports the Cython tour-merge extension (decode, OUT of the denoise
  # path).  It is not built in the read-only reference tree; the golden generator never decodes.
  cm = types.ModuleType("utils.cython_merge.cython_merge")

  def merge_cython(*a, **k):
    raise RuntimeError("cython_merge is outside the denoise path and is not built here")

  cm.merge_cython = merge_cython
  sys.modules["utils.cython_merge.cython_merge"] = cm

  if REFERENCE_ROOT not in sys.path:
    sys.path.insert(0, REFERENCE_ROOT)
