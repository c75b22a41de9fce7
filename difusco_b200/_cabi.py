"""ctypes binding of include/difusco_b200.h.  The ONLY compute backend: if the shared library is
missing or no B200 is present every call fails loudly - there is no eager/CPU fallback."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdifusco_b200.so")

DFB_OK, DFB_E_INVALID, DFB_E_CUDA, DFB_E_UNSUPPORTED, DFB_E_NOMEM = 0, -1, -2, -3, -4
CATEGORICAL, GAUSSIAN = 0, 1
EDGE_IMPL_TC, EDGE_IMPL_FP32, EDGE_IMPL_TC1 = 0, 1, 2
AGGREGATION = {"sum": 0, "mean": 1, "max": 2}

# every symbol include/difusco_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "dfb_abi_version", "dfb_create", "dfb_destroy", "dfb_last_error", "dfb_set_aggregation",
    "dfb_set_edge_impl", "dfb_load_weights", "dfb_prepare_graph", "dfb_set_points",
    "dfb_encoder_forward", "dfb_denoise_step", "dfb_denoise", "dfb_denoise_host",
    "dfb_launch_count", "dfb_profile_begin", "dfb_profile_end", "dfb_debug_edge_gemm",
    "dfb_debug_phase_cycles", "dfb_debug_watchdog", "dfb_knn_graph", "dfb_set_graph_capture",
    "dfb_tsp_merge_sparse", "dfb_tsp_merge_order", "dfb_two_opt", "dfb_write_heatmap_txt",
]

_lib = None


def lib():
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError(
        f"{LIB_PATH} not found: build it with `python -m difusco_b200.build` "
        "(difusco_b200 has no fallback path; the CUDA library is the product)")
  L = C.CDLL(LIB_PATH)
  vp, i32, i64, u64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float
  L.dfb_abi_version.restype = i32
  L.dfb_create.argtypes = [C.POINTER(vp), i32]
  L.dfb_destroy.argtypes = [vp]
  L.dfb_last_error.argtypes = [vp]
  L.dfb_last_error.restype = C.c_char_p
  L.dfb_set_aggregation.argtypes = [vp, i32]
  L.dfb_set_edge_impl.argtypes = [vp, i32]
  L.dfb_load_weights.argtypes = [vp, i32, i32, i32, i32, i32, C.POINTER(C.c_char_p), C.POINTER(vp),
                                 C.POINTER(i64)]
  L.dfb_prepare_graph.argtypes = [vp, vp, i64, i64, i32, vp]
  L.dfb_set_points.argtypes = [vp, vp, vp]
  L.dfb_encoder_forward.argtypes = [vp, vp, f32, vp, vp]
  L.dfb_denoise_step.argtypes = [vp, i32, vp, f32, C.POINTER(f32), i32, vp, u64, i32, vp, vp, vp, vp]
  L.dfb_denoise.argtypes = [vp, i32, vp, i32, C.POINTER(C.c_int32), C.POINTER(f32),
                            C.POINTER(C.c_int32), vp, u64, vp]
  L.dfb_denoise_host.argtypes = [vp, i32, vp, vp, i64, i64, i32, vp, i32, C.POINTER(C.c_int32),
                                 C.POINTER(f32), C.POINTER(C.c_int32), u64, vp, vp]
  L.dfb_launch_count.argtypes = [vp]
  L.dfb_launch_count.restype = i64
  L.dfb_profile_begin.argtypes = [vp]
  L.dfb_profile_end.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
  L.dfb_debug_edge_gemm.argtypes = [vp, i32, vp, vp, vp]
  L.dfb_debug_phase_cycles.argtypes = [vp, C.POINTER(C.c_uint64)]
  L.dfb_debug_watchdog.argtypes = [vp, C.POINTER(C.c_int)]
  L.dfb_set_graph_capture.argtypes = [vp, i32]
  L.dfb_knn_graph.argtypes = [vp, vp, i64, i32, i64, vp, vp]
  L.dfb_tsp_merge_sparse.argtypes = [vp, i64, vp, vp, i64, i32, vp, C.POINTER(i64)]
  L.dfb_tsp_merge_order.argtypes = [i64, vp, i64, vp, C.POINTER(i64)]
  L.dfb_two_opt.argtypes = [vp, vp, i64, vp, i64, i64, C.POINTER(i64), vp]
  L.dfb_write_heatmap_txt.argtypes = [C.c_char_p, i64, vp]
  for name in SYMBOLS:
    fn = getattr(L, name)
    if fn.restype is C.c_int and name not in ("dfb_abi_version",):
      fn.restype = i32
  _lib = L
  return L


class DfbError(RuntimeError):
  pass


def _raise(code, msg):
  msg = msg.decode() if isinstance(msg, bytes) else msg
  if code == DFB_E_INVALID:
    raise ValueError(msg)
  if code == DFB_E_UNSUPPORTED:
    raise NotImplementedError(msg)
  if code == DFB_E_NOMEM:
    raise MemoryError(msg)
  raise DfbError(msg)


MERGE_COMPLETE, MERGE_INCOMPLETE, MERGE_AMBIGUOUS = 0, 1, 2


def tsp_merge_sparse(points, heat, edge_index, mode=0):
  """dfb_tsp_merge_sparse on host arrays -> (status, tour (n+1,) int64, merge_iterations)."""
  points = np.ascontiguousarray(points, dtype=np.float64)
  heat = np.ascontiguousarray(heat, dtype=np.float32).reshape(-1)
  edge_index = np.ascontiguousarray(edge_index, dtype=np.int64)
  n = points.shape[0]
  if edge_index.ndim != 2 or edge_index.shape[0] != 2 or edge_index.shape[1] != heat.shape[0]:
    raise ValueError("edge_index must be (2, E) with one heat value per edge")
  tour = np.empty(n + 1, dtype=np.int64)
  it = C.c_int64(0)
  rc = lib().dfb_tsp_merge_sparse(points.ctypes.data, n, heat.ctypes.data, edge_index.ctypes.data, heat.shape[0],
                                  int(mode), tour.ctypes.data, C.byref(it))
  if rc < 0:
    _raise(rc, "dfb_tsp_merge_sparse: invalid argument (n >= 3, indices inside [0, n), mode 0/1)")
  return rc, tour, it.value


def tsp_merge_order(n, order):
  """dfb_tsp_merge_order: the reference loop over an explicit order of flattened (i*n + j) entries."""
  order = np.ascontiguousarray(order, dtype=np.int64).reshape(-1)
  tour = np.empty(n + 1, dtype=np.int64)
  it = C.c_int64(0)
  rc = lib().dfb_tsp_merge_order(int(n), order.ctypes.data, order.shape[0], tour.ctypes.data, C.byref(it))
  if rc < 0:
    _raise(rc, "dfb_tsp_merge_order: invalid argument or the order does not complete a tour")
  return tour, it.value


def write_heatmap_txt(path, matrix):
  """dfb_write_heatmap_txt: (n, n) matrix -> the tsp_mcts text format.  float32 input is widened exactly."""
  m = np.ascontiguousarray(matrix, dtype=np.float64)
  if m.ndim != 2 or m.shape[0] != m.shape[1]:
    raise ValueError("matrix must be square")
  rc = lib().dfb_write_heatmap_txt(os.fsencode(path), m.shape[0], m.ctypes.data)
  if rc != DFB_OK:
    _raise(rc, f"dfb_write_heatmap_txt: cannot write {path}")


_DEVICE_CTX = {}


def device_context(device_index, prefer=None):
  """The ONE dfb_ctx per device that the helpers around the path (k-NN graph, 2-opt) run on.  A model registers its
  own context with `prefer=` when it is created first, so model + k-NN + 2-opt share one context (and one set of
  kernel attributes / scratch buffers) per GPU; without a model a plain context is created on first use."""
  ctx = _DEVICE_CTX.get(device_index)
  if ctx is None or getattr(ctx, "_h", None) is None:
    ctx = prefer if prefer is not None else Context(device_index)
    _DEVICE_CTX[device_index] = ctx
  return ctx


class Context(object):
  """One dfb_ctx: one GPU, one model, one prepared graph at a time."""

  def __init__(self, device=0):
    L = lib()
    h = C.c_void_p()
    rc = L.dfb_create(C.byref(h), int(device))
    if rc != DFB_OK:
      _raise(rc, L.dfb_last_error(None))
    self._h = h
    self.device = int(device)
    self._keep = []

  def close(self):
    if getattr(self, "_h", None):
      lib().dfb_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  def _ck(self, rc):
    if rc != DFB_OK:
      _raise(rc, lib().dfb_last_error(self._h))

  # ---- model ----
  def load_weights(self, state_dict, n_layers, hidden_dim, out_channels, node_feature_only, consts=None):
    """state_dict: name -> numpy fp32 array (GNNEncoder.state_dict() keys, optional 'model.' prefix)."""
    items = [(k, np.ascontiguousarray(v, dtype=np.float32)) for k, v in state_dict.items()]
    for k, v in (consts or {}).items():
      items.append((k, np.ascontiguousarray(v, dtype=np.float32)))
    n = len(items)
    names = (C.c_char_p * n)(*[k.encode() for k, _ in items])
    ptrs = (C.c_void_p * n)(*[v.ctypes.data for _, v in items])
    numels = (C.c_int64 * n)(*[v.size for _, v in items])
    self._ck(lib().dfb_load_weights(self._h, n_layers, hidden_dim, out_channels, int(bool(node_feature_only)),
                                    n, names, ptrs, numels))

  def set_aggregation(self, name):
    if name not in AGGREGATION:
      raise ValueError(f"unknown aggregation {name}")
    self._ck(lib().dfb_set_aggregation(self._h, AGGREGATION[name]))

  def set_edge_impl(self, impl):
    self._ck(lib().dfb_set_edge_impl(self._h, impl))

  # ---- graph ----
  def prepare_graph(self, edge_index_ptr, num_nodes, num_edges, gn_segments=1, stream=0):
    self._ck(lib().dfb_prepare_graph(self._h, edge_index_ptr, num_nodes, num_edges, gn_segments, stream))

  def set_points(self, points_ptr, stream=0):
    self._ck(lib().dfb_set_points(self._h, points_ptr, stream))

  # ---- compute ----
  def encoder_forward(self, xt_ptr, t, out_ptr, stream=0):
    self._ck(lib().dfb_encoder_forward(self._h, xt_ptr, float(t), out_ptr, stream))

  def denoise_step(self, diffusion, xt_in_ptr, t, consts, last, uniforms_ptr, seed, step_index, xt_out_ptr,
                   p_out_ptr=None, net_out_ptr=None, stream=0):
    c = (C.c_float * 4)(*[float(x) for x in consts])
    self._ck(lib().dfb_denoise_step(self._h, diffusion, xt_in_ptr, float(t), c, int(last), uniforms_ptr,
                                    int(seed) & 0xFFFFFFFFFFFFFFFF, int(step_index), xt_out_ptr, p_out_ptr,
                                    net_out_ptr, stream))

  @staticmethod
  def _sched_arrays(t1, consts, last):
    steps = len(t1)
    t1a = (C.c_int32 * steps)(*[int(x) for x in t1])
    ca = (C.c_float * (4 * steps))(*[float(x) for row in consts for x in row])
    la = (C.c_int32 * steps)(*[int(x) for x in last])
    return steps, t1a, ca, la

  def denoise(self, diffusion, xt_ptr, t1, consts, last, uniforms_ptr=None, seed=0, stream=0):
    steps, t1a, ca, la = self._sched_arrays(t1, consts, last)
    self._ck(lib().dfb_denoise(self._h, diffusion, xt_ptr, steps, t1a, ca, la, uniforms_ptr,
                               int(seed) & 0xFFFFFFFFFFFFFFFF, stream))

  def denoise_host(self, diffusion, points_ptr, edge_index_ptr, num_nodes, num_edges, gn_segments, xt0_ptr,
                   t1, consts, last, seed, heatmap_ptr, stream=0):
    steps, t1a, ca, la = self._sched_arrays(t1, consts, last)
    self._ck(lib().dfb_denoise_host(self._h, diffusion, points_ptr, edge_index_ptr, num_nodes, num_edges,
                                    gn_segments, xt0_ptr, steps, t1a, ca, la,
                                    int(seed) & 0xFFFFFFFFFFFFFFFF, heatmap_ptr, stream))

  # ---- the step before the path ----
  def two_opt(self, points, tours, max_iterations, stream=0):
    """dfb_two_opt on host arrays: returns (tours (B, n+1) int64 copy, iterations)."""
    points = np.ascontiguousarray(points, dtype=np.float64)
    tours = np.array(tours, dtype=np.int64, order="C", copy=True)
    if tours.ndim != 2 or tours.shape[1] != points.shape[0] + 1:
      raise ValueError("tours must be (batch, n + 1)")
    it = C.c_int64(0)
    self._ck(lib().dfb_two_opt(self._h, points.ctypes.data, points.shape[0], tours.ctypes.data, tours.shape[0],
                               int(max_iterations), C.byref(it), stream))
    return tours, it.value

  def knn_graph(self, points_ptr, num_nodes, k, node_offset, edge_index_ptr, stream=0):
    self._ck(lib().dfb_knn_graph(self._h, points_ptr, int(num_nodes), int(k), int(node_offset), edge_index_ptr, stream))

  # ---- accounting ----
  def launch_count(self):
    return int(lib().dfb_launch_count(self._h))

  def profile_begin(self):
    self._ck(lib().dfb_profile_begin(self._h))

  def profile_end(self):
    ms, n = C.c_double(0), C.c_int64(0)
    self._ck(lib().dfb_profile_end(self._h, C.byref(ms), C.byref(n)))
    return ms.value, n.value

  def set_graph_capture(self, enabled):
    self._ck(lib().dfb_set_graph_capture(self._h, int(bool(enabled))))

  def debug_watchdog(self):
    out = (C.c_int * 4)()
    lib().dfb_debug_watchdog(self._h, out)
    return [int(x) for x in out]

  def debug_phase_cycles(self):
    out = (C.c_uint64 * 32)()
    self._ck(lib().dfb_debug_phase_cycles(self._h, out))
    return [int(x) for x in out]

  def debug_edge_gemm(self, layer, e_in_ptr, acc_out_ptr, stream=0):
    self._ck(lib().dfb_debug_edge_gemm(self._h, layer, e_in_ptr, acc_out_ptr, stream))
