"""Timing probes of the fused edge-layer kernel on the BASELINE config[1] shape (not a bench)."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from difusco_b200 import synthetic as syn
from difusco_b200 import _cabi
if os.environ.get("DFB_LIB"):          # A/B of tuning builds: point the binding at another build of the library
  _cabi.LIB_PATH = os.environ["DFB_LIB"]
import gpu_util as G

B = int(os.environ.get("PROBE_B", "16"))
w = syn.make_encoder_weights(0, out_channels=2)
pts, ei = syn.tsp_sparse_batch(500, 50, B, seed=1234)
E = ei.shape[1]
enc = G.encoder(w, 2, impl="tc")
ctx = enc.set_graph(G.cu(ei), pts.shape[0], 1)
enc.set_points(G.cu(pts))
xt = G.cu((syn.initial_noise(E, 0) > 0).astype(np.float32))
out = torch.empty((E, 2), device="cuda")
st = torch.cuda.current_stream().cuda_stream


def t_forward(tag):
  ctx.encoder_forward(xt.data_ptr(), 500.0, out.data_ptr(), st)
  torch.cuda.synchronize()
  ctx.profile_begin()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  reps = int(os.environ.get("PROBE_REPS", "1"))
  e0.record()
  for _ in range(reps):
    ctx.encoder_forward(xt.data_ptr(), 500.0, out.data_ptr(), st)
  e1.record()
  torch.cuda.synchronize()
  ms, n = ctx.profile_end()
  print(f"{tag}: checksum {out.double().sum().item():.10e} {out.double().abs().sum().item():.10e}")
  print(f"{tag}: forward {e0.elapsed_time(e1) / reps:.2f} ms; edge kernel {ms / max(n, 1):.3f} ms/launch x{n}", flush=True)
  if int(os.environ.get("DFB_TC_PROBE", "0")) & 128:
    pc = ctx.debug_phase_cycles()
    ntile_cta = (E + 127) // 128 * 12 * 2      # tiles x 12 layers x (warm-up + timed forward)
    names = ["conv", "waitG1", "E1", "E2", "E3", "waitG2", "E4"]
    print(f"{tag}: cycles/tile " + " ".join(f"{n}={c / ntile_cta:.0f}" for n, c in zip(names, pc)) +
          f" total={sum(pc[:7]) / ntile_cta:.0f}", flush=True)
    npair = (E + 127) // 128 * 11 * 2          # pair kernel: middle layers only
    pn = ["X(conv+E4)", "waitG1", "E1", "E2", "E3"]
    if sum(pc[16:24]):
      print(f"{tag}: pair kernel cycles/tile " + " ".join(f"{n}={c / npair:.0f}" for n, c in zip(pn, pc[16:21])) +
            f" total={sum(pc[16:21]) / npair:.0f}", flush=True)
      xn = ["boxwait", "conv", "waitG2/stage", "E4copy", "preload", "setup", "stagewait"]
      print(f"{tag}: pair kernel X sub-phases cycles/tile " + " ".join(f"{n}={c / npair:.0f}" for n, c in zip(xn, pc[24:31])), flush=True)
      nlead = npair / 2   # tile pairs: the leader CTA's MMA thread
      print(f"{tag}: pair kernel MMA-thread waits per tile pair: A stages {pc[21] / nlead:.0f}  weights {pc[22] / nlead:.0f}  "
            f"GEMM2 A chunks {pc[23] / nlead:.0f}", flush=True)
    L = _cabi.lib()
    if hasattr(L, "dfb_debug_pair_trace"):   # --prof build: time line of cluster 0's leader CTA, tiles 3..6 (last launch)
      import ctypes as C
      buf = (C.c_longlong * 512)()
      L.dfb_debug_pair_trace(buf)
      tr = np.array(buf[:], dtype=np.int64).reshape(4, 8, 16)
      t0 = tr[0, 0, 0]
      wn = ["Xstart", "box0", "conv0", "G2/st0", "E4_0", "pre0", "box1", "conv1", "st1", "E4_1", "pre1", "setup", "gath", "G1", "E1", "E3"]
      for itx in range(4):
        for a in range(4):
          print(f"{tag}: trace tile {itx + 3} part {a}: " + " ".join(f"{n}={tr[itx, a, k] - t0}" for k, n in enumerate(wn)), flush=True)
        print(f"{tag}: trace tile {itx + 3} MMA: G1 chunks " + " ".join(str(tr[itx, 4, k] - t0) for k in range(8)) +
              " | G2 chunks " + " ".join(str(tr[itx, 4, 8 + k] - t0) for k in range(4)) + f" | end {tr[itx, 4, 12] - t0}", flush=True)
        print(f"{tag}: trace tile {itx + 3} box loads: " + " ".join(str(tr[itx, 5, k] - t0) for k in range(8)) +
              " | stores: " + " ".join(str(tr[itx, 6, k] - t0) for k in range(8)), flush=True)
    sub = ["tmemwait", "gatherwait", "math", "gissue", "reduce"]
    print(f"{tag}: E1 sub-phases cycles/tile " + " ".join(f"{n}={c / ntile_cta:.0f}" for n, c in zip(sub, pc[8:13])), flush=True)


def t_gemm(tag):
  x = torch.randn(E, 256, device="cuda")
  acc = torch.empty(E, 256, device="cuda")
  ctx.debug_edge_gemm(1, x.data_ptr(), acc.data_ptr(), st)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(3):
    ctx.debug_edge_gemm(1, x.data_ptr(), acc.data_ptr(), st)
  e1.record()
  torch.cuda.synchronize()
  print(f"{tag}: GEMM1-only (convert + 48 MMA/tile + dump) {e0.elapsed_time(e1) / 3:.3f} ms/launch", flush=True)


mode = sys.argv[1] if len(sys.argv) > 1 else "all"
if mode in ("all", "gemm"):
  t_gemm(f"probe={os.environ.get('DFB_TC_PROBE', '0')}")
if mode in ("all", "fwd"):
  t_forward(f"probe={os.environ.get('DFB_TC_PROBE', '0')}")
