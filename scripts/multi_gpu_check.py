"""N-GPU correctness check: the heat map of every instance must be BITWISE identical no matter how many
ranks share the work - instances never interact across GPUs, sampling is Philox keyed by (seed, step, element).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/multi_gpu_check.py

Every rank runs its contiguous shard through difusco_b200.distributed.denoise_sharded (NCCL all_gather of the heat maps);
rank 0 also computes ALL instances alone (the world-size-1 answer) and compares."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from difusco_b200 import synthetic as syn
from difusco_b200.distributed import denoise_sharded
import gpu_util as G


def main():
  rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
  torch.cuda.set_device(local)
  dist.init_process_group("nccl", device_id=torch.device("cuda", local))
  w = syn.make_encoder_weights(0, out_channels=2)
  steps = 10
  m = G.tsp_model(w, "tc", sparse_factor=50, inference_diffusion_steps=steps)
  n_inst = 2 * world + 1                       # ragged: the last rank gets fewer instances
  inst = []
  for i in range(n_inst):
    pts, ei = syn.tsp_sparse_batch(500, 50, 1, seed=100 + i)
    xt0 = (syn.initial_noise(ei.shape[1], i) > 0).astype(np.float32)
    inst.append((i, pts, ei, xt0))

  def run_batch(block):     # one instance per call: the head GroupNorm couples whatever shares a call
    outs = []
    for (i, pts, ei, xt0) in block:
      outs.append(m.denoise_heatmap(G.cu(pts), G.cu(ei), G.cu(xt0), seed=1000 + i).reshape(-1))
    return torch.cat(outs)

  got = denoise_sharded(inst, run_batch, batch=1)
  flat = torch.cat(got).cpu().numpy()
  ok = True
  if rank == 0:
    ref = run_batch(inst).cpu().numpy()
    ok = bool(np.array_equal(flat, ref))
    print(json.dumps({"world": world, "instances": n_inst, "edges_per_instance": int(inst[0][2].shape[1]), "steps": steps,
                      "bitwise_equal_to_single_gpu": ok, "max_abs_diff": float(np.abs(flat - ref).max()),
                      "heatmap_mean": float(ref.mean())}))
  flag = torch.tensor([1 if ok else 0], device="cuda")
  dist.broadcast(flag, 0)
  dist.destroy_process_group()
  sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
  main()
