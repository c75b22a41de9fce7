"""world_size-2 gloo test of the N>1 host logic (sharding + ragged heatmap gather).  CPU only."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from difusco_b200.distributed import denoise_sharded, gather_heatmaps, shard_range


def test_shard_range_partitions_everything():
  for n in (0, 1, 7, 16, 129):
    for world in (1, 2, 4, 8):
      cover = []
      for r in range(world):
        lo, hi = shard_range(n, r, world)
        assert 0 <= lo <= hi <= n
        cover += list(range(lo, hi))
      assert cover == list(range(n))
      sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
      assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    shard_range(4, 2, 2)


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    # 5 instances with different "edge counts"; the fake engine returns instance id + position
    inst = [{"id": i, "E": 3 + 2 * i} for i in range(5)]

    def run_batch(block):
      return torch.cat([torch.full((b["E"],), float(b["id"])) + torch.arange(b["E"]) * 1e-3 for b in block])
    got = denoise_sharded(inst, run_batch, batch=2)
    flat = torch.cat(got)
    ref = run_batch(inst)
    ok = torch.equal(flat, ref) and len(got) == world
    # explicit sizes path
    lo, hi = shard_range(5, rank, world)
    mine = run_batch(inst[lo:hi]) if hi > lo else torch.zeros(0)
    sizes = [sum(b["E"] for b in inst[slice(*shard_range(5, r, world))]) for r in range(world)]
    ok = ok and torch.equal(torch.cat(gather_heatmaps(mine, sizes=sizes)), ref)
    # more ranks than instances: the rank with the empty shard still joins the collectives (device / dtype explicit)
    one = denoise_sharded(inst[:1], run_batch, batch=2, device=torch.device("cpu"), dtype=torch.float32)
    ok = ok and len(one) == world and torch.equal(torch.cat(one), run_batch(inst[:1])) and one[-1].dtype == torch.float32
    # sync_dist=True metric mean across ranks (pl_tsp_model.py:253-255) without Lightning
    from difusco_b200.pl_meta_model import _Base
    if hasattr(_Base, "logged_metrics"):
      m = _Base()
      for v in ([1.0, 3.0] if rank == 0 else [5.0, 7.0]):
        m.log("test/solved_cost", v, on_epoch=True, sync_dist=True)
      m.log("test/local_only", float(rank + 1))
      got = m.logged_metrics()
      ok = ok and abs(got["test/solved_cost"] - 4.0) < 1e-12 and abs(got["test/local_only"] - (rank + 1)) < 1e-12
    q.put((rank, bool(ok)))
  finally:
    dist.destroy_process_group()


def test_gloo_world2_gather_matches_single_process():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = sorted(q.get(timeout=120) for _ in procs)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert res == [(0, True), (1, True)]
