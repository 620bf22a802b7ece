"""Host-side multi-rank logic (CPU, gloo, world_size 2): env sharding and the episode-statistics gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from so101_sim_b200.sharding import EpisodeStats, gather_episode_stats, rank_seed, shard_range


def test_shard_range_partitions_envs():
  for total in (0, 1, 7, 16, 131072):
    for world in (1, 2, 3, 8):
      spans = [shard_range(total, r, world) for r in range(world)]
      assert spans[0][0] == 0 and spans[-1][1] == total
      assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
      sizes = [b - a for a, b in spans]
      assert max(sizes) - min(sizes) <= 1
  assert shard_range(131072, 3, 8) == (49152, 65536)
  with pytest.raises(ValueError):
    shard_range(8, 2, 2)
  assert rank_seed(5, 0) != rank_seed(5, 1)


def test_episode_stats_follow_step_types():
  s = EpisodeStats(3, 'cpu')
  first = torch.tensor([0, 0, 0], dtype=torch.uint8)
  s.update(first, torch.zeros(3))
  s.update(torch.tensor([1, 1, 1], dtype=torch.uint8), torch.tensor([0.0, 0.0, 0.0]))
  s.update(torch.tensor([1, 2, 1], dtype=torch.uint8), torch.tensor([0.0, 1.0, 0.0]))   # env 1 succeeds on its 2nd step
  s.update(torch.tensor([1, 0, 2], dtype=torch.uint8), torch.tensor([0.0, 0.0, 0.0]))   # env 1 auto-resets, env 2 hits the time limit
  loc = s.local()
  assert loc[1].tolist() == [1.0, 2.0, 1.0, 1.0]
  assert loc[2].tolist() == [0.0, 3.0, 0.0, 1.0]
  assert loc[0].tolist() == [0.0, 0.0, 0.0, 0.0]
  assert s.length.tolist() == [3, 0, 3]


def _worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    lo, hi = shard_range(10, rank, world)
    stats = torch.arange(lo, hi, dtype=torch.float32).unsqueeze(1).repeat(1, 3)
    stats[:, 1] = rank
    out = gather_episode_stats(stats)
    q.put((rank, out.tolist()))
  finally:
    dist.destroy_process_group()


def test_gather_episode_stats_gloo_world2():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs: p.start()
  got = dict(q.get(timeout=120) for _ in range(2))
  for p in procs: p.join(timeout=60)
  assert all(p.exitcode == 0 for p in procs)
  expect = [[float(i), float(i // 5), float(i)] for i in range(10)]
  assert got[0] == expect and got[1] == expect
