"""Multi-GPU layout of the lockstep env step: environments shard across ranks, nothing else does.

One process per GPU (`torchrun`); rank r owns envs [r * N / W, (r + 1) * N / W) and a full copy of the (< 2 MB) model.  The
step itself needs NO collective — envs never interact.  The only exchange on this path is the gather of
episode statistics (return, length, success), off the step's critical path: `all_gather_into_tensor` over NCCL on the
GPUs (NVLink 5 / NVSwitch), gloo in the CPU tests.  The reference has no counterpart (single process, single env).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world: int) -> tuple[int, int]:
  """[start, stop) of the envs rank `rank` owns; the first `total % world` ranks take one extra env."""
  if world <= 0 or not 0 <= rank < world:
    raise ValueError(f'invalid rank {rank} for world size {world}')
  if total_envs < 0:
    raise ValueError('total_envs must be >= 0')
  base, extra = divmod(total_envs, world)
  start = rank * base + min(rank, extra)
  return start, start + base + (1 if rank < extra else 0)


def rank_seed(seed: int, rank: int) -> int:
  """Per-rank Philox stream so that shards draw independent initial states / actions."""
  return int(seed) + 1000 * int(rank)


class EpisodeStats:
  """Per-env running episode return / length / success on the env's device, closed at LAST steps (so100_task.py:292-302)."""

  def __init__(self, num_envs: int, device):
    self.ret = torch.zeros(num_envs, dtype=torch.float32, device=device)
    self.length = torch.zeros(num_envs, dtype=torch.int32, device=device)
    self.done_return = torch.zeros(num_envs, dtype=torch.float32, device=device)
    self.done_length = torch.zeros(num_envs, dtype=torch.int32, device=device)
    self.done_success = torch.zeros(num_envs, dtype=torch.uint8, device=device)
    self.episodes = torch.zeros(num_envs, dtype=torch.int32, device=device)

  def update(self, step_type: torch.Tensor, reward: torch.Tensor):
    """Feed one BatchedTimeStep: FIRST (0) restarts the accumulators, LAST (2) closes the episode."""
    first, last = step_type == 0, step_type == 2
    self.ret = torch.where(first, torch.zeros_like(self.ret), self.ret + reward)
    self.length = torch.where(first, torch.zeros_like(self.length), self.length + 1)
    self.done_return = torch.where(last, self.ret, self.done_return)
    self.done_length = torch.where(last, self.length, self.done_length)
    self.done_success = torch.where(last, (reward >= 1.0).to(torch.uint8), self.done_success)
    self.episodes += last.to(torch.int32)

  def local(self) -> torch.Tensor:
    """[N, 4] float32: last finished episode's return, length, success, and the number of finished episodes."""
    return torch.stack([self.done_return, self.done_length.float(), self.done_success.float(), self.episodes.float()], dim=1).contiguous()


def gather_episode_stats(stats: torch.Tensor) -> torch.Tensor:
  """All ranks receive the [W * N, k] concatenation (rank order) of their [N, k] statistics.  Shards must be equal-sized."""
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
    return stats
  world = dist.get_world_size()
  out = torch.empty((world * stats.shape[0],) + tuple(stats.shape[1:]), dtype=stats.dtype, device=stats.device)
  dist.all_gather_into_tensor(out, stats.contiguous())
  return out
