"""BASELINE.json's full sizes (config 2: 4096 arm-only envs, config 3: 16384 pick-place envs) through size-independent
properties: the oracle cannot step thousands of envs in test time, so these check that an env's trajectory does not depend
on the batch it runs in (same states + actions in a small and in the full-size batch -> bitwise equal), that identical envs
stay identical wherever they sit in the batch, physical invariants over the whole batch, and a sample of envs against the
oracle.  Everything goes through the C-ABI (BatchedEnvironment)."""
import numpy as np
import pytest
import torch

from oracle.oracle import OracleSim

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _make(task, n, **kw):
  from so101_sim_b200.task_suite import create_batched_task_env
  return create_batched_task_env(task, num_envs=n, time_limit=30.0, seed=0, device=DEV, reset_rounds=0, **kw)


def _actions(env, steps, n, seed=1, scale=0.3):
  g = torch.Generator(device=DEV); g.manual_seed(seed)
  spec = env.action_spec()
  lo, hi = torch.tensor(spec.minimum, device=DEV), torch.tensor(spec.maximum, device=DEV)
  return (lo + torch.rand(steps, n, 6, generator=g, device=DEV) * (hi - lo)) * scale


def test_arm4096_batch_independence_and_oracle_sample(built):
  N, steps = 4096, 20
  env = _make('SO100ArmOnly', N)
  q0, v0 = env.sample_arm_initial_states(seed=0)
  # the second half of the batch repeats the first half: copies must stay bitwise identical
  q0[N // 2:] = q0[:N // 2]
  env.set_initial_state(q0, v0); env.reset()
  acts = _actions(env, steps, N)
  acts[:, N // 2:] = acts[:, :N // 2]
  small = _make('SO100ArmOnly', 32)
  small.set_initial_state(q0[:32], v0[:32]); small.reset()
  for t in range(steps):
    ts = env.step(acts[t]); small.step(acts[t, :32])
  q, v = env.get_state(); qs, vs = small.get_state()
  assert torch.equal(q[:N // 2], q[N // 2:]) and torch.equal(v[:N // 2], v[N // 2:])
  assert torch.equal(q[:32], qs) and torch.equal(v[:32], vs)
  assert torch.isfinite(q).all() and ts.step_type.eq(1).all() and ts.reward.eq(0).all() and ts.discount.eq(1).all()
  assert torch.equal(ts.observation['undelayed_joints_pos'], q[:, :6])
  for e in (0, 1777, N // 2 - 1):
    o = OracleSim('so100_arm', collide=False)
    o.set_state(q0[e].double().cpu().numpy(), np.zeros(6))
    for t in range(steps):
      o.control_step(acts[t, e].double().cpu().numpy())
    assert np.abs(o.qpos - q[e].double().cpu().numpy()).max() < 1e-4, e   # stated float32 tolerance (north_star)
  env.close(); small.close()


def test_banana16384_batch_independence_and_invariants(built):
  N, steps = 16384, 4
  env = _make('SO100HandOverBanana', N)
  q0, v0 = env.sample_prop_initial_states(seed=0, settle_steps=5)
  # 64 distinct placements (the first 64 envs), each repeated 256 times across the batch
  idx = torch.arange(N, device=DEV) % 64
  q0, v0 = q0[idx].contiguous(), v0[idx].contiguous()
  env.set_initial_state(q0, v0)
  ts0 = env.reset()
  first = ts0.observation['physics_state'].clone()
  ts0 = env.reset()
  assert torch.equal(first, ts0.observation['physics_state']) and ts0.step_type.eq(0).all()      # reset is idempotent
  acts = _actions(env, steps, N, scale=0.2)[:, idx].contiguous()
  small = _make('SO100HandOverBanana', 8)
  small.set_initial_state(q0[:8], v0[:8]); small.reset()
  for t in range(steps):
    ts = env.step(acts[t]); small.step(acts[t, :8])
  q, v = env.get_state(); qs, vs = small.get_state()
  c = env.counters()
  assert c['diverged'] == 0 and c['contacts_dropped'] == 0
  assert torch.equal(q[:8], qs) and torch.equal(v[:8], vs)                                           # batch-size independence
  rep = q.reshape(N // 64, 64, -1)
  assert torch.equal(rep, rep[:1].expand_as(rep)), 'copies of the same env diverged across the batch'  # position independence
  assert torch.isfinite(q).all() and torch.isfinite(v).all()
  for a in (9, 16):                                                                                   # unit quaternions
    assert float((q[:, a:a + 4].norm(dim=1) - 1).abs().max()) < 1e-5
  assert float(q[:, 8].min()) > 0.415 and float(q[:, 15].min()) > 0.415                               # props rest on the table top (z 0.42)
  assert ts.step_type.eq(1).all() and ts.discount.eq(1).all() and ((ts.reward == 0) | (ts.reward == 1)).all()
  assert torch.equal(ts.observation['physics_state'], torch.cat([q, v], dim=1))
  assert float(ts.observation['joints_pos'].abs().max()) == 0.0                                       # still the initial value: 5-step delay
  # one env of the full batch against the oracle over the same steps (float32 product path: stated tolerance on the arm)
  o = OracleSim('so100_handover_banana', collide=True)
  o.set_state(q0[5].double().cpu().numpy(), v0[5].double().cpu().numpy())
  for t in range(steps):
    o.control_step(acts[t, 5].double().cpu().numpy())
  assert np.abs(o.qpos[:6] - q[5, :6].double().cpu().numpy()).max() < 1e-4
  assert np.abs(o.qpos[6:] - q[5, 6:].double().cpu().numpy()).max() < 2e-3
  env.close(); small.close()
