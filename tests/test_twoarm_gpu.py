"""BASELINE config 4 ("hand-over task, two arms, contact-heavy constraint solve"): the labelled synthetic two-SO100 scene
(so101_sim_b200.task_suite.SO100TwoArmHandOver; the reference has no two-SO100 scene) through the same kernels built for two
arms (libso101_b200_2arm.so), against the float64 oracle - which is generic over joints and bodies - and against the
reference's own arm known-answer (KAT-1, so101_rl.ipynb:219-229), which must hold for EACH of the two identical arms."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.oracle import OracleSim

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
KAT1 = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'kat1_so101_rl.json')))


def _env(built, **kw):
  from so101_sim_b200.task_suite import create_batched_task_env
  args = dict(task_name='SO100TwoArmHandOverBanana', num_envs=2, time_limit=30.0, seed=0, device=DEV, reset_rounds=0)
  args.update(kw)
  return create_batched_task_env(**args)


def _state(env, n, seed=0, clearance=0.002):
  rs = np.random.RandomState(seed)
  q = np.tile(np.asarray(env.model['qpos0'], dtype=np.float64), (n, 1))
  q[:, :12] = 0
  for e in range(n):
    yaw = rs.uniform(-0.3, 0.3)
    q[e, 12:19] = [rs.uniform(0.2, 0.3), rs.uniform(-0.1, 0.1), 0.4217 + clearance, np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)]
    q[e, 19:26] = [rs.uniform(-0.3, -0.22), rs.uniform(-0.1, -0.02), 0.4226 + clearance, 1, 0, 0, 0]
  return torch.tensor(q), torch.zeros(n, 24, dtype=torch.float64)


def _actions(env, steps, seed=1, scale=0.3):
  g = torch.Generator(device=DEV); g.manual_seed(seed)
  spec = env.action_spec()
  lo, hi = torch.tensor(spec.minimum, device=DEV), torch.tensor(spec.maximum, device=DEV)
  return (lo + torch.rand(steps, env.num_envs, env.nu, generator=g, device=DEV) * (hi - lo)) * scale


@pytest.mark.parametrize('precision,rtol', [('f64', 1e-8), ('f32', 2e-5)])
def test_kat1_holds_for_each_arm(built, precision, rtol):
  env = _env(built, precision=precision, calibration_offsets=KAT1['calibration_offsets'])
  assert env.nq == 26 and env.nv == 24 and env.nu == 12 and env.action_spec().shape == (12,)
  q, v = _state(env, 2, clearance=0.0005)
  env.set_initial_state(q, v)
  ts = env.reset()
  assert ts.observation['joints_pos'].shape == (2, 12) and ts.observation['physics_state'].shape == (2, 50)
  act = torch.tensor([KAT1['action'] * 2] * 2, dtype=torch.float32, device=DEV)
  ts = env.step(act)
  np.testing.assert_allclose(ts.observation['commanded_joints_pos'][0].cpu().numpy(), KAT1['commanded_joints_pos'] * 2)
  qq, vv = env.get_state(torch.float64)
  want_q, want_v = np.array(KAT1['physics_state'][:6]), np.array(KAT1['physics_state'][20:26])
  for arm in range(2):
    np.testing.assert_allclose(qq[0, 6 * arm:6 * arm + 6].cpu().numpy(), want_q, rtol=rtol)
    np.testing.assert_allclose(vv[0, 6 * arm:6 * arm + 6].cpu().numpy(), want_v, rtol=rtol)
  env.close()


def test_f64_two_arm_scene_matches_oracle(built):
  """Both arms driven by random targets while the props land on the table: float64 CUDA path vs the float64 oracle, rewards
  exact.  Up to the first DEEP arm-arm interpenetration (> 2 mm: the arms are driven into each other) the states agree to 1e-6;
  such a hit is an ill-conditioned EPA problem (an 8-vertex jaw hull buried in a 525-vertex link hull: a 1e-9 perturbation of the
  poses moves the minimum-translation face, measured 3.52 vs 3.70 mm), after which trajectories separate - the same
  sensitivity MuJoCo's own tolerance-terminated EPA has."""
  env = _env(built, precision='f64')
  q0, v0 = _state(env, 2, seed=3)
  env.set_initial_state(q0, v0); env.reset()
  acts = _actions(env, 25, seed=4)
  sims = []
  for e in range(2):
    o = OracleSim('so100_twoarm_banana', collide=True)
    o.set_state(q0[e].numpy(), v0[e].numpy())
    sims.append(o)
  meta = sims[0].meta
  arm_bodies = set(int(b) for b in meta['jnt_body'][:12])
  gbody = meta['geom_body']
  clean = [True, True]
  worst, ncon_max, clean_steps = 0.0, 0, 0
  for t in range(25):
    ts = env.step(acts[t])
    q, v = env.get_state(torch.float64)
    for e, o in enumerate(sims):
      r = o.control_step(acts[t, e].double().cpu().numpy())
      assert float(ts.reward[e]) == r
      for c in o.contacts():
        if int(gbody[c['geom1']]) in arm_bodies and int(gbody[c['geom2']]) in arm_bodies and c['dist'] < -2e-3:
          clean[e] = False
      if clean[e]:
        worst = max(worst, float(np.abs(q[e].cpu().numpy() - o.qpos).max()))
        clean_steps += 1
    ncon_max = max(ncon_max, int(env.debug_read('ncon').max()))
  print('two-arm f64 rollout: max |dqpos| before a deep arm-arm hit =', worst, 'over', clean_steps, 'env-steps; max contacts', ncon_max)
  assert worst < 1e-6 and clean_steps >= 12 and ncon_max >= 20
  assert env.counters()['diverged'] == 0 and env.counters()['contacts_dropped'] == 0
  env.close()


def test_f32_two_arm_default_creation_and_rollout(built):
  """Factory default (on-device placements + nursery), float32 product arithmetic: 30 random-action steps of 64 envs; one env's
  arms are tracked against the oracle for as long as they are contact-free (north_star tolerance 1e-4)."""
  env = _env(built, num_envs=64, reset_rounds=1)
  ts = env.reset()
  ps = ts.observation['physics_state'].double()
  assert float(ps[:, :12].abs().max()) == 0.0 and abs(float(ps[:, 14].median()) - 0.4217) < 3e-3
  q0, v0 = env.get_state(torch.float64)
  o = OracleSim('so100_twoarm_banana', collide=True)
  o.set_state(q0[0].cpu().numpy(), v0[0].cpu().numpy())
  arm_bodies = set(int(b) for b in o.meta['jnt_body'][:12])
  gbody = o.meta['geom_body']
  acts = _actions(env, 30, seed=2, scale=0.1)
  err, err_touch, free, free_steps, first = 0.0, 0.0, True, 0, None
  for t in range(30):
    ts = env.step(acts[t])
    o.control_step(acts[t, 0].double().cpu().numpy())
    touching = [(c['geom1'], c['geom2']) for c in o.contacts() if int(gbody[c['geom1']]) in arm_bodies or int(gbody[c['geom2']]) in arm_bodies]
    if touching and free:
      free, first = False, (t + 1, touching[0])
    q, _ = env.get_state(torch.float64)
    e = float(np.abs(q[0, :12].cpu().numpy() - o.qpos[:12]).max())
    if free:
      err = max(err, e); free_steps += 1
    elif t < 15:
      err_touch = max(err_touch, e)
  print('two-arm f32: arm error', err, 'over', free_steps, 'contact-free steps; first arm contact (step, geoms):', first, '; error with arm contacts, first 15 steps:', err_touch)
  assert err < 1e-4 and free_steps >= 1 and err_touch < 5e-3, (err, free_steps, err_touch)
  assert torch.isfinite(ts.observation['physics_state']).all() and env.counters()['diverged'] == 0
  assert ts.step_type.eq(1).all()
  env.close()
